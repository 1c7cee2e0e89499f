"""The C-ABI library builds for sm_100a, loads, and exports exactly what include/audio_modem_b200.h declares.
No compute calls here (no GPU); amb_create must refuse to run without a device instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "audio_modem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    for n in ("amb_detect_preamble", "amb_detect_preamble_xcorr", "amb_preprocess_signal", "amb_estimate_channel", "amb_demodulate",
              "amb_crc32", "amb_decode_received_signals", "amb_decode_chunk_frames", "amb_estimate_frame_samples",
              "amb_stream_open", "amb_stream_push", "amb_stream_close", "amb_create", "amb_destroy"):
        assert n in names


def test_library_exports_every_declared_symbol(amb):
    lib = C.CDLL(amb._lib.lib_path())
    for n in declared_functions():
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the ctypes binding covers the same set
    assert sorted(amb._lib.EXPORTS) == declared_functions()


def test_struct_layout_matches_header(amb):
    # 2 int32 + int64 + 14 int32/uint32 + 2 doubles = 88 bytes, no implicit padding
    assert C.sizeof(amb.FrameResult) == 88
    assert amb.FrameResult.preamble_idx.offset == 8 and amb.FrameResult.coarse_metric.offset == 72


def test_sass_is_sm100_and_has_no_tensor_core_ops(amb):
    out = subprocess.run(["cuobjdump", "-sass", amb._lib.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100" in out.stdout
    # nothing on this path is a dense contraction: no HMMA / UTC*MMA / GMMA may appear
    assert not re.search(r"\b(HMMA|UTC[A-Z]*MMA|HGMMA|IGMMA|QGMMA)\b", out.stdout)
    # the scanner front end stages its input tiles with TMA bulk copies (cp.async.bulk -> UBLKCP) completing on an mbarrier
    assert re.search(r"\bUBLKCP\b", out.stdout) and re.search(r"\bSYNCS\b", out.stdout)


def test_host_helpers_need_no_gpu(amb, O):
    L = amb._lib.load()
    assert L.amb_version() == 100
    for mode, (cfg, mod) in {"a": ("standard", "QAM16"), "b": ("acoustic", "BPSK"), "c": ("narrowband", "BPSK")}.items():
        for payload in (11, 27, 280, 523, 4107, 4112):
            for rep in (1, 3):
                assert L.amb_estimate_frame_samples(O.CONFIGS[cfg], O.MODS[mod], payload, rep) == O.estimateFrameSamples(cfg, payload, mod, rep)
    assert L.amb_status_string(2).decode() == "Preamble not detected (low correlation)"


def test_no_cpu_fallback(amb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(amb.AmbError, match="no CUDA device"):
        amb.Modem("standard", "QPSK", 1)
    with pytest.raises(amb.AmbError):
        amb.detectPreamble([0.0] * 1024)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "audio-modem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "modem_oracle" not in text, f


# ---------------------------------------------------------------------------------------------- the boundary from C and from the N-API addon
BOUNDARY_JS = ["setOFDMConfig", "detectPreamble", "detectPreambleCrossCorr", "preprocessSignal", "estimateChannel", "demodulateOFDM", "modulateOFDM",
               "crc32", "decodeReceivedSignal", "decodeChunkFrame", "estimateFrameSamples", "buildTransmitSignal", "buildMetadataFrame",
               "buildDataChunkFrame", "StreamingReceiver"]      # SURVEY.md section 8(b), "Signatures the drop-in keeps"


def build_cabi_smoke(amb, tmp_path):
    """tests/cabi_smoke.c compiled as strict C99 against the public header and linked to the library"""
    exe = str(tmp_path / "cabi_smoke")
    lib_dir = os.path.dirname(amb._lib.lib_path())
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cabi_smoke.c"), "-L", lib_dir, "-laudiomodem_b200", "-Wl,-rpath," + lib_dir, "-o", exe])
    return exe


def test_header_is_c_and_the_library_links_from_c(amb, tmp_path):
    """the public header compiles as C99 (no C++ in the boundary) and a C program linked against the library runs its host-side
    entry points; without a GPU amb_create refuses (AMB_E_NODEVICE) instead of falling back"""
    exe = build_cabi_smoke(amb, tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "host side ok" in out.stdout


def test_napi_addon_compiles_and_covers_the_boundary(amb, tmp_path):
    """node/binding.cc is compiled (-Wall -Wextra -Werror) against a declaration-only node_api.h (tests/stubs) and linked against
    the library: every amb_* symbol it calls exists with the signature the header declares, and it registers every entry point
    the JS wrapper uses.  (Node.js itself is not in the image: the addon is never loaded here.)"""
    obj, so = str(tmp_path / "binding.o"), str(tmp_path / "audio_modem_b200.node")
    lib_dir = os.path.dirname(amb._lib.lib_path())
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-Wno-comment", "-fPIC", "-c", "-I", os.path.join(ROOT, "tests", "stubs"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "node", "binding.cc"), "-o", obj])
    subprocess.check_call(["g++", "-shared", obj, "-L", lib_dir, "-laudiomodem_b200", "-o", so])
    undefined = subprocess.run(["nm", "-u", obj], capture_output=True, text=True).stdout
    called = sorted(set(re.findall(r"\b(amb_[a-z0-9_]+)\b", undefined)))
    assert len(called) >= 20 and set(called) <= set(declared_functions()), set(called) - set(declared_functions())
    src = open(os.path.join(ROOT, "node", "binding.cc")).read()
    registered = set(re.findall(r'\{"([A-Za-z0-9]+)", nullptr, ', src))
    js = open(os.path.join(ROOT, "node", "modem_b200.js")).read()
    used = set(re.findall(r"addon\.([A-Za-z0-9]+)\(", js))
    assert used <= registered, used - registered
    exported = set(re.findall(r"^  ([A-Za-z0-9]+)(?:\(|:)", js, flags=re.M))
    assert set(BOUNDARY_JS) <= exported, set(BOUNDARY_JS) - exported
