"""Shared test helpers: synthetic signal builders (through the oracle's TX) and result comparison."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the reference's five UI modes (app.js:60-66) with their chunk sizes (app.js:195-199)
MODES = {
    "QPSK": ("standard", "QPSK", 1, 2048),
    "16-QAM": ("standard", "QAM16", 1, 4096),
    "BPSK-ACOUSTIC": ("acoustic", "BPSK", 1, 512),
    "BPSK-REPEAT": ("acoustic", "BPSK", 3, 512),
    "BPSK-NARROW": ("narrowband", "BPSK", 3, 512),
}

RESULT_INT_FIELDS = ["status", "frame_type", "preamble_idx", "seq_num", "data_len", "total_chunks", "total_file_size",
                     "chunk_size", "name_len", "name_off", "data_off", "expected_crc", "actual_crc", "crc_valid", "n_bytes",
                     "err_arg"]


def load_emu():
    """Build and load the test-only host emulation of the device logic (tests/emu/emu.cpp)."""
    d = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(d, "libamb_emu.so")
    srcs = [os.path.join(d, "emu.cpp"), os.path.join(ROOT, "audio-modem_b200", "csrc", "amb_tables.cpp")]
    hdrs = [os.path.join(ROOT, "audio-modem_b200", "csrc", f) for f in ("amb_device.cuh", "amb_walker.cuh", "amb_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-DAMB_EMU_CHECK", "-shared", "-o", so] + srcs)
    return C.CDLL(so)


def pad_blocks(x, block=4096, tail=8192):
    n = (len(x) + tail + block - 1) // block * block
    return np.concatenate([np.asarray(x, np.float32), np.zeros(n - len(x), np.float32)])


def make_stream(O, mode, n_frames, seed, snr_db=25.0, lead=0, gaps=None, meta=True, noise=True):
    """metadata frame + n_frames data chunks (buildMetadataFrame / buildDataChunkFrame), AWGN at snr_db relative to
    the first data frame's core; returns (samples padded to whole blocks, payload list)."""
    cfg, mod, rep, chunk = MODES[mode]
    parts = [np.zeros(lead, np.float32)]
    if meta:
        parts.append(O.buildMetadataFrame(cfg, n_frames, n_frames * chunk, chunk, "file.bin", mod, rep))
    payloads = []
    first = None
    for i in range(n_frames):
        d = O.randomBytes(seed * 1000 + i, chunk)
        payloads.append(bytes(d))
        f = O.buildDataChunkFrame(cfg, d, i, mod, rep)
        first = f if first is None else first
        parts.append(f)
        if gaps:
            parts.append(np.zeros(gaps[i % len(gaps)], np.float32))
    x = np.concatenate(parts)
    if noise and first is not None:
        p = O.power(first[2205:2205 + O.estimateFrameSamples(cfg, chunk + 11, mod, rep)])
        x = O.addAwgn(x, np.sqrt(p / 10 ** (snr_db / 10)), seed + 77)
    return pad_blocks(x), payloads


def results_equal(r1, b1, r2, b2, metric_tol=1e-6, check_metrics=True):
    """Field-by-field comparison of two frame results (+ decoded bytes); returns a list of differences."""
    diffs = []
    for n in RESULT_INT_FIELDS:
        a, b = getattr(r1, n), getattr(r2, n)
        if a != b:
            diffs.append((n, a, b))
    if check_metrics:
        for n in ("coarse_metric", "fine_metric"):
            a, b = getattr(r1, n), getattr(r2, n)
            if not (abs(a - b) <= metric_tol * max(1.0, abs(a))):
                diffs.append((n, a, b))
    # decoded bytes: everything the parser consumed (header, name, payload, CRC) must be identical; bytes beyond the frame's
    # extent come from symbols after the frame (noise in the rest of the collection window), which no result depends on and
    # the CUDA path does not re-decide in float64 (include/audio_modem_b200.h: amb_frame_result)
    b1 = np.asarray(b1, np.uint8)
    b2 = np.asarray(b2, np.uint8)
    ext = min(frame_extent(r1), len(b1))         # (a caller may hand in rows clipped to its own row length)
    if len(b2) < ext or not np.array_equal(b1[:ext], b2[:ext]):
        diffs.append(("bytes", len(b1), len(b2), ext))
    return diffs


def frame_extent(r):
    """number of decoded bytes the reference's parsers consume for this result (0 for an error result)"""
    if r.status != 0:
        return 0
    if r.frame_type == 0xFE:
        return r.name_off + r.name_len + 4
    return r.data_off + r.data_len + 4


def P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def emu_decode_chunk_frame(emu, O, cfg, mod, rep, x):
    x = np.ascontiguousarray(x, np.float32)
    r = O.OrcResult()
    b = np.zeros(70000, np.uint8)
    emu.emu_decode_chunk_frame(O.CONFIGS[cfg], O.MODS[mod], rep, P(x, C.c_float), C.c_int64(len(x)), C.byref(r), P(b, C.c_uint8), C.c_int64(len(b)))
    return r, b[: max(0, r.n_bytes)]


def emu_stream(emu, O, cfg, mod, rep, x, block=4096, push=None, stride=70000, max_res=512):
    x = np.ascontiguousarray(x, np.float32)
    push = push or len(x)
    res = (O.OrcResult * max_res)()
    b = np.zeros((max_res, stride), np.uint8)
    cnt = np.zeros(8, np.int64)
    n = emu.emu_stream(O.CONFIGS[cfg], O.MODS[mod], rep, P(x, C.c_float), C.c_int64(len(x)), C.c_int64(block), C.c_int64(push),
                       res, max_res, P(b, C.c_uint8), C.c_int64(stride), P(cnt, C.c_int64))
    out = []
    for i in range(n):
        r = O.OrcResult()
        C.memmove(C.byref(r), C.byref(res[i]), C.sizeof(r))
        out.append((r, b[i, : max(0, r.n_bytes)].copy()))
    keys = ["framesDecoded", "frameErrors", "falsePositives", "receivedCount", "crcErrors", "duplicates", "metaReceived", "acScanPos"]
    return out, dict(zip(keys, [int(v) for v in cnt]))
