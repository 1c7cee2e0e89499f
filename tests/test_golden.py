"""Committed golden fixtures.  tests/golden/golden.json (tests/golden/make_golden.py, from the oracle): the oracle must still
reproduce it (drift pin), and the CUDA path must reproduce it through the C-ABI (GPU).  tests/golden/reference_golden.json
and reference_cases.json (tools/qjs/dump_golden_qjs.py, from the REAL modem.js / app.js): the oracle must equal them - the
pin against the reference itself."""
import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest

from helpers import MODES, make_stream

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = json.load(open(os.path.join(HERE, "golden.json")))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_reproduces_golden(O):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    now = json.loads(json.dumps(mg.make(), sort_keys=True))
    assert now == GOLD
    x = np.load(os.path.join(HERE, "test_frame_qpsk.npy"))
    assert _sha(x) == GOLD["modes"]["QPSK"]["test_signal"]["sha256"]


def test_reference_fixture_known_answers():
    # the reference's loopback packet: [4]'test'[0,0,0,16] 00..0F and its CRC
    assert GOLD["crc32"] == {"123456789": 0xCBF43926, "fixture_packet": 0x9D5BE7ED}
    assert GOLD["lcg"]["42"] == [1250496027, 1116302080, 1964818176, 1500480256]
    assert GOLD["lcg"]["43"] == [206527624, 86771232, 2052885472, 1531313664]
    assert GOLD["lcg"]["44"] == [1310042869, 1204723712, 1203392512, 783564800]
    for mode in ("QPSK", "16-QAM", "BPSK-ACOUSTIC", "BPSK-REPEAT"):
        d = GOLD["modes"][mode]["test_decode"]
        assert d["crcValid"] and d["data"] == bytes(range(16)).hex() and d["fileName"] == "test"
        assert d["expectedCRC"] == 0x9D5BE7ED
        lb = GOLD["modes"][mode]["loopback"]
        assert lb["detected"] and lb["quality"] == "excellent" and lb["ber"] == 0.0 and lb["correlation"] > 0.999
        assert lb["startIdx"] == (13230 if mode in ("QPSK", "16-QAM") else 22050)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(MODES))
def test_gpu_reproduces_golden(O, amb, mode):
    cfg, mod, rep, chunk = MODES[mode]
    e = GOLD["modes"][mode]
    m = amb.Modem(cfg, mod, rep, device=0)
    try:
        # transmit side: bit-identical frames
        assert _sha(m.build_metadata_frame(3, 3 * chunk, chunk, "golden.bin")) == e["meta_frame"]["sha256"]
        assert _sha(m.build_data_chunk_frames([O.randomBytes(4242, chunk)], seq0=2)[0]) == e["data_frame"]["sha256"]
        # the reference's loopback fixture through decodeReceivedSignal
        sig = O.buildTransmitSignal(cfg, bytes(range(16)), mod, "test", rep)
        assert _sha(sig) == e["test_signal"]["sha256"]
        got = m.decode_received_signals([sig])[0]
        if "data" in got:
            got["data"] = got["data"].hex()
        assert got == e["test_decode"]
        assert int(m.detect_preamble(m.preprocess_signals([sig]))[0][0]) == e["detectPreamble"]
        lb = m.analyze_loopback([sig])[0]
        mags = lb.pop("channelMagnitude")
        assert _sha(mags) == e["loopback"]["channelMagnitude_sha256"] and len(mags) == e["loopback"]["channelMagnitude_n"]
        for k, v in lb.items():
            assert e["loopback"][k] == (pytest.approx(v, rel=1e-12) if k == "correlation" else v), (mode, k, v, e["loopback"][k])
        # streaming events
        x, _ = make_stream(O, mode, 3, 21)
        assert _sha(x) == e["stream"]["input"]["sha256"]
        m.stream_open(1, 4096, len(x))
        ev = m.stream_push(x[None, :])
        m.stream_close()
        assert len(ev) == len(e["stream"]["events"])
        for g, w in zip(ev, e["stream"]["events"]):
            g.pop("stream")
            data = g.pop("data", None)
            if data is not None:
                assert hashlib.sha256(data).hexdigest() == w["data_sha256"]
            for k, v in g.items():
                assert w[k] == v, (mode, k)
    finally:
        m.close()


REF = os.path.join(HERE, "reference_golden.json")


def _close(a, b, rel=1e-9):
    if a is None or b is None:                   # JSON.stringify writes NaN as null
        return (a is None or a != a) and (b is None or b != b)
    if isinstance(a, float) or isinstance(b, float):
        return a == b or abs(a - b) <= rel * max(abs(a), abs(b))
    return a == b


def _diff(ref, orc, path=""):
    """differences between the reference's dump and the oracle's fixtures: keys the reference does not have (our extras, e.g.
    loopback.startIdx) are ignored, floats compare to 1e-9 relative, everything else exactly"""
    out = []
    if isinstance(ref, dict) and isinstance(orc, dict):
        for k, v in ref.items():
            if k in ("note", "counters"):
                continue
            if k not in orc:
                out.append((path + "/" + k, "missing in the oracle's fixtures"))
            else:
                out += _diff(v, orc[k], path + "/" + k)
    elif isinstance(ref, list) and isinstance(orc, list):
        if len(ref) != len(orc):
            out.append((path, f"length {len(ref)} != {len(orc)}"))
        for i, (a, b) in enumerate(zip(ref, orc)):
            out += _diff(a, b, f"{path}[{i}]")
    elif not _close(ref, orc):
        out.append((path, ref, orc))
    return out


def test_reference_golden_matches_the_oracle():
    """THE pin against modem.js: tests/golden/reference_golden.json was written by the REAL reference - modem.js and the
    RingBuffer / StreamingReceiver classes of app.js, evaluated by Qt's QJSEngine (tools/qjs/dump_golden_qjs.py; the same
    dump runs under Node with node/dump_golden.js).  Every value in it must equal the oracle's fixture: transmit frames by
    sha-256 of their Float32 samples, decodes, detector indices, the loopback analysis, frame lengths, CRCs, LCG states and
    the streaming events of the five modes."""
    assert os.path.exists(REF), "tests/golden/reference_golden.json is missing: parity would be unpinned (DESIGN.md section 2)"
    ref = json.load(open(REF))
    assert "evaluated by" in ref["note"] or "from the reference" in ref["note"]
    d = _diff(ref, GOLD)
    assert d == [], d[:20]
    n_exact, n = _count_exact(ref, GOLD)
    assert n >= 450 and n_exact == n, (n_exact, n)         # floats included: the oracle reproduces the reference bit for bit


def _count_exact(ref, orc):
    """(leaves that are bit-equal, leaves compared); NaN equals null (JSON.stringify writes NaN as null)"""
    if isinstance(ref, dict):
        a = [_count_exact(v, orc[k]) for k, v in ref.items() if k not in ("note",) and isinstance(orc, dict) and k in orc]
        return sum(x for x, _ in a), sum(y for _, y in a)
    if isinstance(ref, list):
        a = [_count_exact(x, y) for x, y in zip(ref, orc)]
        return sum(x for x, _ in a), sum(y for _, y in a)
    same = ref == orc or ((ref is None or ref != ref) and (orc is None or orc != orc))
    return int(same), 1


REF_CASES = os.path.join(HERE, "reference_cases.json")


def test_reference_cases_match_the_oracle(O):
    """the wider pin: the cases of tests/refcases.py - recordings through multipath + CFO + AWGN, truncated and empty
    recordings, BASELINE config 1 at full size, every decodeChunkFrame error string, analyzeLoopback through channels and its
    early returns, the single operators (estimateChannel, demodulateOFDM with the equalised points it hands to the demapper,
    modulateOFDM, fft, crc32, majorityVote, ...), and 24 streams (gaps, plateau ties, no metadata, DC offset, ring overwritten, chunkSize 8192,
    duplicates + second metadata frame, a NaN sample, a CRC error, a periodic tone, the C4 noise floor) - were run through the
    REAL reference (tests/golden/reference_cases.json, written by tools/qjs/dump_golden_qjs.py).  The oracle must reproduce
    every value on the same inputs (inputs are re-made here and checked by hash)."""
    import refcases

    ref = json.load(open(REF_CASES))
    cases = refcases.build_cases(O)
    assert len(cases) == len(ref["cases"]) >= 180
    n_exact = n = 0
    for c in cases:
        key = c["mode"] + ":" + c["name"]
        for k, a in c["inputs"].items():
            assert ref["inputs_sha256"][f"{key}:{k}"] == refcases.sha(a), ("input drifted", key, k)
        got = json.loads(json.dumps(refcases.oracle_outputs(O, c)))
        d = _diff(ref["cases"][key], got)
        assert d == [], (key, d[:6])
        e, t = _count_exact(ref["cases"][key], got)
        n_exact, n = n_exact + e, n + t
    assert n >= 1700 and n_exact == n, (n_exact, n)


def test_reference_golden_comparison_logic():
    """the comparison itself, exercised on the oracle's own fixtures (so that a future reference_golden.json is judged by
    tested code): identical -> no difference; a changed hash, a changed float and a dropped event are all reported"""
    import copy

    assert _diff(GOLD, GOLD) == []
    g = copy.deepcopy(GOLD)
    g["modes"]["QPSK"]["test_signal"]["sha256"] = "0" * 64
    g["modes"]["QPSK"]["loopback"]["correlation"] *= 1.0 + 1e-6
    g["modes"]["16-QAM"]["stream"]["events"].pop()
    del g["modes"]["QPSK"]["loopback"]["startIdx"]          # a key the reference does not produce: ignored
    paths = [x[0] for x in _diff(g, GOLD)]
    assert "/modes/QPSK/test_signal/sha256" in paths and "/modes/QPSK/loopback/correlation" in paths and "/modes/16-QAM/stream/events" in paths
    assert len(paths) == 3
