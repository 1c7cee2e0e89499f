"""The CUDA path against the REAL reference's outputs (no oracle in between): tests/golden/reference_cases.json holds what
playok/audio-modem's own modem.js / app.js returned for the cases of tests/refcases.py (tools/qjs/dump_golden_qjs.py); here the
same inputs go through the C-ABI library (via the Python mirror of the reference's interface) and every field must match:
result objects, error strings, detector indices, decoded bytes (by hash), stream events and counters exactly; H to 1e-12."""
import hashlib
import json
import os

import numpy as np
import pytest

import refcases
from helpers import MODES

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = json.load(open(os.path.join(HERE, "reference_cases.json")))


@pytest.fixture(scope="module")
def cases(O):
    return refcases.build_cases(O)         # (the oracle's transmit side + channel simulator only make the INPUTS; hashes checked below)


def _check_inputs(c):
    key = c["mode"] + ":" + c["name"]
    for k, a in c["inputs"].items():
        assert REF["inputs_sha256"][f"{key}:{k}"] == refcases.sha(a), ("input drifted", key, k)
    return REF["cases"][key]


def _obj(d):
    d = dict(d)
    d.pop("stream", None)
    if "data" in d:
        d["data_sha256"] = hashlib.sha256(bytes(d.pop("data"))).hexdigest()
    return d


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(MODES))
def test_batch_operators_equal_the_reference(amb, cases, mode):
    cfg, mod, rep, _ = MODES[mode]
    m = amb.Modem(cfg, mod, rep, device=0)
    n = 0
    try:
        mine = [c for c in cases if c["mode"] == mode]
        dec = [c for c in mine if c["kind"] == "decode"]
        got = m.decode_received_signals([c["inputs"]["x"] for c in dec])
        pre = m.preprocess_signals([c["inputs"]["x"] for c in dec])
        sc = m.detect_preamble(pre)
        xc = m.detect_preamble_xcorr(pre)
        for i, c in enumerate(dec):
            want = _check_inputs(c)
            assert _obj(got[i]) == want["result"], (mode, c["name"], got[i], want["result"])
            assert int(sc[0][i]) == want["detectPreamble"], (mode, c["name"])
            assert int(xc[0][i]) == want["detectPreambleCrossCorr"], (mode, c["name"])
            n += 3
        ck = [c for c in mine if c["kind"] == "chunkframe"]
        got = m.decode_chunk_frames([c["inputs"]["x"] for c in ck])
        for i, c in enumerate(ck):
            want = _check_inputs(c)
            assert _obj(got[i]) == want["result"], (mode, c["name"], got[i], want["result"])
            n += 1
        sl = m.SYMBOL_LEN
        for c in mine:
            if c["kind"] == "ops":
                want = _check_inputs(c)
                x = c["inputs"]["x"]
                h_re, h_im = m.estimate_channel([x[2 * sl: 3 * sl]])
                hh = np.concatenate([np.asarray(h_re[0], np.float64), np.asarray(h_im[0], np.float64)])
                assert hashlib.sha256(hh.tobytes()).hexdigest() == want["H_sha256"], (mode, "H is not bit-identical to the reference's")
                bits, eq = m.demodulate([x[3 * sl:]], h_re, h_im, with_eq=True)
                bits = np.asarray(bits[0], np.uint8)[: want["n_bits"]]
                assert hashlib.sha256(bits.tobytes()).hexdigest() == want["bits_sha256"], (mode, "demodulateOFDM bits")
                # the equalised points the reference hands to constellationDemap (north star: within 1e-4 relative; here bit-identical)
                eq = np.ascontiguousarray(np.asarray(eq[0], np.float64)[: want["eq_n"]])
                assert len(eq) == want["eq_n"]
                assert np.allclose(eq[:3].ravel(), want["eq_first"], rtol=1e-4, atol=0)
                assert hashlib.sha256(eq.tobytes()).hexdigest() == want["eq_sha256"], (mode, "equalised points are not bit-identical to the reference's")
                n += 3
            elif c["kind"] == "loopback":
                want = _check_inputs(c)
                lb = m.analyze_loopback([c["inputs"]["x"]])[0]
                mags = np.asarray(lb.pop("channelMagnitude"), np.float64)
                lb.pop("startIdx", None)
                assert (hashlib.sha256(mags.tobytes()).hexdigest(), len(mags)) == (want["channelMagnitude_sha256"], want["channelMagnitude_n"]), (mode, c["name"])
                for k, v in lb.items():
                    w = want[k]
                    assert (abs(v - w) <= 1e-12 * abs(w) if isinstance(w, float) else v == w), (mode, c["name"], k, v, w)
                n += 1
            elif c["kind"] == "modulate":
                want = _check_inputs(c)
                amb.setOFDMConfig(cfg)
                b = [int(v) for v in c["inputs"]["bits"]]
                r = amb.modulateOFDM(b, mod)
                sym = np.concatenate([np.asarray(s, np.float32) for s in r["samples"]])
                assert (r["numSymbols"], r["bitsPerSymbol"], len(b)) == (want["numSymbols"], want["bitsPerSymbol"], want["padded_len"])
                assert hashlib.sha256(sym.tobytes()).hexdigest() == want["samples_sha256"], (mode, "modulateOFDM samples")
                n += 4
    finally:
        m.close()
    assert n >= 20


@pytest.mark.gpu
def test_streams_equal_the_reference(amb, cases):
    """every stream case, pushed whole and in 8-block sections: events (fields, error strings, hashes of the data) and counters
    are the reference's"""
    n = 0
    for c in cases:
        if c["kind"] != "stream":
            continue
        want = _check_inputs(c)
        cfg, mod, rep, _ = MODES[c["mode"]]
        x = c["inputs"]["x"]
        x = x[: len(x) // 4096 * 4096]
        for per in (len(x), 8 * 4096):
            m = amb.Modem(cfg, mod, rep, device=0)
            try:
                m.stream_open(1, 4096, per, max_chunk=8192 if "8192" in c["name"] else 4096)
                ev = []
                for off in range(0, len(x), per):
                    ev += m.stream_push(x[None, off: off + per])
                ev = [_obj(e) for e in ev]
                assert len(ev) == len(want["events"]), (c["mode"], c["name"], per, len(ev), len(want["events"]))
                for g, w in zip(ev, want["events"]):
                    assert g == w, (c["mode"], c["name"], per, g, w)
                cnt = m.stream_counters()
                for k in ("framesDecoded", "frameErrors", "receivedCount", "crcErrors", "duplicates", "metaReceived"):
                    assert cnt[k] == want["counters"][k], (c["mode"], c["name"], k, cnt[k], want["counters"][k])
                n += 1
            finally:
                m.close()
    assert n >= 48


@pytest.mark.gpu
def test_tables_equal_the_reference(amb, cases):
    c = [c for c in cases if c["kind"] == "misc"][0]
    want = _check_inputs(c)
    m = amb.Modem("standard", "QPSK", 1, device=0)
    try:
        for n_, v in want["crc32"].items():
            assert int(m.crc32([c["inputs"]["bytes_a"][: int(n_)]])[0]) == v
    finally:
        m.close()
    for key, v in want["estimateFrameSamples"].items():
        cfg, mod, rep, p = key.split("/")
        amb.setOFDMConfig(cfg)
        assert amb.estimateFrameSamples(int(p), mod, int(rep)) == v, key
    amb.setOFDMConfig("standard")
