"""The wider list of parity cases evaluated by the REAL reference (tools/qjs/dump_golden_qjs.py -> tests/golden/reference_cases.json)
and by the oracle (tests/test_golden.py), and - for the decode / stream kinds - by the CUDA path (tests/test_gpu_reference.py).

A case is a dict: name, mode, kind, args (JSON-able) and inputs (float32 / uint8 arrays, made here deterministically from
the oracle's transmit side and channel simulator: the reference has no noise source of its own, and its transmit side is
already pinned bit for bit by reference_golden.json).  `oracle_outputs(O, case)` produces what tools/qjs/cases_body.js
produces from the reference for the same case, in the same schema.
"""
import hashlib

import numpy as np

from helpers import MODES, make_stream, pad_blocks


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _c3_channel(O, sig, seed, snr_db=20.0, df=0.03, lead=0, tail=3000):
    """BASELINE config 3's channel: multipath 1.0@0 0.5@23 0.25@57, CFO, AWGN relative to the frame's power"""
    x = np.concatenate([np.zeros(lead, np.float32), sig, np.zeros(tail, np.float32)])
    x = O.multipath(x, [0, 23, 57], [1.0, 0.5, 0.25])
    x = O.cfo(x, df)
    p = O.power(x, lead + 2205, lead + len(sig) - 2205) if len(sig) > 6000 else O.power(x)
    return O.addAwgn(x, np.sqrt(p / 10 ** (snr_db / 10)), seed)


def build_cases(O, quick=False):
    cases = []

    def add(name, mode, kind, args=None, **inputs):
        cases.append({"name": name, "mode": mode, "kind": kind, "args": args or {}, "inputs": inputs})

    for mode, (cfg, mod, rep, chunk) in MODES.items():
        narrow = cfg == "narrowband"
        sl = O.config(cfg)["SYMBOL_LEN"]
        # ---- decodeReceivedSignal (modem.js:557-654): recordings through the C3 channel, ragged leads, failures
        for i, (n_bytes, snr, df, lead) in enumerate([(200, 20.0, 0.03, 0), (512, 20.0, -0.05, 1777), (64, 12.0, 1.0, 333), (37, 3.0, 0.0, 50)]):
            if narrow and i >= 2:
                continue
            sig = O.buildTransmitSignal(cfg, O.randomBytes(1000 + i, n_bytes), mod, "rec%d.bin" % i, rep)
            add(f"decode/c3_{i}", mode, "decode", x=_c3_channel(O, sig, 500 + i, snr, df, lead))
        sig = O.buildTransmitSignal(cfg, O.randomBytes(77, 48), mod, "t", rep)
        p0 = O.detectPreamble(cfg, O.preprocessSignal(sig))                 # the leading silence of buildTransmitSignal
        add("decode/truncated_in_ce", mode, "decode", x=sig[: p0 + 2 * sl + 100])          # Signal too short for CE
        add("decode/truncated_after_ce", mode, "decode", x=sig[: p0 + 3 * sl])              # No data after CE
        add("decode/truncated_payload", mode, "decode", x=sig[: p0 + 4 * sl + 5])            # a header and nothing behind it
        add("decode/noise_only", mode, "decode", x=(np.random.default_rng(5).standard_normal(40000) * 0.05).astype(np.float32))
        add("decode/silence", mode, "decode", x=np.zeros(9000, np.float32))
        add("decode/dc_offset", mode, "decode", x=(O.addAwgn(sig, 0.002, 3) + np.float32(0.21)).astype(np.float32))
        # ---- analyzeLoopback (modem.js:975-1081) on generateTestSignal's fixture through channels, its early returns and its
        # detector fallback (the cases of tests/test_gpu_parity.py::_loopback_cases)
        t = O.generateTestSignal(cfg, mod, rep)["signal"]
        lead_t = 13230 if cfg == "standard" else 22050
        pw = O.power(t[lead_t:][:5000])
        lb = {"clean": t, "awgn_20db": O.addAwgn(t, np.sqrt(pw / 10 ** 2.0), 3),
              "multipath_12db": O.addAwgn(O.multipath(t, [0, 23, 57], [1.0, 0.5, 0.25]), np.sqrt(pw / 10 ** 1.2), 4),
              "awgn_3db": O.addAwgn(t, np.sqrt(pw / 10 ** 0.3), 5), "ce_does_not_fit": t[: lead_t + 2 * sl + 10],
              "no_data_symbols": t[: lead_t + 3 * sl], "packet_cut_short": t[: lead_t + 3 * sl + 5 * sl // 2],
              "silence": np.zeros(30000, np.float32)}
        spike = t.copy()
        spike[5] = 60.0                      # the energy gate of detectPreamble fails -> detectPreambleCrossCorr fallback (modem.js:983-987)
        lb["spike_fallback"] = spike
        for k, v in lb.items():
            if narrow and k in ("multipath_12db", "awgn_3db", "spike_fallback"):
                continue
            add("loopback/" + k, mode, "loopback", x=np.ascontiguousarray(v, np.float32))
        # ---- decodeChunkFrame (modem.js:770-849): the edge frames of tests/test_gpu_parity.py::test_decode_chunk_frames
        lead = 2205 - 4
        frames = {
            "data_noisy": O.addAwgn(O.buildDataChunkFrame(cfg, O.randomBytes(9, chunk), 5, mod, rep), 0.01, 5)[2205:],
            "meta_utf8_name": O.addAwgn(O.buildMetadataFrame(cfg, 7, 1234, chunk, "näme.bin", mod, rep), 0.005, 6)[lead:],
            "meta_hopeless_snr": O.addAwgn(O.buildMetadataFrame(cfg, 7, 1234, chunk, "x", mod, rep), 0.3, 6)[lead:],
            "signed_seqnum": O.buildDataChunkFrame(cfg, bytes(8), -2, mod, rep)[2205:],
            "garbage": (np.random.default_rng(1).standard_normal(20000) * 0.1).astype(np.float32),
            "zeros": np.zeros(5000, np.float32),
            "too_short_for_ce": np.zeros(100, np.float32),
            "no_data_after_ce": np.zeros(3 * sl, np.float32),
            "decoded_too_short": np.zeros(3 * sl + 10, np.float32),
            "unknown_type": O.buildRawChunkFrame(cfg, bytes([0x42]) + bytes(40), mod, rep)[2205:],
            "data_truncated": O.buildRawChunkFrame(cfg, bytes([0xFF, 0, 0, 0, 1, 0xFF, 0xFF]) + bytes(30), mod, rep)[2205:],
            "meta_truncated": O.buildRawChunkFrame(cfg, bytes([0xFE]) + bytes(10) + bytes([200]) + bytes(20), mod, rep)[2205:],
        }
        for k, f in frames.items():
            if narrow and k in ("meta_hopeless_snr", "garbage"):
                continue
            add("chunkframe/" + k, mode, "chunkframe", x=f)
        # ---- the single operators on one noisy frame: estimateChannel (:421), demodulateOFDM (:365), modulateOFDM (:322)
        f = frames["data_noisy"]
        add("ops/channel_demod", mode, "ops", x=f[: (3 + (8 if narrow else 40)) * sl + 17])
        add("ops/modulate", mode, "modulate", bits=(O.randomBytes(11, 700) & 1).astype(np.uint8))
        # ---- StreamingReceiver (app.js:706-981)
        if not narrow:
            x, _ = make_stream(O, mode, 5, 41, lead=777, gaps=[0, 3000, 129, 7000])
            add("stream/gaps_and_lead", mode, "stream", x=x)
            x, _ = make_stream(O, mode, 2, 35, noise=False, lead=1234)
            add("stream/noise_free_plateau_ties", mode, "stream", x=x)
            x, _ = make_stream(O, mode, 3, 43, meta=False)
            add("stream/no_metadata", mode, "stream", x=x)
            x, _ = make_stream(O, mode, 3, 44, snr_db=14.0)
            add("stream/dc_offset_low_snr", mode, "stream", x=(x + np.float32(0.2)).astype(np.float32))
        else:
            x, _ = make_stream(O, mode, 1, 35, noise=False, lead=1234)
            add("stream/noise_free_plateau_ties", mode, "stream", x=x)
    # ---- stream edge cases (one mode each)
    cfg, mod, rep, _ = MODES["QPSK"]
    huge = 20000
    parts = [O.buildMetadataFrame(cfg, 2, 2 * huge, huge, "huge.bin", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(90 + i, 2048), i, mod, rep) for i in range(6)]
    add("stream/ring_overwritten", "QPSK", "stream", x=pad_blocks(O.addAwgn(np.concatenate(parts), 0.003, 10), tail=64 * 4096))
    big = 8192
    parts = [O.buildMetadataFrame(cfg, 2, 2 * big, big, "big.bin", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(70 + i, big), i, mod, rep) for i in range(2)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(80, 2048), 2, mod, rep)]
    add("stream/chunk_size_8192", "QPSK", "stream", x=pad_blocks(O.addAwgn(np.concatenate(parts), 0.003, 9), tail=4 * 4096))
    parts = [O.buildMetadataFrame(cfg, 4, 4 * 2048, 2048, "a.bin", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(60 + i, 2048), s, mod, rep) for i, s in enumerate([0, 1, 1, 3, 9, 0])]
    parts += [O.buildMetadataFrame(cfg, 2, 2 * 1024, 1024, "b.bin", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(50 + i, 1024), s, mod, rep) for i, s in enumerate([1, 1, 0])]
    add("stream/duplicates_and_second_metadata", "QPSK", "stream", x=pad_blocks(O.addAwgn(np.concatenate(parts), 0.003, 12)))
    x, _ = make_stream(O, "16-QAM", 4, 31)
    x = x.copy()
    x[16416 + 28431 + 9000] = np.nan
    add("stream/nan_sample", "16-QAM", "stream", x=x)
    x, _ = make_stream(O, "16-QAM", 4, 32)
    x = x.copy()
    x[16416 + 28431 + 5000: 16416 + 28431 + 5600] = 0.0            # a dropout inside a data frame: CRC error
    add("stream/crc_error", "16-QAM", "stream", x=x)
    n = 24 * 4096
    tone = (0.3 * np.sin(2 * np.pi * np.arange(n) * 5 / 256.0)).astype(np.float32)
    x, _ = make_stream(O, "QPSK", 2, 33)
    add("stream/periodic_tone_then_frames", "QPSK", "stream", x=np.concatenate([tone, x]))
    cfg, mod, rep, _ = MODES["BPSK-NARROW"]
    quiet = O.addAwgn(np.zeros(40 * 4096, np.float32), 10 ** (-30 / 20), 7)           # an idle stretch over a -30 dBFS floor (C4)
    f = O.buildDataChunkFrame(cfg, O.randomBytes(5, 512), 0, mod, rep)
    add("stream/c4_noise_floor_then_frame", "BPSK-NARROW", "stream",
        x=pad_blocks(np.concatenate([quiet, O.addAwgn(f, 10 ** (-30 / 20), 8)])))
    # ---- BASELINE config 1 at full size: buildTransmitSignal(32 KB, QPSK) -> decodeReceivedSignal (392,418 samples, noise-free)
    add("decode/c1_32kb_loopback", "QPSK", "decode", x=O.buildTransmitSignal("standard", O.randomBytes(1, 32768), "QPSK", "file", 1))
    # ---- integer / table functions
    add("misc/tables", "QPSK", "misc", bytes_a=O.randomBytes(3, 4107), bits_a=(O.randomBytes(4, 999) & 1).astype(np.uint8),
        fft_re=np.random.default_rng(7).standard_normal(512).astype(np.float32), fft_im=np.random.default_rng(8).standard_normal(512).astype(np.float32))
    if quick:
        cases = [c for c in cases if c["mode"] in ("QPSK", "BPSK-REPEAT") and c["kind"] != "stream"][:12]
    return cases


def _res(d):
    d = dict(d)
    if "data" in d:
        d["data_sha256"] = hashlib.sha256(bytes(d.pop("data"))).hexdigest()
    return d


def stream_outputs(events, counters, dc_mean):
    ev = []
    for r in events:
        ev.append(_res(r))
    keys = ("framesDecoded", "frameErrors", "receivedCount", "crcErrors", "duplicates", "metaReceived", "acScanPos")
    return {"events": ev, "counters": {k: int(counters[k]) for k in keys}, "dcMean": dc_mean}


def oracle_outputs(O, case):
    cfg, mod, rep, chunk = MODES[case["mode"]]
    kind, inp = case["kind"], case["inputs"]
    if kind == "decode":
        x = inp["x"]
        pre = O.preprocessSignal(x)
        return {"result": _res(O.decodeReceivedSignal(cfg, x, mod, rep)), "detectPreamble": O.detectPreamble(cfg, pre),
                "detectPreambleCrossCorr": O.detectPreambleCrossCorr(cfg, pre), "preprocess_sha256": sha(pre)}
    if kind == "chunkframe":
        return {"result": _res(O.decodeChunkFrame(cfg, inp["x"], mod, rep))}
    if kind == "ops":
        x = inp["x"]
        sl = O.config(cfg)["SYMBOL_LEN"]
        ce = O.generateChannelEstSymbol(cfg)
        h_re, h_im = O.estimateChannel(cfg, x[2 * sl: 3 * sl], ce["knownRe"], ce["knownIm"])
        bits, eq = O.demodulateOFDM(cfg, x[3 * sl:], mod, h_re, h_im, with_eq=True)
        eq = np.ascontiguousarray(eq, np.float64)
        return {"H_sha256": sha(np.concatenate([h_re, h_im])), "H_first": [float(h_re[O.config(cfg)["SUB_START"]]), float(h_im[O.config(cfg)["SUB_START"]])],
                "bits_sha256": sha(np.asarray(bits, np.uint8)), "n_bits": int(len(bits)),
                # the equalised, phase-corrected points handed to constellationDemap (modem.js:410-412), as (re, im) float64 pairs
                "eq_sha256": sha(eq), "eq_n": int(len(eq)), "eq_first": [float(v) for v in eq[:3].ravel()]}
    if kind == "loopback":
        r = O.analyzeLoopback(cfg, inp["x"], mod, rep)
        mags = np.asarray(r.pop("channelMagnitude"), np.float64)
        r.pop("startIdx", None)
        r.update(channelMagnitude_sha256=sha(mags), channelMagnitude_n=int(len(mags)))
        return r
    if kind == "modulate":
        bits = inp["bits"]
        sym = O.modulateOFDM(cfg, bits, mod)
        c = O.config(cfg)
        bps = {"BPSK": 1, "QPSK": 2, "QAM16": 4}[mod]
        per = c["NUM_DATA_SUBS"] * bps
        return {"numSymbols": int(sym.shape[0]), "bitsPerSymbol": per, "padded_len": int(-(-len(bits) // per) * per), "samples_sha256": sha(sym)}
    if kind == "stream":
        rx = O.StreamingReceiver(cfg, mod, rep)
        ev = rx.push(inp["x"], 4096)
        return stream_outputs(ev, rx.counters(), rx.dcMean())
    if kind == "misc":
        out = {"crc32": {str(n): O.crc32(inp["bytes_a"][:n]) for n in (0, 1, 7, 4107)}}
        re, im = O.fft(inp["fft_re"].astype(np.float64), inp["fft_im"].astype(np.float64))
        out["fft_sha256"] = sha(np.concatenate([re, im]))
        re, im = O.fft(inp["fft_re"].astype(np.float64), inp["fft_im"].astype(np.float64), inverse=True)
        out["ifft_sha256"] = sha(np.concatenate([re, im]))
        out["majorityVote3_sha256"] = sha(O.majorityVote(inp["bits_a"], 3))
        out["bitsToBytes_sha256"] = sha(O.bitsToBytes(inp["bits_a"]))
        out["estimateFrameSamples"] = {}
        out["configs"] = {}
        for c in ("standard", "acoustic", "narrowband"):
            k = O.config(c)
            out["configs"][c] = {"FFT_SIZE": k["FFT_SIZE"], "CP_LEN": k["CP_LEN"], "SYMBOL_LEN": k["SYMBOL_LEN"], "SUB_START": k["SUB_START"],
                                 "SUB_END": k["SUB_END"], "PILOTS": list(k["PILOTS"]), "numDataSubs": k["NUM_DATA_SUBS"],
                                 "pre2_sha256": sha(O.generatePreambleSymbol2(c))}
            for m in ("BPSK", "QPSK", "QAM16"):
                for r in (1, 3):
                    for p in (0, 1, 280, 4112):
                        out["estimateFrameSamples"][f"{c}/{m}/{r}/{p}"] = int(O.estimateFrameSamples(c, p, m, r))
        out["constellations"] = {m: [[float(a), float(b)] for a, b in O.constellationPoints(m)[1]] for m in ("BPSK", "QPSK", "QAM16")}
        st, _ = O.lcg(12345, 1000)
        out["lcg_12345_state_1000"] = int(st[-1])
        return out
    raise ValueError(kind)
