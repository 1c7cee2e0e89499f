"""Parity tests proper: the CUDA path (through the C-ABI library and its Python mirror) against the oracle on the
same seeded float32 inputs.  Bars (BASELINE.json north_star): preamble indices exact, decoded bytes bit-exact, CRC
verdicts identical, constellation points and correlation metrics within 1e-4 relative (the float64-faithful
device path is in fact held to 1e-9 or exact equality here)."""
import os

import numpy as np
import pytest

from helpers import MODES, make_stream, pad_blocks, results_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def modems(amb):
    cache = {}

    def get(mode_or_cfg, mod=None, rep=None):
        if mod is None:
            cfg, mod, rep, _ = MODES[mode_or_cfg]
        else:
            cfg = mode_or_cfg
        key = (cfg, mod, rep)
        if key not in cache:
            cache[key] = amb.Modem(cfg, mod, rep, device=0)
        return cache[key]

    yield get
    for m in cache.values():
        m.close()


def test_extension_is_loaded_and_launches_kernels(amb, modems):
    import os

    m = modems("QPSK")
    before = m.kernel_launches
    m.crc32([b"123456789"])
    assert m.kernel_launches > before
    maps = open(f"/proc/{os.getpid()}/maps").read()
    assert "libaudiomodem_b200.so" in maps


def test_scan_batch_selftest(modems):
    """warp-parallel candidate logic of the streaming scanner vs the reference-order loop, on the device (32 lanes)"""
    bad, commits = modems("QPSK").selftest(0, seed=7, n=200000)
    assert bad == 0 and commits > 10000, (bad, commits)


def test_crc32(O, modems):
    import zlib

    rng = np.random.default_rng(0)
    rows = [rng.integers(0, 256, n, dtype=np.uint8) for n in (0, 1, 9, 255, 4103, 32777)]
    rows[2] = np.frombuffer(b"123456789", np.uint8)
    got = modems("QPSK").crc32(rows)
    assert got[2] == 0xCBF43926
    assert [int(g) for g in got] == [zlib.crc32(r.tobytes()) for r in rows] == [O.crc32(r) for r in rows]


def test_preprocess_signal(O, modems):
    rng = np.random.default_rng(1)
    rows = [(rng.standard_normal(n) * 0.2 + 0.05).astype(np.float32) for n in (700, 4096, 100001, 392418)]
    rows.append(np.full(3000, 1e-8, np.float32))         # mx <= 1e-6: no division (modem.js:227)
    got = modems("QPSK").preprocess_signals(rows)
    for g, r in zip(got, rows):
        w = O.preprocessSignal(r)
        # the mean is a regrouped float64 sum here and a sequential one in the reference: a float32 rounding of x-mean
        # may flip for a vanishing fraction of samples (DESIGN.md numerics); in practice the arrays are identical
        neq = int((g != w).sum())
        assert neq <= max(1, len(r) // 100000), neq
        assert np.allclose(g, w, rtol=2e-7, atol=0)


@pytest.mark.parametrize("mode", list(MODES))
def test_detect_preamble(O, modems, mode):
    cfg, mod, rep, _ = MODES[mode]
    m = modems(mode)
    rows = []
    for i, sigma in enumerate((0.0, 0.002, 0.02, 0.08)):
        tx = O.buildTransmitSignal(cfg, bytes(O.randomBytes(20 + i, 150) | 0x11), mod, "f", rep)
        rows.append(O.preprocessSignal(O.addAwgn(tx, sigma, i) if sigma else tx))
    rows.append((np.random.default_rng(2).standard_normal(30000) * 0.05).astype(np.float32))    # no preamble
    rows.append(np.zeros(400, np.float32))                                                        # < 512 samples
    rows.append(np.zeros(5000, np.float32))                                                       # energy gate never opens
    idx, met = m.detect_preamble(rows)
    for i, r in enumerate(rows):
        wi, wm = O.detectPreamble(cfg, r, with_metric=True)
        assert idx[i] == wi, (mode, i, idx[i], wi, met[i], wm)
        assert abs(met[i] - wm) <= 1e-9 * max(1.0, wm)


def test_detect_preamble_tie_between_separate_offsets(O, modems):
    # narrowband + repeat-3 + zero bytes: a perfect junction peak ties with the preamble plateau to the last ulp;
    # the sequential float64 replay must settle it exactly as the reference's own recurrence does
    x = O.preprocessSignal(O.buildTransmitSignal("narrowband", bytes(300), "BPSK", "rec", 3))
    wi, wm = O.detectPreamble("narrowband", x, with_metric=True)
    idx, met = modems("BPSK-NARROW").detect_preamble([x])
    assert idx[0] == wi and met[0] == wm


def test_detect_preamble_xcorr(O, modems):
    tx = O.addAwgn(O.buildTransmitSignal("standard", bytes(O.randomBytes(3, 100)), "QPSK", "f", 1), 0.01, 1)
    rows = [tx, np.concatenate([np.zeros(51, np.float32), tx]), np.zeros(300, np.float32),
            (np.random.default_rng(3).standard_normal(20000) * 0.1).astype(np.float32)]
    idx, met = modems("QPSK").detect_preamble_xcorr(rows)
    for i, r in enumerate(rows):
        wi, wm = O.detectPreambleCrossCorr("standard", r, with_metric=True)
        assert idx[i] == wi and abs(met[i] - wm) <= 1e-12 * max(1.0, abs(wm)), (i, idx[i], wi, met[i], wm)
    tx = O.addAwgn(O.buildTransmitSignal("acoustic", bytes(O.randomBytes(3, 100)), "BPSK", "f", 1), 0.01, 1)
    x = np.concatenate([np.zeros(30, np.float32), tx])
    idx, _ = modems("BPSK-ACOUSTIC").detect_preamble_xcorr([x])
    assert idx[0] == O.detectPreambleCrossCorr("acoustic", x) == 22080


@pytest.mark.parametrize("mode", list(MODES))
def test_estimate_channel_and_demodulate(O, modems, mode):
    cfg, mod, rep, chunk = MODES[mode]
    m = modems(mode)
    c = O.config(cfg)
    sl = c["SYMBOL_LEN"]
    f = O.addAwgn(O.multipath(O.buildDataChunkFrame(cfg, O.randomBytes(5, chunk), 3, mod, rep), [0, 23, 57], [1.0, 0.5, 0.25]), 0.004, 8)[2205:]
    ce = O.generateChannelEstSymbol(cfg)
    want_h = O.estimateChannel(cfg, f[2 * sl:3 * sl], ce["knownRe"], ce["knownIm"])
    got_h = m.estimate_channel([f[2 * sl:3 * sl], f[2 * sl:3 * sl][:100]], ce["knownRe"], ce["knownIm"])
    assert np.array_equal(got_h[0][0], want_h[0]) and np.array_equal(got_h[1][0], want_h[1])      # float64-faithful: bit-identical
    short = O.estimateChannel(cfg, f[2 * sl:3 * sl][:100], ce["knownRe"], ce["knownIm"])
    assert np.array_equal(got_h[0][1], short[0]) and np.array_equal(got_h[1][1], short[1])
    # default known symbol == generateChannelEstSymbol
    got_d = m.estimate_channel([f[2 * sl:3 * sl]])
    assert np.array_equal(got_d[0][0], want_h[0])
    data = f[3 * sl:]
    rows = [data, data[: 5 * sl + 17], data[:sl - 1]]
    bits, eq = m.demodulate(rows, want_h[0], want_h[1], with_eq=True)
    for i, r in enumerate(rows):
        wb, we = O.demodulateOFDM(cfg, r, mod, want_h[0], want_h[1], with_eq=True)
        assert np.array_equal(bits[i], wb), (mode, i)
        assert eq[i].shape == we.shape
        assert np.allclose(eq[i], we, rtol=1e-9, atol=1e-12)
        assert np.array_equal(eq[i], we)


@pytest.mark.parametrize("mode", list(MODES))
def test_decode_chunk_frames(O, modems, mode):
    cfg, mod, rep, chunk = MODES[mode]
    m = modems(mode)
    lead = 13230 if cfg == "standard" else 22050
    sl = O.config(cfg)["SYMBOL_LEN"]
    frames = [
        O.addAwgn(O.buildDataChunkFrame(cfg, O.randomBytes(9, chunk), 5, mod, rep), 0.01, 5)[2205:],
        O.addAwgn(O.buildMetadataFrame(cfg, 7, 1234, chunk, "näme.bin", mod, rep), 0.005, 6)[lead:],
        O.addAwgn(O.buildMetadataFrame(cfg, 7, 1234, chunk, "x", mod, rep), 0.3, 6)[lead:],              # hopeless SNR
        O.buildDataChunkFrame(cfg, bytes(8), -2, mod, rep)[2205:],                                        # signed seqNum
        (np.random.default_rng(1).standard_normal(20000) * 0.1).astype(np.float32),                      # garbage
        np.zeros(5000, np.float32),
        np.zeros(100, np.float32),                                                                        # Frame too short for CE
        np.zeros(3 * sl, np.float32),                                                                     # No data after CE
        np.zeros(3 * sl + 10, np.float32),                                                                # Decoded data too short
        O.buildRawChunkFrame(cfg, bytes([0x42]) + bytes(40), mod, rep)[2205:],                            # Unknown frame type
        O.buildRawChunkFrame(cfg, bytes([0xFF, 0, 0, 0, 1, 0xFF, 0xFF]) + bytes(30), mod, rep)[2205:],    # Data chunk truncated
        O.buildRawChunkFrame(cfg, bytes([0xFE]) + bytes(10) + bytes([200]) + bytes(20), mod, rep)[2205:], # Metadata frame truncated
    ]
    got = m.decode_chunk_frames(frames, raw=True)
    objs = m.decode_chunk_frames(frames)
    for i, f in enumerate(frames):
        r1, b1 = O.decodeChunkFrame(cfg, f, mod, rep, raw=True)
        assert results_equal(r1, b1, got[i][0], got[i][1], check_metrics=False) == [], (mode, i)
        assert objs[i] == O.decodeChunkFrame(cfg, f, mod, rep), (mode, i)


@pytest.mark.parametrize("mode", list(MODES))
def test_fast_demodulator_carries_clean_frames(O, modems, mode):
    """The float32 symbol pipeline must decide clean frames itself (its float64 fallback is for symbols near a decision
    boundary): bit-exact results AND (almost) no fallback symbols at 25 dB; pure noise must still be bit-exact."""
    cfg, mod, rep, chunk = MODES[mode]
    m = modems(mode)
    frames = []
    for i in range(12):
        f = O.buildDataChunkFrame(cfg, O.randomBytes(300 + i, chunk), i, mod, rep)[2205:]
        p = O.power(f[: O.estimateFrameSamples(cfg, chunk + 11, mod, rep)])
        # windows up to 11 samples early (inside the cyclic prefix): a phase ramp the channel estimate absorbs
        frames.append(np.concatenate([np.zeros(i, np.float32), O.addAwgn(f, np.sqrt(p / 10 ** 2.5), 40 + i)]))
    before = m.debug_counter(1)
    got = m.decode_chunk_frames(frames, raw=True)
    fallbacks = m.debug_counter(1) - before
    n_sym = sum(len(f) // O.config(cfg)["SYMBOL_LEN"] - 3 for f in frames)
    for i, f in enumerate(frames):
        r1, b1 = O.decodeChunkFrame(cfg, f, mod, rep, raw=True)
        assert results_equal(r1, b1, got[i][0], got[i][1], check_metrics=False) == [], (mode, i)
    # the symbol after the frame (trailing silence + noise) is garbage and may fall back: one per frame
    assert fallbacks <= len(frames) + max(2, n_sym // 50), (fallbacks, n_sym)
    noise = [(np.random.default_rng(50 + i).standard_normal(30000) * 0.05).astype(np.float32) for i in range(4)]
    got = m.decode_chunk_frames(noise, raw=True)
    for i, f in enumerate(noise):
        r1, b1 = O.decodeChunkFrame(cfg, f, mod, rep, raw=True)
        assert results_equal(r1, b1, got[i][0], got[i][1], check_metrics=False) == [], (mode, i)


def test_decode_received_signal_c1(O, amb):
    """BASELINE config 1: QPSK loopback of a 32 KB payload (392,418 samples, noise-free)."""
    data = bytes(O.randomBytes(1, 32768))
    x = O.buildTransmitSignal("standard", data, "QPSK", "file", 1)
    amb.setOFDMConfig("standard")
    got = amb.decodeReceivedSignal(x, "QPSK", 1)
    want = O.decodeReceivedSignal("standard", x, "QPSK", 1)
    assert got == want
    assert got["preambleIdx"] == 13230 and got["crcValid"] and got["data"] == data and got["fileName"] == "file"


@pytest.mark.parametrize("mode", list(MODES))
def test_decode_received_signals_batch(O, modems, mode):
    """BASELINE config 3 in small: recordings through multipath + CFO + AWGN, ragged lengths, one batch."""
    cfg, mod, rep, _ = MODES[mode]
    m = modems(mode)
    recs = []
    rng = np.random.default_rng(4)
    for i in range(6):
        tx = O.buildTransmitSignal(cfg, bytes(O.randomBytes(1000 + i, 64 + 50 * i) | 0x11), mod, "rec%d" % i, rep)
        y = O.multipath(tx, [0, 23, 57], [1.0, 0.5, 0.25])
        df = (rng.uniform(-0.05, 0.05) if i != 3 else 1.0)
        y = O.cfo(y, df)
        p = O.power(tx[13230 if cfg == "standard" else 22050:][:5000])
        recs.append(O.addAwgn(y, np.sqrt(p / 100.0), 50 + i))      # 20 dB
    recs.append((rng.standard_normal(40000) * 0.05).astype(np.float32))                     # Preamble not detected
    recs.append(np.zeros(3000, np.float32))
    recs.append(recs[0][: (13230 if cfg == "standard" else 22050) + 2 * O.config(cfg)["SYMBOL_LEN"] + 10])   # Signal too short for CE
    got = m.decode_received_signals(recs, raw=True)
    objs = m.decode_received_signals(recs)
    for i, r in enumerate(recs):
        r1, b1 = O.decodeReceivedSignal(cfg, r, mod, rep, raw=True)
        d = results_equal(r1, b1, got[i][0], got[i][1], metric_tol=1e-9)
        assert d == [], (mode, i, d)
        assert objs[i] == O.decodeReceivedSignal(cfg, r, mod, rep), (mode, i)


@pytest.mark.parametrize("mode", ["BPSK-ACOUSTIC", "BPSK-REPEAT", "16-QAM"])
def test_decode_received_signals_stop_at_the_frames_extent(O, modems, mode):
    """Recordings much longer than the frame they hold (BASELINE config 3: 88,260 samples around a frame of ~30,000): the header
    symbols are demodulated first, the extent is read from the header, and only the symbols inside it are demodulated
    (k_demod_limit / k_extent).  Every result field and every byte the parsers consume equal the oracle's, n_bytes is the
    length of the whole decoded recording, and the bytes behind the extent are 0."""
    from helpers import frame_extent

    cfg, mod, rep, _ = MODES[mode]
    m = modems(mode)
    rng = np.random.default_rng(9)
    recs = []
    for i in range(5):
        tx = O.buildTransmitSignal(cfg, bytes(O.randomBytes(2000 + i, 40 + 90 * i)), mod, "r%d.bin" % i, rep)
        p = O.power(tx[13230 if cfg == "standard" else 22050:][:5000])
        tail = (rng.standard_normal(2 * len(tx) + 997 * i) * np.sqrt(p / 100.0)).astype(np.float32)      # noise at -20 dB behind the frame
        recs.append(np.concatenate([O.addAwgn(tx, np.sqrt(p / 100.0), 70 + i), tail]))
    got = m.decode_received_signals(recs, raw=True)
    for i, r in enumerate(recs):
        r1, b1 = O.decodeReceivedSignal(cfg, r, mod, rep, raw=True)
        r2, b2 = got[i]
        assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], (mode, i)
        assert r1.status == 0 and r2.n_bytes == r1.n_bytes == len(b1), (mode, i, r1.status, r2.n_bytes, len(b1))
        ext = frame_extent(r1)
        b2 = np.asarray(b2, np.uint8)
        assert len(b1) > 2 * ext                                           # (the recording really is much longer than its frame)
        # behind the extent, rounded up to whole symbols, nothing was demodulated
        sym_bytes = -(-O.config(cfg)["NUM_DATA_SUBS"] * {"BPSK": 1, "QPSK": 2, "QAM16": 4}[mod] // (8 * rep))       # bytes per OFDM symbol, rounded up
        assert not b2[ext + 2 * sym_bytes + 16:].any(), (mode, i, ext)


def _loopback_cases(O, cfg, mod, rep):
    """generateTestSignal (modem.js:913-975) clean and through channels, plus the early-return and fallback cases"""
    t = O.generateTestSignal(cfg, mod, rep)["signal"]
    lead = 13230 if cfg == "standard" else 22050
    sym = O.config(cfg)["SYMBOL_LEN"]
    p = O.power(t[lead:][:5000])
    cases = [t]
    cases.append(O.addAwgn(t, np.sqrt(p / 10 ** 2.0), 3))                                    # 20 dB
    cases.append(O.addAwgn(O.multipath(t, [0, 23, 57], [1.0, 0.5, 0.25]), np.sqrt(p / 10 ** 1.2), 4))
    cases.append(O.addAwgn(t, np.sqrt(p / 10 ** 0.3), 5))                                    # 3 dB: bit errors
    spike = t.copy(); spike[5] = 60.0                                                        # energy gate of detectPreamble fails -> cross-correlation fallback
    cases.append(spike)
    cases.append(t[: lead + 2 * sym + 10])                                                   # CE does not fit
    cases.append(t[: lead + 3 * sym])                                                        # channel, but no data symbols
    cases.append(t[: lead + 3 * sym + 5 * sym // 2])                                         # packet cut short
    cases.append(np.zeros(30000, np.float32))                                                # nothing detected
    cases.append((np.random.default_rng(1).standard_normal(20000) * 0.1).astype(np.float32))
    return cases


@pytest.mark.parametrize("mode", list(MODES))
def test_analyze_loopback(O, modems, mode):
    """analyzeLoopback (modem.js:978-1081): every field of the report against the oracle; float64 fields bit-equal
    (the device path repeats the reference's operation order; snrEstimate's log10 runs in the C library on both sides)."""
    cfg, mod, rep, _ = MODES[mode]
    cases = _loopback_cases(O, cfg, mod, rep)
    got = modems(mode).analyze_loopback(cases)
    seen = set()
    for i, (x, g) in enumerate(zip(cases, got)):
        w = O.analyzeLoopback(cfg, x, mod, rep)
        for k in ("detected", "quality", "startIdx", "ber"):
            assert g[k] == w[k] or (k == "ber" and np.isnan(g[k]) and np.isnan(w[k])), (mode, i, k, g[k], w[k])
        assert g["correlation"] == pytest.approx(w["correlation"], rel=1e-12, abs=1e-15), (mode, i)
        assert len(g["channelMagnitude"]) == len(w["channelMagnitude"]), (mode, i)
        assert np.array_equal(g["channelMagnitude"], w["channelMagnitude"]), (mode, i)
        assert g["snrEstimate"] == w["snrEstimate"], (mode, i, g["snrEstimate"], w["snrEstimate"])
        seen.add((w["detected"], len(w["channelMagnitude"]) > 0, w["quality"]))
    if cfg != "narrowband":
        assert got[0]["quality"] == "excellent" and got[0]["ber"] == 0.0
    assert (False, False, "poor") in seen and (True, False, "poor") in seen


def test_analyze_loopback_global_function(O, amb):
    amb.setOFDMConfig("standard")
    t = O.generateTestSignal("standard", "QPSK", 1)
    g = amb.analyzeLoopback(t["signal"], "QPSK", 1, t["testData"])
    w = O.analyzeLoopback("standard", t["signal"], "QPSK", 1)
    assert g["quality"] == w["quality"] == "excellent" and g["startIdx"] == 13230
    wrong = bytes(range(1, 17))
    assert amb.analyzeLoopback(t["signal"], "QPSK", 1, wrong)["ber"] == O.analyzeLoopback("standard", t["signal"], "QPSK", 1, wrong)["ber"] > 0


def _stream_compare(O, amb, mode, x, pushes=None, n_streams=1):
    cfg, mod, rep, _ = MODES[mode]
    xs = x if isinstance(x, list) else [x]
    n = len(xs[0])
    m = amb.Modem(cfg, mod, rep, device=0)
    pushes = pushes or [n]
    m.stream_open(len(xs), 4096, max(pushes))
    got = [[] for _ in xs]
    off = 0
    for p in pushes:
        blk = np.stack([xx[off:off + p] for xx in xs])
        for r, b in m.stream_push(blk, raw=True):
            got[r.stream].append((r, b))
        off += p
    assert off == n
    for s, xx in enumerate(xs):
        rx = O.StreamingReceiver(cfg, mod, rep)
        want = rx.push(xx, 4096, raw=True)
        g = sorted(got[s], key=lambda t: t[0].preamble_idx)
        assert len(g) == len(want), (mode, s, len(g), len(want))
        for (r1, b1), (r2, b2) in zip(want, g):
            d = results_equal(r1, b1, r2, b2, metric_tol=1e-9)
            assert d == [], (mode, s, r1.preamble_idx, d)
        c, cc = rx.counters(), m.stream_counters(s)
        for k in ("framesDecoded", "frameErrors", "falsePositives", "crcErrors", "metaReceived", "receivedCount", "duplicates"):
            assert c[k] == cc[k], (mode, s, k, c[k], cc[k])
        assert cc["samplesWritten"] == n
    m.stream_close()
    m.close()
    return got


@pytest.mark.parametrize("mode", list(MODES))
def test_stream_matches_oracle(O, amb, mode):
    x, _ = make_stream(O, mode, 3, 3)
    _stream_compare(O, amb, mode, x)
    n = len(x)
    pushes = [8 * 4096] * (n // (8 * 4096)) + ([n % (8 * 4096)] if n % (8 * 4096) else [])
    _stream_compare(O, amb, mode, x, pushes=pushes)


def test_stream_single_block_pushes(O, amb):
    """processAudioBlock cadence: one 4096-sample block per call, like the ScriptProcessor callback (app.js:1107)."""
    x, pay = make_stream(O, "16-QAM", 2, 21)
    got = _stream_compare(O, amb, "16-QAM", x, pushes=[4096] * (len(x) // 4096))
    assert [bytes(b[r.data_off:r.data_off + r.data_len]) for r, b in got[0] if r.frame_type == 0xFF] == pay


@pytest.mark.parametrize("snr_db,gaps", [(25.0, None), (25.0, [0, 2500, 40, 9000]), (7.0, [0, 333])])
def test_stream_phase_a_walks_stand_for_phase_b(O, amb, snr_db, gaps):
    """Dense streams after their metadata frame: phase A starts every run where it expects the receiver to be (the frame end
    predicted from the screening peak of the run before, or from the carried state) and phase B takes its walk over where the
    exact start state is that very state; everything else is walked again.  Events and counters equal the oracle's either way."""
    mode = "16-QAM"
    cfg, mod, rep, _ = MODES[mode]
    x, _ = make_stream(O, mode, 12, 31, snr_db=snr_db, gaps=gaps)
    n = len(x)
    first = 6 * 4096                                   # the metadata frame (16,416 samples) arrives in a push of its own
    pushes = [first, (n - first) // 4096 // 2 * 4096]
    pushes.append(n - sum(pushes))
    m = amb.Modem(cfg, mod, rep, device=0)
    m.stream_open(1, 4096, max(pushes))
    got, off = [], 0
    for p in pushes:
        got += list(m.stream_push(x[None, off:off + p], raw=True))
        off += p
    taken, walked = m.debug_counter(6), m.debug_counter(7)
    rx = O.StreamingReceiver(cfg, mod, rep)
    want = rx.push(x, 4096, raw=True)
    assert len(got) == len(want) >= 10, (len(got), len(want))
    for (r1, b1), (r2, b2) in zip(want, got):
        assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], r1.preamble_idx
    c, cc = rx.counters(), m.stream_counters(0)
    for k in ("framesDecoded", "frameErrors", "falsePositives", "crcErrors", "metaReceived", "receivedCount", "duplicates"):
        assert c[k] == cc[k], (k, c[k], cc[k])
    if snr_db >= 20:
        assert taken >= 6, (taken, walked)              # most links of the two dense pushes
        assert walked >= 1, (taken, walked)             # (the last link of a push at least: no frame completes in it)
    # (at 7 dB the metadata frame fails its CRC: the stream stays on the sequential pre-metadata path, results as the oracle's)
    m.stream_close()
    m.close()


@pytest.mark.parametrize("lead", [0, 777, 3000, 4000])
def test_stream_block_alignment(O, amb, lead):
    x, _ = make_stream(O, "16-QAM", 5, 11, lead=lead, gaps=[0, 100, 5000, 333])
    _stream_compare(O, amb, "16-QAM", x)


def test_stream_edge_cases(O, amb):
    x, _ = make_stream(O, "QPSK", 4, 12, snr_db=8.0)
    _stream_compare(O, amb, "QPSK", x)
    x, _ = make_stream(O, "16-QAM", 3, 14, noise=False)               # exact plateau ties
    _stream_compare(O, amb, "16-QAM", x)
    x, _ = make_stream(O, "16-QAM", 3, 15, meta=False, lead=5000)     # no metadata: 280-byte window for every frame
    _stream_compare(O, amb, "16-QAM", x)
    rng = np.random.default_rng(5)
    x, _ = make_stream(O, "QPSK", 2, 16)
    noise = (rng.standard_normal(60 * 4096) * 0.03).astype(np.float32)  # block-boundary defect accumulates over 60 idle calls
    _stream_compare(O, amb, "QPSK", np.concatenate([noise, x]))
    _stream_compare(O, amb, "QPSK", np.zeros(16 * 4096, np.float32))      # silence: nothing found
    dc = pad_blocks(make_stream(O, "16-QAM", 2, 17)[0] + np.float32(0.2))  # strong DC offset: the EMA matters
    _stream_compare(O, amb, "16-QAM", dc)


def test_stream_second_metadata_changes_chunk_size(O, amb):
    cfg, mod, rep, _ = MODES["QPSK"]
    parts = [O.buildMetadataFrame(cfg, 2, 4096, 2048, "a", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(i, 2048), i, mod, rep) for i in range(2)]
    parts.append(O.buildMetadataFrame(cfg, 3, 1536, 512, "b", mod, rep)[13230 - 2205:])
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(10 + i, 512), i, mod, rep) for i in range(3)]
    x = pad_blocks(O.addAwgn(np.concatenate(parts), 0.003, 3))
    _stream_compare(O, amb, "QPSK", x)


def test_chunk_bitmap_duplicates_and_reset(O, amb):
    """ChunkAssembler semantics without storage (app.js:610-643): duplicates, out-of-range chunks, bitmap reset by a new
    metadata frame; in one push (metadata inside the batch: ordered path) and in small pushes (parallel marking)"""
    cfg, mod, rep, chunk = MODES["QPSK"]
    seqs = [0, 1, 1, 2, 0, 7, 2]
    parts = [O.buildMetadataFrame(cfg, 4, 4 * chunk, chunk, "a.bin", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(50 + q, chunk), q, mod, rep) for q in seqs]
    parts.append(O.buildMetadataFrame(cfg, 4, 4 * chunk, chunk, "a.bin", mod, rep)[13230 - 2205:])
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(50 + q, chunk), q, mod, rep) for q in (1, 3, 3)]
    x = pad_blocks(O.addAwgn(np.concatenate(parts), 0.003, 5))
    _stream_compare(O, amb, "QPSK", x)
    n = len(x)
    _stream_compare(O, amb, "QPSK", x, pushes=[8 * 4096] * (n // (8 * 4096)) + ([n % (8 * 4096)] if n % (8 * 4096) else []))


def _assemble_from_events(events):
    """ChunkAssembler.handleMetadataFrame / handleDataChunk / assembleFile (app.js:610-682) over a list of result objects"""
    total = fsize = csize = 0
    seen, store = None, {}
    for r in events:
        if "error" in r:
            continue
        if r["frameType"] == 0xFE and r["crcValid"]:
            total, fsize, csize, seen, store = r["totalChunks"], r["totalFileSize"], r["chunkSize"], set(), {}
        elif r["frameType"] == 0xFF and seen is not None and not r["seqNum"] >= total and r["crcValid"] and r["seqNum"] not in seen:
            seen.add(r["seqNum"])
            store[r["seqNum"]] = r["data"]
    out = bytearray(max(fsize, 0))
    for i in range(max(total, 0)):
        if i in store:
            out[i * csize:i * csize + len(store[i])] = store[i]
    return bytes(out[:max(fsize, 0)])


def test_chunk_store_assembles_the_file(O, amb):
    """the storage-free ChunkAssembler on the device: chunks out of order, a duplicate, a missing chunk (zeros), a short last
    chunk; then a second file announced by a new metadata frame replaces the first"""
    cfg, mod, rep, chunk = MODES["QPSK"]
    data = [bytes(O.randomBytes(70 + i, chunk if i < 4 else 1000)) for i in range(5)]
    fsize = 4 * chunk + 1000
    parts = [O.buildMetadataFrame(cfg, 5, fsize, chunk, "f.bin", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, data[q], q, mod, rep) for q in (0, 2, 2, 4, 1)]
    x1 = pad_blocks(O.addAwgn(np.concatenate(parts), 0.003, 8))
    second = [bytes(O.randomBytes(90 + i, 512)) for i in range(3)]
    parts2 = [O.buildMetadataFrame(cfg, 3, 3 * 512, 512, "g.bin", mod, rep)[13230 - 2205:]]
    parts2 += [O.buildDataChunkFrame(cfg, second[q], q, mod, rep) for q in (2, 0, 1)]
    x2 = pad_blocks(O.addAwgn(np.concatenate(parts2), 0.003, 9))
    m = amb.Modem(cfg, mod, rep, device=0)
    try:
        m.stream_open(1, 4096, max(len(x1), len(x2)))
        m.stream_enable_assembly(1 << 16)
        assert m.stream_assemble(0) == b""
        m.stream_push(x1[None, :])
        rx = O.StreamingReceiver(cfg, mod, rep)
        ev = rx.push(x1, 4096)
        want = _assemble_from_events(ev)
        got = m.stream_assemble(0)
        assert len(got) == fsize and got == want
        assert got[:chunk] == data[0] and got[3 * chunk:4 * chunk] == bytes(chunk) and got[4 * chunk:] == data[4]
        n = len(x2)
        for p0 in range(0, n, 8 * 4096):                       # the second file in small pushes
            m.stream_push(x2[None, p0:p0 + 8 * 4096])
        ev += rx.push(x2, 4096)
        assert m.stream_assemble(0) == _assemble_from_events(ev)
        c, w = m.stream_counters(), rx.counters()
        assert c["receivedCount"] == w["receivedCount"] and c["duplicates"] == w["duplicates"]
    finally:
        m.close()


def test_many_streams_one_handle(O, amb):
    """BASELINE config 5 in small: independent streams (different payload, noise seed and lead-in) side by side."""
    xs, pays = [], []
    for s in range(12):
        x, pay = make_stream(O, "16-QAM", 3, 5000 + s, lead=137 * s)
        xs.append(x)
        pays.append(pay)
    n = max(len(x) for x in xs)
    xs = [np.concatenate([x, np.zeros(n - len(x), np.float32)]) for x in xs]
    got = _stream_compare(O, amb, "16-QAM", xs, pushes=[16 * 4096] * (n // (16 * 4096)) + ([n % (16 * 4096)] if n % (16 * 4096) else []))
    for s in range(12):
        data = [bytes(b[r.data_off:r.data_off + r.data_len]) for r, b in sorted(got[s], key=lambda t: t[0].preamble_idx) if r.frame_type == 0xFF]
        assert data == pays[s]


def test_round_trip_properties_at_scale(O, amb):
    """Size-independent properties on a longer stream than the oracle comparison covers comfortably: every frame
    found at its known position, every CRC valid, payload bytes equal to what was sent, counters consistent."""
    x, pay = make_stream(O, "16-QAM", 60, 31)
    m = amb.Modem("standard", "QAM16", 1)
    m.stream_open(1, 4096, len(x))
    ev = m.stream_push(x[None, :])
    assert len(ev) == 61 and all(e["crcValid"] for e in ev)
    assert [e["preambleIdx"] for e in ev] == [13230] + [16416 + 28431 * i + 2205 for i in range(60)]
    assert [e["data"] for e in ev[1:]] == pay and [e["seqNum"] for e in ev[1:]] == list(range(60))
    c = m.stream_counters()
    assert c["framesDecoded"] == 61 and c["frameErrors"] == 0 and c["receivedCount"] == 60 and c["crcErrors"] == 0
    m.stream_close()
    m.close()


def test_api_argument_errors(amb):
    m = amb.Modem("standard", "QPSK", 1)
    with pytest.raises(amb.AmbError):
        m.stream_push(np.zeros((1, 4096), np.float32))          # stream not open
    with pytest.raises(amb.AmbError):
        m.stream_open(1, 1000, 4000)                            # block_len must be 4096
    m.stream_open(1, 4096, 8192)
    with pytest.raises(amb.AmbError):
        m.stream_push(np.zeros((1, 5000), np.float32))          # not a multiple of block_len
    with pytest.raises(amb.AmbError):
        m.stream_push(np.zeros((1, 16384), np.float32))         # larger than max_push
    m.stream_close()
    m.close()
    with pytest.raises(amb.AmbError):
        amb.Modem("standard", "QAM64", 1)


@pytest.mark.parametrize("mode", list(MODES))
def test_tx_frames_bit_identical(O, modems, mode):
    """GPU transmit path (buildDataChunkFrame / buildMetadataFrame) against the oracle's frames, bit for bit."""
    cfg, mod, rep, chunk = MODES[mode]
    m = modems(mode)
    chunks = [O.randomBytes(60 + i, chunk) for i in range(5)]
    frames = m.build_data_chunk_frames(chunks, seq0=-2)
    for i, c in enumerate(chunks):
        assert np.array_equal(frames[i], O.buildDataChunkFrame(cfg, c, -2 + i, mod, rep)), (mode, i)
    short = m.build_data_chunk_frames([O.randomBytes(1, 17)], seq0=9)
    assert np.array_equal(short[0], O.buildDataChunkFrame(cfg, O.randomBytes(1, 17), 9, mod, rep))
    syn, data = m.build_data_chunk_frames(n=3, data_len=chunk, seq0=40, synth_seed=77, return_data=True)
    for i in range(3):
        assert np.array_equal(data[i], O.synthBytes(77, 40 + i, chunk))
        assert np.array_equal(syn[i], O.buildDataChunkFrame(cfg, data[i], 40 + i, mod, rep))
    meta = m.build_metadata_frame(9, 9 * chunk - 1, chunk, "ünï.bin")
    assert np.array_equal(meta, O.buildMetadataFrame(cfg, 9, 9 * chunk - 1, chunk, "ünï.bin", mod, rep))
    # noise: statistically the requested sigma, and the frames still decode
    noisy = m.build_data_chunk_frames(chunks, seq0=0, noise_sigma=0.001, noise_seed=5)
    d = noisy[0] - frames[0] * 0 - O.buildDataChunkFrame(cfg, chunks[0], 0, mod, rep)
    assert abs(float(d.std()) - 0.001) < 1e-4 and abs(float(d.mean())) < 1e-4
    r = m.decode_chunk_frames([noisy[0][2205:]])[0]
    assert r["crcValid"] and r["data"] == bytes(chunks[0])


@pytest.mark.parametrize("mode", list(MODES))
def test_tx_transmit_signal_bit_identical(O, amb, modems, mode):
    """buildTransmitSignal / generateTestSignal (modem.js:498-555, :913-975) on the device: bit-identical to the oracle's
    frames, and the loop closes through decodeReceivedSignal and analyzeLoopback"""
    cfg, mod, rep, _ = MODES[mode]
    m = modems(mode)
    datas = [bytes(O.randomBytes(300 + i, 200) | 0x01) for i in range(3)]
    sigs = m.build_transmit_signals(datas, "ünï-name.bin")
    for i, d in enumerate(datas):
        assert np.array_equal(sigs[i], O.buildTransmitSignal(cfg, d, mod, "ünï-name.bin", rep)), (mode, i)
    one = m.build_transmit_signals([b"\x07"], None)
    assert np.array_equal(one[0], O.buildTransmitSignal(cfg, b"\x07", mod, "file", rep))
    amb.setOFDMConfig(cfg)
    t = amb.generateTestSignal(mod, rep)
    assert np.array_equal(t["signal"], O.generateTestSignal(cfg, mod, rep)["signal"]) and bytes(t["testData"]) == bytes(range(16))
    full = amb.buildTransmitSignal(datas[0], mod, "a.bin", rep)
    assert full["dataLen"] == 200 and full["totalBits"] == (1 + 5 + 4 + 200 + 4) * 8 * rep
    assert full["numSymbols"] == (len(full["signal"]) - len(O.buildTransmitSignal(cfg, b"", mod, "a.bin", rep))) // O.config(cfg)["SYMBOL_LEN"] + \
        -(-((1 + 5 + 4 + 4) * 8 * rep) // full["bitsPerSymbol"])
    if cfg != "narrowband":
        r = amb.decodeReceivedSignal(full["signal"], mod, rep)
        assert r["crcValid"] and r["data"] == datas[0] and r["fileName"] == "a.bin"
        assert amb.analyzeLoopback(t["signal"], mod, rep, t["testData"])["quality"] == "excellent"
    amb.setOFDMConfig("standard")


# ---------------------------------------------------------------------------------------------- edge cases added in round 2
def _stream_events(O, amb, mode, x, pushes, max_chunk=4096, clip_ok=False):
    """events of the CUDA path and of the oracle for one stream; with clip_ok the decoded-byte rows may be shorter than the
    frame (the fields, CRC verdict and n_bytes must still be the oracle's, the bytes equal up to the row length)"""
    cfg, mod, rep, _ = MODES[mode]
    m = amb.Modem(cfg, mod, rep, device=0)
    try:
        m.stream_open(1, 4096, max(pushes), max_chunk=max_chunk)
        got, off = [], 0
        for p in pushes:
            got += m.stream_push(x[None, off:off + p], raw=True)
            off += p
        assert off == len(x)
        rx = O.StreamingReceiver(cfg, mod, rep)
        want = rx.push(x, 4096, raw=True)
        assert len(got) == len(want), (len(got), len(want))
        for (r1, b1), (r2, b2) in zip(want, got):
            if clip_ok and len(b2) < len(b1):
                assert len(b2) == m.bytes_stride and r2.n_bytes == r1.n_bytes
                b1 = b1[:len(b2)]
            assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], (mode, r1.preamble_idx, results_equal(r1, b1, r2, b2, metric_tol=1e-9))
        c, w = m.stream_counters(), rx.counters()
        for k in ("framesDecoded", "frameErrors", "falsePositives", "crcErrors", "metaReceived", "receivedCount", "duplicates"):
            assert c[k] == w[k], (k, c, w)
        return want
    finally:
        m.close()


def _pushes(n, per):
    return [per] * (n // per) + ([n % per] if n % per else [])


def test_stream_chunk_size_larger_than_the_rows(O, amb):
    """a metadata frame announcing chunkSize 8192 (rows sized for 4096): the collection window follows it (app.js:892-896); the
    results - fields, CRC verdict, n_bytes - do not depend on the caller's row length (k_parse used to parse the
    clipped row), and with rows sized for the chunk the bytes are the oracle's"""
    cfg, mod, rep, _ = MODES["QPSK"]
    big = 8192
    parts = [O.buildMetadataFrame(cfg, 2, 2 * big, big, "big.bin", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(70 + i, big), i, mod, rep) for i in range(2)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(80, 2048), 2, mod, rep)]          # a short chunk inside the long window
    x = pad_blocks(O.addAwgn(np.concatenate(parts), 0.003, 9), tail=4 * 4096)
    ev = _stream_events(O, amb, "QPSK", x, _pushes(len(x), 16 * 4096), clip_ok=True)
    assert sum(1 for r, _ in ev if r.status == 0 and r.frame_type == 0xFF and r.crc_valid and r.data_len == big) == 2
    _stream_events(O, amb, "QPSK", x, [len(x)], max_chunk=big)


def test_stream_ring_overwritten(O, amb):
    """RingBuffer.getRange -> null (app.js:581-582, :912-916): a metadata frame announcing a chunkSize whose collection window
    is longer than the ring (3 * estimateFrameSamples(4112) + 8192, app.js:712-715) makes every later frame a frameError
    without a message"""
    cfg, mod, rep, _ = MODES["QPSK"]
    huge = 20000
    parts = [O.buildMetadataFrame(cfg, 2, 2 * huge, huge, "huge.bin", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(90 + i, 2048), i, mod, rep) for i in range(6)]
    x = pad_blocks(O.addAwgn(np.concatenate(parts), 0.003, 10), tail=64 * 4096)
    ev = _stream_events(O, amb, "QPSK", x, _pushes(len(x), 8 * 4096), clip_ok=True)
    assert any(r.status == 14 for r, _ in ev), [r.status for r, _ in ev]
    _stream_events(O, amb, "QPSK", x, [len(x)], clip_ok=True)


def test_stream_nan_samples(O, amb):
    """a NaN sample poisons the reference's DC mean for good (app.js:750-755): frames before it decode, the frame it falls
    into is demodulated with `x || 0` (modem.js:377), nothing is found afterwards - also in later pushes"""
    x, _ = make_stream(O, "16-QAM", 4, 31)
    x = x.copy()
    bad = 16416 + 28431 + 9000               # inside the second data frame
    x[bad] = np.nan
    _stream_events(O, amb, "16-QAM", x, [len(x)])
    _stream_events(O, amb, "16-QAM", x, _pushes(len(x), 4 * 4096))
    y = x.copy()
    y[bad + 100000 // 2] = np.nan             # a second one, in a later span of the same push
    _stream_events(O, amb, "16-QAM", y, [len(y)])


def test_stream_slot_table_overflow(O, amb):
    """every tile of the push is hot (a 256-periodic tone: metric 1 everywhere), far more than the slot table holds
    (tiles / 8 + 64): the tiles without a slot are evaluated by the walker itself (amb_eval_windows), same events"""
    n = 96 * 4096
    t = np.arange(n)
    tone = (0.3 * np.sin(2 * np.pi * t * 5 / 256.0)).astype(np.float32)
    x, _ = make_stream(O, "QPSK", 2, 33)
    sig = np.concatenate([tone, x])
    _stream_events(O, amb, "QPSK", sig, [len(sig)])
    _stream_events(O, amb, "QPSK", sig, _pushes(len(sig), 32 * 4096))


def test_stream_near_ties_take_the_exact_replay(O, amb):
    """noise-free frames: the Schmidl-Cox plateau ties to the last ulp, and which sample wins is decided by the reference's own
    sliding sums (strict '>', app.js:811).  The call is replayed with them (amb_walk_scan_exact): coarse position and metric
    are the oracle's exactly, for every mode and for pushes of any length"""
    for mode in MODES:
        x, _ = make_stream(O, mode, 2, 35, noise=False, lead=1234)
        want = _stream_events(O, amb, mode, x, [len(x)])
        _stream_events(O, amb, mode, x, _pushes(len(x), 4096))
        cfg, mod, rep, _ = MODES[mode]
        m = amb.Modem(cfg, mod, rep, device=0)
        try:
            m.stream_open(1, 4096, len(x))
            got = m.stream_push(x[None, :], raw=True)
            exact = 0
            for (r1, _), (r2, _) in zip(want, got):
                assert abs(r1.coarse_metric - r2.coarse_metric) <= 1e-10 * r1.coarse_metric, (mode, r1.coarse_metric, r2.coarse_metric)
                exact += int(r1.coarse_metric == r2.coarse_metric)
            # a replayed call reports the reference's own value bit for bit (the standard and acoustic plateaus tie; a call without
            # a near-tie keeps the directly evaluated sums, 1e-13 away)
            print(mode, "events with a bit-equal coarse metric:", exact, "of", len(want))
        finally:
            m.close()


def test_c_abi_from_compiled_c(amb, tmp_path):
    """tests/cabi_smoke.c (strict C99, linked against the library) does create -> buildTransmitSignal -> decodeReceivedSignal ->
    preprocess + detectPreamble -> stream push -> destroy on the GPU and checks the decoded bytes itself"""
    import subprocess
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_abi import build_cabi_smoke

    exe = build_cabi_smoke(amb, tmp_path)
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "round trip on the GPU ok" in out.stdout
