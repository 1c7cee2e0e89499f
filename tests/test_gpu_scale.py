"""Parity at larger sizes, shaped like BASELINE.json's configs 2-5 (reduced so the oracle finishes in seconds), plus
size-independent properties on what the oracle does not replay.  GPU only."""
import numpy as np
import pytest

from helpers import MODES, make_stream, pad_blocks, results_equal

pytestmark = pytest.mark.gpu


def _events_equal(O, amb, mode, x, pushes):
    cfg, mod, rep, _ = MODES[mode]
    m = amb.Modem(cfg, mod, rep, device=0)
    try:
        m.stream_open(1, 4096, max(pushes))
        got, off = [], 0
        for p in pushes:
            got += m.stream_push(x[None, off:off + p], raw=True)
            off += p
        assert off == len(x)
        rx = O.StreamingReceiver(cfg, mod, rep)
        want = rx.push(x, 4096, raw=True)
        assert len(got) == len(want), (len(got), len(want))
        for (r1, b1), (r2, b2) in zip(want, got):
            assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], (mode, r1.preamble_idx)
        c, w = m.stream_counters(), rx.counters()
        for k in ("framesDecoded", "frameErrors", "falsePositives", "crcErrors", "metaReceived", "receivedCount", "duplicates"):
            assert c[k] == w[k], (k, c, w)
        return want
    finally:
        m.close()


def test_c2_chunk_stream_400_frames(O, amb):
    """config 2 in small: 16-QAM, 4 KB chunks, AWGN 25 dB, 400 frames, three unequal pushes - every event equal to the oracle's"""
    x, pay = make_stream(O, "16-QAM", 400, 77)
    n = len(x)
    a, b = (n // 3) // 4096 * 4096, (n // 2) // 4096 * 4096
    ev = _events_equal(O, amb, "16-QAM", x, [a, b, n - a - b])
    data = [bytes(bb[r.data_off:r.data_off + r.data_len]) for r, bb in ev if r.status == 0 and r.frame_type == 0xFF and r.crc_valid]
    assert len(data) >= 398 and set(data) <= set(pay)


def test_c4_narrowband_sparse_capture(O, amb):
    """config 4 in small: narrowband frames far apart in a long noisy capture (scan dominated)"""
    cfg, mod, rep, chunk = MODES["BPSK-NARROW"]
    rng = np.random.default_rng(9)
    parts = [np.zeros(5 * 4410, np.float32), O.buildMetadataFrame(cfg, 3, 3 * chunk, chunk, "cap.bin", mod, rep)]
    for i, gap in enumerate((300000, 441000, 250000)):
        parts.append(np.zeros(gap, np.float32))
        parts.append(O.buildDataChunkFrame(cfg, O.randomBytes(600 + i, chunk), i, mod, rep))
    x = np.concatenate(parts)
    x = pad_blocks((x + rng.standard_normal(len(x)).astype(np.float32) * np.float32(10 ** (-30 / 20))).astype(np.float32), tail=3 * 4096)
    n = len(x)
    pushes = [64 * 4096] * (n // (64 * 4096)) + ([n % (64 * 4096)] if n % (64 * 4096) else [])
    ev = _events_equal(O, amb, "BPSK-NARROW", x, pushes)
    assert sum(1 for r, _ in ev if r.status == 0 and r.crc_valid) >= 3


@pytest.mark.parametrize("mode", ["BPSK-ACOUSTIC", "BPSK-REPEAT"])
def test_c3_recordings_batch(O, amb, mode):
    """config 3 in small: 48 recordings through multipath, CFO (10 % at +-1 Hz: the reference has no CFO correction, the
    failures must be identical) and AWGN 20 dB, decoded as one batch"""
    cfg, mod, rep, _ = MODES[mode]
    rng = np.random.default_rng(11)
    recs = []
    for i in range(48):
        tx = O.buildTransmitSignal(cfg, O.randomBytes(1000 + i, 512), mod, "rec", rep)
        y = O.multipath(tx, [0, 23, 57], [1.0, 0.5, 0.25])
        y = O.cfo(y, rng.choice([-1.0, 1.0]) if i % 10 == 3 else rng.uniform(-0.05, 0.05))
        recs.append(O.addAwgn(y, float(np.sqrt(O.power(tx[22050:][:20000]) / 100.0)), 1000 + i))
    m = amb.Modem(cfg, mod, rep, device=0)
    try:
        got = m.decode_received_signals(recs, raw=True)
    finally:
        m.close()
    ok = 0
    for i, r in enumerate(recs):
        r1, b1 = O.decodeReceivedSignal(cfg, r, mod, rep, raw=True)
        assert results_equal(r1, b1, got[i][0], got[i][1], metric_tol=1e-9) == [], (mode, i)
        ok += int(r1.status == 0 and r1.crc_valid)
    assert ok >= 30


def test_c5_many_streams(O, amb):
    """config 5 in small: 192 independent 16-QAM streams built on the device (metadata + 3 chunks, own payload and noise),
    pushed in sections; eight of them replayed by the oracle, all of them checked through the counters"""
    import torch

    cfg, mod, rep, chunk = MODES["16-QAM"]
    ns, K = 192, 3
    dev = torch.device("cuda", 0)
    m = amb.Modem(cfg, mod, rep, device=0)
    try:
        meta_len = m.chunk_frame_samples(1 + 4 + 4 + 2 + 1 + len("s.bin") + 4, True)
        flen = m.chunk_frame_samples(chunk + 11, False)
        push = 8 * 4096
        n = (meta_len + K * flen + 8192 + push - 1) // push * push
        x = torch.zeros((ns, n), dtype=torch.float32, device=dev)
        clean = torch.zeros((1, flen), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()          # the library writes on its own stream: torch's fills must be done first
        m.build_data_chunk_frames(n=1, data_len=chunk, seq0=0, synth_seed=2, out=clean)
        m.synchronize()
        sigma = float((clean[0, 2205:flen - 882].double().pow(2).mean() / 10 ** 2.5).sqrt())       # 25 dB below the frame core
        for s in range(ns):
            m.build_metadata_frame(K, K * chunk, chunk, "s.bin", noise_sigma=sigma, noise_seed=5000 + s, noise_g0=0, out=x[s, :meta_len])
            m.build_data_chunk_frames(n=K, data_len=chunk, seq0=0, synth_seed=9000 + s, noise_sigma=sigma, noise_seed=5000 + s, noise_g0=meta_len,
                                      out=x[s, meta_len:meta_len + K * flen].view(K, flen))
        m.synchronize()
        m.stream_open(ns, 4096, push)
        got = [[] for _ in range(ns)]
        for p0 in range(0, n, push):
            for r, b in m.stream_push(x[:, p0:p0 + push], raw=True):
                got[r.stream].append((r, b))
        c = m.stream_counters()
        assert c["metaReceived"] == ns and c["framesDecoded"] == ns * (K + 1) and c["frameErrors"] == 0
        assert c["receivedCount"] + c["crcErrors"] == ns * K
        host = x[:8].cpu().numpy()
        for s in range(8):
            want = O.StreamingReceiver(cfg, mod, rep).push(host[s], 4096, raw=True)
            g = sorted(got[s], key=lambda t: t[0].preamble_idx)
            assert len(g) == len(want) == K + 1
            for (r1, b1), (r2, b2) in zip(want, g):
                assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], (s, r1.preamble_idx)
        m.stream_close()
    finally:
        m.close()


def test_sectioned_host_push_equals_one_push(O, amb):
    """large host pushes are cut into sections whose copies overlap the processing (amb_stream_push and
    amb_stream_push_pcm16): same events as the oracle, also when the results of several sections land in the caller's
    buffers back to back"""
    import os
    import subprocess
    import sys

    # the threshold is read once per process: run the comparison in a child with a small one
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "import audio_modem_b200 as amb\n"
        "from oracle import oracle as O\n"
        "from helpers import make_stream, results_equal\n"
        "x, pay = make_stream(O, '16-QAM', 40, 123)\n"
        "m = amb.Modem('standard', 'QAM16', 1, device=0)\n"
        "m.stream_open(1, 4096, len(x))\n"
        "got = m.stream_push(x[None, :], raw=True)\n"
        "want = O.StreamingReceiver('standard', 'QAM16', 1).push(x, 4096, raw=True)\n"
        "assert len(got) == len(want) == 41, (len(got), len(want))\n"
        "for (r1, b1), (r2, b2) in zip(want, got):\n"
        "    assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], r1.preamble_idx\n"
        "pcm = np.clip(np.rint(x.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)\n"
        "m.stream_close(); m.stream_open(1, 4096, len(x))\n"
        "got16 = m.stream_push(pcm[None, :], raw=True)\n"
        "want16 = O.StreamingReceiver('standard', 'QAM16', 1).push(pcm.astype(np.float32) / np.float32(32768.0), 4096, raw=True)\n"
        "assert len(got16) == len(want16) == 41, (len(got16), len(want16))\n"
        "for (r1, b1), (r2, b2) in zip(want16, got16):\n"
        "    assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], r1.preamble_idx\n"
        "print('sections ok')\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AMB_SECTION_MIN_BYTES="65536")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "sections ok" in out.stdout, out.stdout + out.stderr


def _quantise(x):
    """float samples -> 16-bit PCM and the float32 array an audio front end makes of it (s / 32768, exact)"""
    pcm = np.clip(np.rint(np.asarray(x, np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    return pcm, (pcm.astype(np.float32) / np.float32(32768.0))


@pytest.mark.parametrize("mode", ["16-QAM", "BPSK-REPEAT"])
def test_pcm16_push_equals_float_push(O, amb, mode):
    """wire-format ingest: 16-bit PCM pushed from host memory, from device memory and in several pushes gives
    the oracle's events for the float32 samples s / 32768"""
    import torch

    cfg, mod, rep, _ = MODES[mode]
    xs = []
    for s in range(3):
        x, _ = make_stream(O, mode, 6, 40 + s, lead=1000 * s + 17)
        xs.append(x)
    n = max(len(x) for x in xs)
    pcm = np.zeros((3, n), np.int16)
    flt = np.zeros((3, n), np.float32)
    for s, x in enumerate(xs):
        pcm[s, : len(x)], flt[s, : len(x)] = _quantise(x)
    want = [O.StreamingReceiver(cfg, mod, rep).push(flt[s], 4096, raw=True) for s in range(3)]
    assert all(len(w) >= 5 for w in want)
    blocks = n // 4096
    splits = [blocks // 3 * 4096, (blocks - blocks // 3) * 4096]
    for where in ("host", "device", "device_unaligned"):
        m = amb.Modem(cfg, mod, rep, device=0)
        try:
            m.stream_open(3, 4096, max(splits))
            got = [[] for _ in range(3)]
            off = 0
            for p in splits:
                blk = np.ascontiguousarray(pcm[:, off:off + p])
                if where == "device":
                    blk = torch.from_numpy(blk).cuda()
                elif where == "device_unaligned":
                    t = torch.zeros((3, p + 3), dtype=torch.int16, device="cuda")
                    t[:, 3:] = torch.from_numpy(blk).cuda()
                    blk = t[:, 3:]
                for r, b in m.stream_push(blk, raw=True):
                    got[r.stream].append((r, b))
                off += p
            for s in range(3):
                assert len(got[s]) == len(want[s]), (where, s, len(got[s]))
                for (r1, b1), (r2, b2) in zip(want[s], got[s]):
                    assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], (where, s, r1.preamble_idx)
            assert m.stream_counters()["samplesWritten"] == 3 * n
        finally:
            m.close()


# ---------------------------------------------------------------------------------------------- SURVEY.md section 8(d) sizes
def _pool_map(fn, items, threads=8):
    """the oracle's C++ runs under ctypes with the GIL released: a thread pool keeps its replays to seconds"""
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(fn, items))


def test_c2_chunk_stream_2000_frames(O, amb):
    """config 2 at the parity size of SURVEY 8(d): the first 2,000 frames of the 16-QAM stream (57 M samples, AWGN 25 dB), one
    push of 1,000 frames and pushes of 128 blocks - every event, every decoded byte and every counter equal to the oracle's"""
    cfg, mod, rep, chunk = MODES["16-QAM"]
    n_frames = 2000
    parts = [O.buildMetadataFrame(cfg, n_frames, n_frames * chunk, chunk, "c2.bin", mod, rep)]
    frames = _pool_map(lambda i: O.buildDataChunkFrame(cfg, O.synthBytes(2, i, chunk), i, mod, rep), range(n_frames))
    x = np.concatenate(parts + frames)
    p = O.power(frames[0][2205:2205 + 25344])
    x = pad_blocks(O.addAwgnCtr(x, (p / 10 ** 2.5) ** 0.5, 3))
    n = len(x)
    half = (n // 2) // 4096 * 4096
    pushes = [half] + [128 * 4096] * ((n - half) // (128 * 4096)) + ([(n - half) % (128 * 4096)] if (n - half) % (128 * 4096) else [])
    ev = _events_equal(O, amb, "16-QAM", x, pushes)
    ok = [r for r, _ in ev if r.status == 0 and r.frame_type == 0xFF and r.crc_valid]
    assert len(ev) == n_frames + 1 and len(ok) >= n_frames - 8
    assert sorted(r.seq_num for r in ok) == sorted(set(r.seq_num for r in ok))


@pytest.mark.parametrize("mode", ["BPSK-ACOUSTIC", "BPSK-REPEAT"])
def test_c3_recordings_batch_1000(O, amb, mode):
    """config 3 at a tenth of its size: 500 recordings per mode through multipath, CFO (10 % at +-1 Hz) and AWGN 20 dB, one batch;
    preambleIdx, verdict, every result field and the decoded bytes of every recording equal to the oracle's"""
    cfg, mod, rep, _ = MODES[mode]
    rng = np.random.default_rng(12)
    dfs = [float(rng.choice([-1.0, 1.0])) if i % 10 == 3 else float(rng.uniform(-0.05, 0.05)) for i in range(500)]

    def make(i):
        tx = O.buildTransmitSignal(cfg, O.randomBytes(3000 + i, 512), mod, "rec", rep)
        y = O.cfo(O.multipath(tx, [0, 23, 57], [1.0, 0.5, 0.25]), dfs[i])
        return O.addAwgn(y, float(np.sqrt(O.power(tx[22050:][:20000]) / 100.0)), 3000 + i)

    recs = _pool_map(make, range(500))
    m = amb.Modem(cfg, mod, rep, device=0)
    try:
        got = m.decode_received_signals(recs, raw=True)
    finally:
        m.close()
    want = _pool_map(lambda r: O.decodeReceivedSignal(cfg, r, mod, rep, raw=True), recs)
    ok = 0
    for i, ((r1, b1), (r2, b2)) in enumerate(zip(want, got)):
        assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], (mode, i, results_equal(r1, b1, r2, b2, metric_tol=1e-9))
        ok += int(r1.status == 0 and r1.crc_valid)
    assert ok >= 300


def test_c4_long_capture_20m_samples(O, amb):
    """config 4 at an eighth of its length: a 21 M-sample narrowband capture (8 minutes) over a -30 dBFS noise floor with a
    metadata frame and four data frames far apart, as one push and as pushes of 1,024 blocks: the long idle stretches make the
    app.js:826 defect random-walk (false positives, frame errors) - every event and counter, falsePositives and frameErrors
    included, equal to the oracle's"""
    cfg, mod, rep, chunk = MODES["BPSK-NARROW"]
    n = 5120 * 4096
    rng = np.random.default_rng(4)
    x = (rng.standard_normal(n) * 10 ** (-30 / 20)).astype(np.float32)
    pos = 3 * 44100
    meta = O.buildMetadataFrame(cfg, 4, 4 * chunk, chunk, "cap.bin", mod, rep)
    x[pos:pos + len(meta)] += meta
    pos += len(meta)
    for i, gap_s in enumerate((12, 47, 95, 140)):
        pos += gap_s * 44100
        f = O.buildDataChunkFrame(cfg, O.randomBytes(700 + i, chunk), i, mod, rep)
        x[pos:pos + len(f)] += f
        pos += len(f)
    assert pos < n - 8 * 4096
    ev = _events_equal(O, amb, "BPSK-NARROW", x, [n])
    assert sum(1 for r, _ in ev if r.status == 0 and r.crc_valid) >= 4
    _events_equal(O, amb, "BPSK-NARROW", x, [1024 * 4096] * 5)


def test_c5_many_streams_1024(O, amb):
    """config 5 at 1/64 of its stream count: 1,024 independent 16-QAM streams (metadata + 8 chunks each, own payload and noise)
    built on the device and pushed as [n_streams][8 x 4096] sections; 64 of them replayed by the oracle event for event, all of
    them checked through the counters"""
    import torch

    cfg, mod, rep, chunk = MODES["16-QAM"]
    ns, K = 1024, 8
    dev = torch.device("cuda", 0)
    m = amb.Modem(cfg, mod, rep, device=0)
    try:
        meta_len = m.chunk_frame_samples(1 + 4 + 4 + 2 + 1 + len("s.bin") + 4, True)
        flen = m.chunk_frame_samples(chunk + 11, False)
        push = 8 * 4096
        n = (meta_len + K * flen + 8192 + push - 1) // push * push
        x = torch.zeros((ns, n), dtype=torch.float32, device=dev)
        clean = torch.zeros((1, flen), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        m.build_data_chunk_frames(n=1, data_len=chunk, seq0=0, synth_seed=2, out=clean)
        m.synchronize()
        sigma = float((clean[0, 2205:flen - 882].double().pow(2).mean() / 10 ** 2.5).sqrt())
        for s in range(ns):
            m.build_metadata_frame(K, K * chunk, chunk, "s.bin", noise_sigma=sigma, noise_seed=5000 + s, noise_g0=0, out=x[s, :meta_len])
            m.build_data_chunk_frames(n=K, data_len=chunk, seq0=0, synth_seed=9000 + s, noise_sigma=sigma, noise_seed=5000 + s, noise_g0=meta_len,
                                      out=x[s, meta_len:meta_len + K * flen].view(K, flen))
        m.synchronize()
        m.stream_open(ns, 4096, push)
        got = [[] for _ in range(ns)]
        for p0 in range(0, n, push):
            for r, b in m.stream_push(x[:, p0:p0 + push], raw=True):
                got[r.stream].append((r, b))
        c = m.stream_counters()
        assert c["metaReceived"] == ns and c["framesDecoded"] == ns * (K + 1) and c["frameErrors"] == 0
        assert c["receivedCount"] + c["crcErrors"] == ns * K
        pick = list(range(0, ns, 16))
        host = x[pick].cpu().numpy()
        wants = _pool_map(lambda i: O.StreamingReceiver(cfg, mod, rep).push(host[i], 4096, raw=True), range(len(pick)))
        for i, s in enumerate(pick):
            g = sorted(got[s], key=lambda t: t[0].preamble_idx)
            assert len(g) == len(wants[i]) == K + 1
            for (r1, b1), (r2, b2) in zip(wants[i], g):
                assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == [], (s, r1.preamble_idx)
        m.stream_close()
    finally:
        m.close()
