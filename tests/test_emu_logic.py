"""Device LOGIC vs the oracle, on the CPU: tests/emu/emu.cpp compiles the very .cuh sources the CUDA kernels are
built from (warp loops run sequentially) and drives them as the kernels do.  This checks the faithful FFT
schedule, equaliser, demap slicing, vote/pack/parse/CRC, the streaming walker (tile skipping, the reference's
block-boundary quirk, refine/collect/reset) and the metadata rewind rule; the GPU parity tests then only have to
establish that the parallel decomposition computes the same thing."""
import numpy as np
import pytest

from helpers import MODES, emu_decode_chunk_frame, emu_stream, make_stream, results_equal


@pytest.mark.parametrize("mode", list(MODES))
def test_decode_chunk_frame_bit_exact(O, emu, mode):
    cfg, mod, rep, chunk = MODES[mode]
    lead = 13230 if cfg == "standard" else 22050
    cases = [
        O.addAwgn(O.buildDataChunkFrame(cfg, O.randomBytes(9, chunk), 5, mod, rep), 0.01, 5)[2205:],
        O.addAwgn(O.buildMetadataFrame(cfg, 7, 1234, chunk, "näme.bin", mod, rep), 0.3, 6)[lead:],      # hopeless SNR
        (np.random.default_rng(1).standard_normal(20000) * 0.1).astype(np.float32),                     # garbage
        np.zeros(5000, np.float32),                                                                      # silence: eq = 0/0 guards
        O.buildDataChunkFrame(cfg, O.randomBytes(9, 16), 1, mod, rep)[2205:][: 3 * O.config(cfg)["SYMBOL_LEN"] + 5],
        np.zeros(100, np.float32),
    ]
    for i, x in enumerate(cases):
        r1, b1 = O.decodeChunkFrame(cfg, x, mod, rep, raw=True)
        r2, b2 = emu_decode_chunk_frame(emu, O, cfg, mod, rep, x)
        assert results_equal(r1, b1, r2, b2, check_metrics=False) == [], (mode, i)


def _compare_stream(O, emu, mode, x, push=None):
    cfg, mod, rep, _ = MODES[mode]
    rx = O.StreamingReceiver(cfg, mod, rep)
    want = rx.push(x, 4096, raw=True)
    got, cnt = emu_stream(emu, O, cfg, mod, rep, x, push=push)
    assert len(got) == len(want)
    for (r1, b1), (r2, b2) in zip(want, got):
        assert results_equal(r1, b1, r2, b2, metric_tol=1e-9) == []
    c = rx.counters()
    for k in ("framesDecoded", "frameErrors", "falsePositives", "crcErrors", "metaReceived", "receivedCount", "duplicates"):
        assert c[k] == cnt[k], k
    return want


@pytest.mark.parametrize("mode", list(MODES))
def test_stream_matches_oracle(O, emu, mode):
    x, _ = make_stream(O, mode, 3, 3)
    _compare_stream(O, emu, mode, x)
    _compare_stream(O, emu, mode, x, push=8 * 4096)


@pytest.mark.parametrize("lead", [0, 777, 3000, 4000])
def test_stream_block_alignment(O, emu, lead):
    # different alignments of the frames against the 4096-sample call cadence move the end-of-buffer commits
    x, pay = make_stream(O, "16-QAM", 5, 11, lead=lead, gaps=[0, 100, 5000, 333])
    ev = _compare_stream(O, emu, "16-QAM", x)
    assert sum(1 for r, _ in ev if r.crc_valid) == 6


def test_stream_low_and_high_snr(O, emu):
    x, _ = make_stream(O, "QPSK", 4, 12, snr_db=8.0)
    _compare_stream(O, emu, "QPSK", x)
    x, _ = make_stream(O, "16-QAM", 4, 13, snr_db=60.0)
    _compare_stream(O, emu, "16-QAM", x)


def test_stream_noise_free_plateau(O, emu):
    # exact ties on the plateau: the first plateau sample must win (reference: sums unchanged by exact zeros)
    x, _ = make_stream(O, "16-QAM", 3, 14, noise=False)
    _compare_stream(O, emu, "16-QAM", x)


def test_stream_without_metadata(O, emu):
    # data frames only: every frame is cut by the 280-byte window and fails to parse the same way
    x, _ = make_stream(O, "16-QAM", 3, 15, meta=False, lead=5000)
    _compare_stream(O, emu, "16-QAM", x)


def test_stream_long_idle_noise_drift(O, emu):
    # 60 calls of pure noise before the frames: the reference's block-boundary defect accumulates in its sliding
    # sums; the walker must carry the same defect (Ep, Ea, Eb)
    rng = np.random.default_rng(5)
    x, _ = make_stream(O, "QPSK", 2, 16)
    noise = (rng.standard_normal(60 * 4096) * 0.03).astype(np.float32)
    _compare_stream(O, emu, "QPSK", np.concatenate([noise, x]))


def test_stream_second_metadata_changes_chunk_size(O, emu):
    cfg, mod, rep, _ = MODES["QPSK"]
    parts = [O.buildMetadataFrame(cfg, 2, 4096, 2048, "a", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(i, 2048), i, mod, rep) for i in range(2)]
    m2 = O.buildMetadataFrame(cfg, 3, 1536, 512, "b", mod, rep)
    parts.append(m2[13230 - 2205:])
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(10 + i, 512), i, mod, rep) for i in range(3)]
    from helpers import pad_blocks

    x = pad_blocks(O.addAwgn(np.concatenate(parts), 0.003, 3))
    ev = _compare_stream(O, emu, "QPSK", x)
    assert sum(1 for r, _ in ev if r.frame_type == 0xFE and r.crc_valid) >= 1


@pytest.mark.parametrize("mode", list(MODES))
def test_tx_frames_bit_identical(O, emu, mode):
    """Transmit-side device logic (amb_tx.cuh) against the oracle's buildDataChunkFrame / buildMetadataFrame."""
    import ctypes as C

    from helpers import P

    cfg, mod, rep, chunk = MODES[mode]
    emu.emu_build_data_chunk_frame.restype = C.c_int64
    emu.emu_build_metadata_frame.restype = C.c_int64
    for seq, n in ((0, chunk), (123456, 17), (-3, 1)):
        d = O.randomBytes(40 + n, n)
        want = O.buildDataChunkFrame(cfg, d, seq, mod, rep)
        out = np.zeros(len(want), np.float32)
        got_n = emu.emu_build_data_chunk_frame(O.CONFIGS[cfg], O.MODS[mod], rep, P(d, C.c_uint8), n, seq, C.c_uint64(0), P(out, C.c_float), C.c_int64(len(out)))
        assert got_n == len(want) and np.array_equal(out, want), (mode, seq, n)
    # synthetic chunk bytes: the device formula equals the oracle's mirror
    d = O.synthBytes(77, 5, chunk)
    want = O.buildDataChunkFrame(cfg, d, 5, mod, rep)
    out = np.zeros(len(want), np.float32)
    emu.emu_build_data_chunk_frame(O.CONFIGS[cfg], O.MODS[mod], rep, None, chunk, 5, C.c_uint64(77), P(out, C.c_float), C.c_int64(len(out)))
    assert np.array_equal(out, want)
    name = "ünï.bin".encode()
    want = O.buildMetadataFrame(cfg, 9, 9 * chunk - 1, chunk, name, mod, rep)
    out = np.zeros(len(want), np.float32)
    nb = np.frombuffer(name, np.uint8)
    assert emu.emu_build_metadata_frame(O.CONFIGS[cfg], O.MODS[mod], rep, 9, 9 * chunk - 1, chunk, P(nb, C.c_uint8), len(nb), P(out, C.c_float), C.c_int64(len(out))) == len(want)
    assert np.array_equal(out, want)


def test_counter_noise_matches_oracle_mirror(O, emu):
    import ctypes as C

    emu.emu_synth_normal.restype = C.c_double
    x = O.addAwgnCtr(np.zeros(4096, np.float32), 1.0, 99, g0=1000)
    z = np.array([emu.emu_synth_normal(C.c_uint64(99), C.c_int64(1000 + i)) for i in range(4096)])
    assert np.allclose(x, z.astype(np.float32), rtol=1e-6, atol=1e-7)
    big = O.addAwgnCtr(np.zeros(200000, np.float32), 1.0, 5)
    assert abs(float(big.mean())) < 0.01 and abs(float(big.std()) - 1.0) < 0.01


def test_chunked_crc32_matches_zlib(emu):
    """the end-aligned chunking + GF(2) combining tree used by the CTA-wide CRC-32 (k_parse, k_crc32)"""
    import ctypes as C
    import zlib

    emu.emu_crc32_chunked.restype = C.c_uint32
    emu.emu_crc32_chunked.argtypes = [C.c_void_p, C.c_int64, C.c_int]
    rng = np.random.default_rng(12)
    for n in [0, 1, 2, 3, 9, 127, 128, 129, 255, 256, 1000, 4103, 4107, 32777, 70001]:
        d = rng.integers(0, 256, max(n, 1), dtype=np.uint8)
        for nt in (1, 2, 32, 128):
            assert emu.emu_crc32_chunked(d.ctypes.data, n, nt) == zlib.crc32(d[:n].tobytes()), (n, nt)
    d = np.frombuffer(b"123456789", np.uint8).copy()
    assert emu.emu_crc32_chunked(d.ctypes.data, 9, 128) == 0xCBF43926


def test_scan_batch_reduction_agrees_with_serial_loop(emu):
    """amb_scan_batch (three reductions) vs the reference-order loop on random, plateau, near-tie and ramp shapes;
    the emulation is built with AMB_EMU_CHECK, which aborts the process on any disagreement"""
    import ctypes as C

    emu.emu_scan_batch_fuzz.restype = C.c_int
    emu.emu_scan_batch_fuzz.argtypes = [C.c_uint64, C.c_int]
    commits = sum(emu.emu_scan_batch_fuzz(seed, 20000) for seed in range(1, 6))
    assert commits > 1000


def test_chunk_bitmap_duplicates_and_reset(O, emu):
    """ChunkAssembler semantics without storage (app.js:610-643): retransmitted chunks count as duplicates, a chunk beyond
    totalChunks is ignored, a new metadata frame starts a fresh bitmap"""
    cfg, mod, rep, chunk = MODES["QPSK"]
    seqs = [0, 1, 1, 2, 0, 7, 2]                      # 7 >= totalChunks (4): ignored
    parts = [O.buildMetadataFrame(cfg, 4, 4 * chunk, chunk, "a.bin", mod, rep)]
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(50 + q, chunk), q, mod, rep) for q in seqs]
    parts.append(O.buildMetadataFrame(cfg, 4, 4 * chunk, chunk, "a.bin", mod, rep)[13230 - 2205:])     # same chunk size: no rewind, bitmap reset
    parts += [O.buildDataChunkFrame(cfg, O.randomBytes(50 + q, chunk), q, mod, rep) for q in (1, 3, 3)]
    x = O.addAwgn(np.concatenate(parts), 0.003, 5)
    from helpers import pad_blocks
    x = pad_blocks(x)
    rx = O.StreamingReceiver(cfg, mod, rep)
    rx.push(x, 4096)
    c = rx.counters()
    assert c["duplicates"] >= 3 and c["receivedCount"] >= 1, c          # the fixture does exercise duplicates and the reset
    _compare_stream(O, emu, "QPSK", x)
    _compare_stream(O, emu, "QPSK", x, push=8 * 4096)
