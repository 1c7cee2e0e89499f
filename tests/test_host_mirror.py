"""Host-side logic of the Python mirror (no GPU): result objects, error strings, row packing, sharding."""
import ctypes as C

import numpy as np
import pytest


def test_result_objects_follow_the_reference(amb, O):
    dec = np.frombuffer(bytes([3]) + b"abc" + (5).to_bytes(4, "big") + b"hello" + bytes(4), np.uint8)
    r = amb.FrameResult(status=0, frame_type=0, preamble_idx=13230, data_len=5, name_len=3, name_off=1, data_off=8,
                        expected_crc=7, actual_crc=7, crc_valid=1, n_bytes=len(dec))
    o = amb.result_to_object(r, dec)
    assert o == {"data": b"hello", "dataLen": 5, "fileName": "abc", "crcValid": True, "expectedCRC": 7, "actualCRC": 7,
                 "preambleIdx": 13230, "frameType": "legacy"}
    r = amb.FrameResult(status=0, frame_type=0xFF, preamble_idx=-1, seq_num=-2, data_len=2, data_off=7, crc_valid=0, expected_crc=1, actual_crc=2)
    assert amb.result_to_object(r, np.arange(16, dtype=np.uint8)) == {
        "frameType": 0xFF, "seqNum": -2, "data": bytes([7, 8]), "dataLen": 2, "crcValid": False, "expectedCRC": 1, "actualCRC": 2}
    r = amb.FrameResult(status=0, frame_type=0xFE, preamble_idx=-1, total_chunks=9, total_file_size=99, chunk_size=512, name_len=1, name_off=12, crc_valid=1)
    o = amb.result_to_object(r, np.frombuffer(bytes(12) + b"x" + bytes(4), np.uint8))
    assert (o["frameType"], o["totalChunks"], o["totalFileSize"], o["chunkSize"], o["fileName"]) == (0xFE, 9, 99, 512, "x")


def test_error_strings_are_the_reference_strings(amb, O):
    # the oracle's strings were typed from modem.js; the library's come from amb_status_string
    for status, text in O.ERROR_STRINGS.items():
        for arg in (0, 0x42, -1190902218):
            r = amb.FrameResult(status=status, err_arg=arg, preamble_idx=-1)
            want = text.format(arg=arg) if status != 9 else "Unknown frame type: 0x%x" % arg if arg >= 0 else None
            if want is None:
                continue
            assert amb.result_to_object(r, b"")["error"] == want
    assert amb.result_to_object(amb.FrameResult(status=9, err_arg=0x42, preamble_idx=-1), b"") == {"error": "Unknown frame type: 0x42", "frameType": 0x42}


def test_pack_rows(amb):
    from audio_modem_b200.modem import _pack_rows

    x, lens = _pack_rows([np.ones(3), np.ones(5), np.ones(0)])
    assert x.shape == (3, 5) and list(lens) == [3, 5, 0] and x.dtype == np.float32 and x[0, 3] == 0
    x, lens = _pack_rows(np.zeros((2, 7), np.float64))
    assert x.shape == (2, 7) and lens is None and x.dtype == np.float32


def test_set_ofdm_config_falls_back_to_standard(amb):
    amb.setOFDMConfig("acoustic")
    assert amb.estimateFrameSamples(523, "BPSK", 3) == (3 + 197) * 640
    amb.setOFDMConfig("bogus")            # modem.js:96
    assert amb.estimateFrameSamples(4107, "QAM16", 1) == 25344
    amb.setOFDMConfig("standard")


def test_host_side_symbols_match_oracle(amb, O):
    L = amb._lib.load()
    for cfg in ("standard", "acoustic", "narrowband"):
        n = O.config(cfg)["SYMBOL_LEN"]
        a = np.zeros(n, np.float32)
        L.amb_preamble_symbol1(O.CONFIGS[cfg], C.c_void_p(a.ctypes.data))
        assert np.array_equal(a, O.generatePreambleSymbol1(cfg))
        k = np.zeros(512)
        L.amb_channel_est_symbol(O.CONFIGS[cfg], C.c_void_p(a.ctypes.data), C.c_void_p(k.ctypes.data))
        ce = O.generateChannelEstSymbol(cfg)
        assert np.array_equal(a, ce["samples"]) and np.array_equal(k, ce["knownRe"])


def test_tables_match_oracle(emu, O):
    from helpers import P

    for cfg in ("standard", "acoustic", "narrowband"):
        for mod in ("BPSK", "QPSK", "QAM16"):
            pre1 = np.zeros(768, np.float32)
            known, twr, twi, pts, sc = np.zeros(512), np.zeros(512), np.zeros(512), np.zeros(32), np.zeros(8)
            assert emu.emu_tables(O.CONFIGS[cfg], O.MODS[mod], 1, P(pre1, C.c_float), P(known, C.c_double), P(twr, C.c_double),
                                  P(twi, C.c_double), P(pts, C.c_double), P(sc, C.c_double)) == 0
            bps, opts = O.constellationPoints(mod)
            assert np.array_equal(pts[: 2 << bps].reshape(-1, 2), opts)
            assert sc[3] == O.config(cfg)["NUM_DATA_SUBS"]
            p1 = O.generatePreambleSymbol1(cfg)
            assert sc[0] == float(np.sum(p1.astype(np.float64) ** 2 * 1.0)) or abs(sc[0] - float((p1.astype(np.float64) ** 2).sum())) < 1e-12
    # twiddles reproduce the reference's running product: a transform built from them equals the oracle's fft bit for bit
    rng = np.random.default_rng(3)
    x = rng.standard_normal(512)
    re, im = x.copy(), np.zeros(512)
    rev = [int(format(i, "09b")[::-1], 2) for i in range(512)]
    re = re[rev]
    half = 1
    while half < 512:
        for base in range(0, 512, 2 * half):
            for j in range(half):
                a, b = base + j, base + j + half
                wr, wi = twr[half + j], twi[half + j]
                tr = wr * re[b] - wi * im[b]
                ti = wr * im[b] + wi * re[b]
                re[b], im[b] = re[a] - tr, im[a] - ti
                re[a], im[a] = re[a] + tr, im[a] + ti
        half *= 2
    ore, oim = O.fft(x, np.zeros(512))
    assert np.array_equal(re, ore) and np.array_equal(im, oim)


def test_shard_range(amb):
    sh = amb.shard
    for n in (0, 1, 7, 8, 65536, 10000):
        for w in (1, 2, 3, 4, 8):
            rs = [sh.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
            for item in range(0, n, max(1, n // 37)):
                r = sh.owner_of(item, n, w)
                assert rs[r][0] <= item < rs[r][1]
    with pytest.raises(ValueError):
        sh.shard_range(4, 4, 4)
    assert sh.gather_counters({"a": 3, "b": 4}) == {"a": 3, "b": 4}
