"""ctypes front end of the CPU oracle (oracle/modem_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.  Function names follow modem.js so that the parity tests
read like calls into the reference.  Pinned bit for bit against the outputs of
the real modem.js / app.js (tests/golden/reference_*.json; header of modem_oracle.cpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmodem_oracle.so")

CONFIGS = {"standard": 0, "acoustic": 1, "narrowband": 2}
MODS = {"BPSK": 0, "QPSK": 1, "QAM16": 2}

ERROR_STRINGS = {
    1: "Preamble not detected",
    2: "Preamble not detected (low correlation)",
    3: "Signal too short for CE",
    4: "No data after CE",
    5: "Decoded data too short",
    6: "Decoded data too short for header",
    7: "Invalid data length: {arg}",
    8: "Frame too short for CE",
    9: "Unknown frame type: 0x{arg:x}",
    10: "Metadata frame too short",
    11: "Metadata frame truncated",
    12: "Data chunk frame too short",
    13: "Data chunk truncated",
    14: "(ring buffer overwritten)",
}


class OrcResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("frame_type", C.c_int32), ("preamble_idx", C.c_int64),
        ("seq_num", C.c_int32), ("data_len", C.c_int32), ("total_chunks", C.c_int32),
        ("total_file_size", C.c_int32), ("chunk_size", C.c_int32), ("name_len", C.c_int32),
        ("name_off", C.c_int32), ("data_off", C.c_int32), ("expected_crc", C.c_uint32),
        ("actual_crc", C.c_uint32), ("crc_valid", C.c_int32), ("n_bytes", C.c_int32),
        ("err_arg", C.c_int32), ("stream", C.c_int32), ("coarse_metric", C.c_double),
        ("fine_metric", C.c_double),
    ]


class OrcLoopback(C.Structure):
    _fields_ = [("detected", C.c_int32), ("quality", C.c_int32), ("start_idx", C.c_int64), ("n_mag", C.c_int32),
                ("pad", C.c_int32), ("correlation", C.c_double), ("ber", C.c_double), ("snr_estimate", C.c_double)]


QUALITY = ("poor", "good", "excellent")


def build(force=False):
    """Compile the oracle if needed (gcc only; never touches /root/reference)."""
    src = os.path.join(_HERE, "modem_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p, f64p, u8p, i64p, i32p = (C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int))
        sig = {
            "orc_result_size": (C.c_int, []),
            "orc_fft": (None, [f64p, f64p, C.c_int, C.c_int]),
            "orc_fft_stage_seeds": (None, [f64p, f64p]),
            "orc_lcg": (None, [C.c_int, C.c_int, i64p, f64p]),
            "orc_config": (C.c_int, [C.c_int, i32p, i32p]),
            "orc_constellation": (C.c_int, [C.c_int, f64p]),
            "orc_preamble1": (None, [C.c_int, f32p]),
            "orc_preamble2": (None, [C.c_int, f32p]),
            "orc_ce_symbol": (None, [C.c_int, f32p, f64p]),
            "orc_preprocess": (None, [f32p, C.c_int64, f32p]),
            "orc_detect_preamble": (C.c_int64, [C.c_int, f32p, C.c_int64, f64p]),
            "orc_detect_preamble_xcorr": (C.c_int64, [C.c_int, f32p, C.c_int64, f64p]),
            "orc_autocorr_metric": (None, [C.c_int, f32p, C.c_int64, C.c_double, f64p]),
            "orc_xcorr_metric": (None, [C.c_int, f32p, C.c_int64, C.c_int64, f64p]),
            "orc_estimate_channel": (None, [C.c_int, f32p, C.c_int64, f64p, f64p, f64p, f64p]),
            "orc_demodulate": (C.c_int64, [C.c_int, C.c_int, f32p, C.c_int64, f64p, f64p, u8p, f64p]),
            "orc_crc32": (C.c_uint32, [u8p, C.c_int64]),
            "orc_majority_vote": (C.c_int64, [u8p, C.c_int64, C.c_int, u8p]),
            "orc_bits_to_bytes": (C.c_int64, [u8p, C.c_int64, u8p]),
            "orc_modulate": (C.c_int64, [C.c_int, C.c_int, u8p, C.c_int64, f32p]),
            "orc_build_transmit_signal": (C.c_int64, [C.c_int, C.c_int, C.c_int, u8p, C.c_int64, u8p, C.c_int, f32p, C.c_int64]),
            "orc_build_metadata_frame": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32, u8p, C.c_int, f32p, C.c_int64]),
            "orc_build_data_chunk_frame": (C.c_int64, [C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int32, f32p, C.c_int64]),
            "orc_build_raw_chunk_frame": (C.c_int64, [C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, f32p, C.c_int64]),
            "orc_estimate_frame_samples": (C.c_int64, [C.c_int, C.c_int, C.c_int64, C.c_int]),
            "orc_decode_chunk_frame": (None, [C.c_int, C.c_int, C.c_int, f32p, C.c_int64, C.POINTER(OrcResult), u8p, C.c_int64]),
            "orc_decode_received_signal": (None, [C.c_int, C.c_int, C.c_int, f32p, C.c_int64, C.POINTER(OrcResult), u8p, C.c_int64]),
            "orc_loopback_size": (C.c_int, []),
            "orc_analyze_loopback": (None, [C.c_int, C.c_int, C.c_int, f32p, C.c_int64, u8p, C.c_int, C.POINTER(OrcLoopback), f64p]),
            "orc_stream_new": (C.c_void_p, [C.c_int, C.c_int, C.c_int]),
            "orc_stream_free": (None, [C.c_void_p]),
            "orc_stream_push": (None, [C.c_void_p, f32p, C.c_int64, C.c_int64]),
            "orc_stream_num_events": (C.c_int64, [C.c_void_p]),
            "orc_stream_event": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(OrcResult), u8p, C.c_int64]),
            "orc_stream_clear_events": (None, [C.c_void_p]),
            "orc_stream_counters": (None, [C.c_void_p, i64p]),
            "orc_stream_dc_mean": (C.c_double, [C.c_void_p]),
            "orc_random_bytes": (None, [C.c_uint64, u8p, C.c_int64]),
            "orc_add_awgn": (None, [f32p, C.c_int64, C.c_double, C.c_uint64]),
            "orc_power": (C.c_double, [f32p, C.c_int64, C.c_int64]),
            "orc_multipath": (None, [f32p, C.c_int64, i32p, f64p, C.c_int, f32p]),
            "orc_cfo": (None, [f32p, C.c_int64, C.c_double, C.c_double, f32p]),
            "orc_synth_bytes": (None, [C.c_uint64, C.c_int64, u8p, C.c_int64]),
            "orc_add_awgn_ctr": (None, [f32p, C.c_int64, C.c_double, C.c_uint64, C.c_int64]),
            "orc_bench_stream": (C.c_int64, [C.c_int, C.c_int, C.c_int, f32p, C.c_int64, C.c_int64, i64p]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        assert L.orc_result_size() == C.sizeof(OrcResult)
        assert L.orc_loopback_size() == C.sizeof(OrcLoopback)
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def _u8(x):
    if isinstance(x, (bytes, bytearray)):
        x = np.frombuffer(bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def _cfg(c):
    return CONFIGS.get(c, 0) if isinstance(c, str) else int(c)


def _mod(m):
    return MODS[m] if isinstance(m, str) else int(m)


def result_to_dict(r, bytes_arr):
    """Shape an OrcResult like the reference's result objects (modem.js:644-653, :822-827, :843-848)."""
    b = bytes_arr
    if r.status != 0:
        d = {"error": ERROR_STRINGS[r.status].format(arg=r.err_arg)}
        if r.status == 9:
            d["frameType"] = r.err_arg
        return d
    out = {"crcValid": bool(r.crc_valid), "expectedCRC": int(r.expected_crc), "actualCRC": int(r.actual_crc)}
    if r.frame_type == 0:
        out.update(frameType="legacy", dataLen=r.data_len, data=bytes(b[r.data_off:r.data_off + r.data_len]),
                   fileName=bytes(b[r.name_off:r.name_off + r.name_len]).decode("utf-8", "replace"),
                   preambleIdx=int(r.preamble_idx))
    elif r.frame_type == 0xFE:
        out.update(frameType=0xFE, totalChunks=r.total_chunks, totalFileSize=r.total_file_size, chunkSize=r.chunk_size,
                   fileName=bytes(b[r.name_off:r.name_off + r.name_len]).decode("utf-8", "replace"))
        if r.preamble_idx >= 0:
            out["preambleIdx"] = int(r.preamble_idx)
    elif r.frame_type == 0xFF:
        out.update(frameType=0xFF, seqNum=r.seq_num, dataLen=r.data_len,
                   data=bytes(b[r.data_off:r.data_off + r.data_len]))
        if r.preamble_idx >= 0:
            out["preambleIdx"] = int(r.preamble_idx)
    return out


# ------------------------------------------------------------------ modem.js names
def fft(re, im, inverse=False):
    re = np.array(re, dtype=np.float64)
    im = np.array(im, dtype=np.float64)
    lib().orc_fft(_p(re, C.c_double), _p(im, C.c_double), len(re), int(inverse))
    return re, im


def fft_stage_seeds():
    a = np.zeros(9)
    b = np.zeros(9)
    lib().orc_fft_stage_seeds(_p(a, C.c_double), _p(b, C.c_double))
    return a, b


def lcg(seed, n):
    st = np.zeros(n, dtype=np.int64)
    v = np.zeros(n)
    lib().orc_lcg(seed, n, _p(st, C.c_int64), _p(v, C.c_double))
    return st, v


def config(cfg):
    out = (C.c_int * 8)()
    pil = (C.c_int * 32)()
    lib().orc_config(_cfg(cfg), out, pil)
    keys = ["FFT_SIZE", "CP_LEN", "SYMBOL_LEN", "SAMPLE_RATE", "SUB_START", "SUB_END", "N_PILOTS", "NUM_DATA_SUBS"]
    d = dict(zip(keys, list(out)))
    d["PILOTS"] = list(pil)[: d["N_PILOTS"]]
    return d


def constellationPoints(mod):
    pts = np.zeros(32)
    bps = lib().orc_constellation(_mod(mod), _p(pts, C.c_double))
    return bps, pts[: 2 << bps].reshape(-1, 2).copy()


def generatePreambleSymbol1(cfg):
    out = np.zeros(config(cfg)["SYMBOL_LEN"], dtype=np.float32)
    lib().orc_preamble1(_cfg(cfg), _p(out, C.c_float))
    return out


def generatePreambleSymbol2(cfg):
    out = np.zeros(config(cfg)["SYMBOL_LEN"], dtype=np.float32)
    lib().orc_preamble2(_cfg(cfg), _p(out, C.c_float))
    return out


def generateChannelEstSymbol(cfg):
    out = np.zeros(config(cfg)["SYMBOL_LEN"], dtype=np.float32)
    known = np.zeros(512)
    lib().orc_ce_symbol(_cfg(cfg), _p(out, C.c_float), _p(known, C.c_double))
    return {"samples": out, "knownRe": known, "knownIm": np.zeros(512)}


def preprocessSignal(x):
    x = _f32(x)
    out = np.empty_like(x)
    lib().orc_preprocess(_p(x, C.c_float), len(x), _p(out, C.c_float))
    return out


def detectPreamble(cfg, x, with_metric=False):
    x = _f32(x)
    best = C.c_double(0)
    idx = lib().orc_detect_preamble(_cfg(cfg), _p(x, C.c_float), len(x), C.byref(best))
    return (int(idx), best.value) if with_metric else int(idx)


def detectPreambleCrossCorr(cfg, x, with_metric=False):
    x = _f32(x)
    best = C.c_double(0)
    idx = lib().orc_detect_preamble_xcorr(_cfg(cfg), _p(x, C.c_float), len(x), C.byref(best))
    return (int(idx), best.value) if with_metric else int(idx)


def autocorrMetric(cfg, x, min_energy):
    x = _f32(x)
    m = np.full(max(0, len(x) - 512 + 1), np.nan)
    lib().orc_autocorr_metric(_cfg(cfg), _p(x, C.c_float), len(x), min_energy, _p(m, C.c_double))
    return m


def xcorrMetric(cfg, x, d0, d1):
    x = _f32(x)
    m = np.zeros(d1 - d0 + 1)
    lib().orc_xcorr_metric(_cfg(cfg), _p(x, C.c_float), d0, d1, _p(m, C.c_double))
    return m


def estimateChannel(cfg, rx, knownRe, knownIm=None):
    rx = _f32(rx)
    kr = np.ascontiguousarray(knownRe, dtype=np.float64)
    ki = np.zeros(512) if knownIm is None else np.ascontiguousarray(knownIm, dtype=np.float64)
    cr = np.zeros(512)
    ci = np.zeros(512)
    lib().orc_estimate_channel(_cfg(cfg), _p(rx, C.c_float), len(rx), _p(kr, C.c_double), _p(ki, C.c_double),
                               _p(cr, C.c_double), _p(ci, C.c_double))
    return cr, ci


def demodulateOFDM(cfg, x, mod, chRe, chIm, with_eq=False):
    x = _f32(x)
    c = config(cfg)
    bps = [1, 2, 4][_mod(mod)]
    nsym = len(x) // c["SYMBOL_LEN"]
    bits = np.zeros(max(1, nsym * c["NUM_DATA_SUBS"] * bps), dtype=np.uint8)
    eq = np.zeros(max(1, nsym * c["NUM_DATA_SUBS"] * 2)) if with_eq else None
    cr = np.ascontiguousarray(chRe, dtype=np.float64)
    ci = np.ascontiguousarray(chIm, dtype=np.float64)
    n = lib().orc_demodulate(_cfg(cfg), _mod(mod), _p(x, C.c_float), len(x), _p(cr, C.c_double), _p(ci, C.c_double),
                             _p(bits, C.c_uint8), _p(eq, C.c_double) if with_eq else None)
    bits = bits[:n]
    return (bits, eq[: 2 * (n // bps)].reshape(-1, 2)) if with_eq else bits


def crc32(data):
    d = _u8(data)
    return int(lib().orc_crc32(_p(d, C.c_uint8), len(d)))


def majorityVote(bits, n):
    b = _u8(bits)
    out = np.zeros(len(b) // n + 1, dtype=np.uint8)
    k = lib().orc_majority_vote(_p(b, C.c_uint8), len(b), n, _p(out, C.c_uint8))
    return out[:k]


def bitsToBytes(bits):
    b = _u8(bits)
    out = np.zeros(len(b) // 8 + 1, dtype=np.uint8)
    k = lib().orc_bits_to_bytes(_p(b, C.c_uint8), len(b), _p(out, C.c_uint8))
    return out[:k]


def modulateOFDM(cfg, bits, mod):
    b = _u8(bits)
    n = lib().orc_modulate(_cfg(cfg), _mod(mod), _p(b, C.c_uint8), len(b), None)
    out = np.zeros((n, config(cfg)["SYMBOL_LEN"]), dtype=np.float32)
    lib().orc_modulate(_cfg(cfg), _mod(mod), _p(b, C.c_uint8), len(b), _p(out, C.c_float))
    return out


def _emit(call):
    n = call(None, 0)
    out = np.zeros(n, dtype=np.float32)
    call(_p(out, C.c_float), n)
    return out


def buildTransmitSignal(cfg, data, mod, name="file", rep=1):
    d = _u8(data)
    nm = _u8(name.encode("utf-8") if isinstance(name, str) else name)
    return _emit(lambda o, cap: lib().orc_build_transmit_signal(_cfg(cfg), _mod(mod), rep, _p(d, C.c_uint8), len(d),
                                                                 _p(nm, C.c_uint8), len(nm), o, cap))


def buildMetadataFrame(cfg, totalChunks, totalFileSize, chunkSize, name, mod, rep=1):
    nm = _u8(name.encode("utf-8") if isinstance(name, str) else name)
    return _emit(lambda o, cap: lib().orc_build_metadata_frame(_cfg(cfg), _mod(mod), rep, totalChunks, totalFileSize,
                                                                chunkSize, _p(nm, C.c_uint8), len(nm), o, cap))


def buildDataChunkFrame(cfg, chunk, seq, mod, rep=1):
    d = _u8(chunk)
    return _emit(lambda o, cap: lib().orc_build_data_chunk_frame(_cfg(cfg), _mod(mod), rep, _p(d, C.c_uint8), len(d), seq, o, cap))


def buildRawChunkFrame(cfg, payload, mod, rep=1, first=False):
    d = _u8(payload)
    return _emit(lambda o, cap: lib().orc_build_raw_chunk_frame(_cfg(cfg), _mod(mod), rep, _p(d, C.c_uint8), len(d), int(first), o, cap))


def estimateFrameSamples(cfg, payloadBytes, mod, rep=1):
    return int(lib().orc_estimate_frame_samples(_cfg(cfg), _mod(mod), payloadBytes, rep))


def _decode(fn, cfg, x, mod, rep, raw=False):
    x = _f32(x)
    c = config(cfg)
    cap = (len(x) // c["SYMBOL_LEN"] + 1) * c["NUM_DATA_SUBS"] * 4 // 8 + 16
    buf = np.zeros(cap, dtype=np.uint8)
    r = OrcResult()
    fn(_cfg(cfg), _mod(mod), rep, _p(x, C.c_float), len(x), C.byref(r), _p(buf, C.c_uint8), cap)
    b = buf[: max(0, r.n_bytes)]
    return (r, b) if raw else result_to_dict(r, b)


def decodeChunkFrame(cfg, frame, mod, rep=1, raw=False):
    return _decode(lib().orc_decode_chunk_frame, cfg, frame, mod, rep, raw)


def decodeReceivedSignal(cfg, x, mod, rep=1, raw=False):
    return _decode(lib().orc_decode_received_signal, cfg, x, mod, rep, raw)


TEST_DATA = bytes(range(16))      # modem.js:914-916


def generateTestSignal(cfg, mod, rep=1):
    """modem.js:913-975: the same construction as buildTransmitSignal over 00..0F named 'test'."""
    return {"signal": buildTransmitSignal(cfg, TEST_DATA, mod, "test", rep), "testData": np.frombuffer(TEST_DATA, dtype=np.uint8).copy()}


def analyzeLoopback(cfg, recorded, mod, rep=1, testData=TEST_DATA):
    """modem.js:978-1081 -> {detected, correlation, ber, channelMagnitude, snrEstimate, quality} (+ startIdx)."""
    x = _f32(recorded)
    td = _u8(testData)
    c = config(cfg)
    mags = np.zeros(c["SUB_END"] - c["SUB_START"] + 1, dtype=np.float64)
    r = OrcLoopback()
    lib().orc_analyze_loopback(_cfg(cfg), _mod(mod), rep, _p(x, C.c_float), len(x), _p(td, C.c_uint8), len(td), C.byref(r),
                               _p(mags, C.c_double))
    return {"detected": bool(r.detected), "correlation": r.correlation, "ber": r.ber, "channelMagnitude": mags[: r.n_mag].copy(),
            "snrEstimate": r.snr_estimate, "quality": QUALITY[r.quality], "startIdx": int(r.start_idx)}


class StreamingReceiver:
    """app.js:706-981 without DOM/IndexedDB; frame results are returned instead of logged."""

    def __init__(self, cfg, mod, rep=1):
        self.cfg, self.mod, self.rep = cfg, mod, rep
        self._h = lib().orc_stream_new(_cfg(cfg), _mod(mod), rep)
        c = config(cfg)
        self._cap = (estimateFrameSamples(cfg, 4096 + 16, mod, rep) // c["SYMBOL_LEN"] + 1) * c["NUM_DATA_SUBS"] * 4 // 8 + 16

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_stream_free(self._h)
            self._h = None

    def processAudioBlock(self, block):
        return self.push(block, len(block))

    def push(self, x, block_len=4096, raw=False):
        x = _f32(x)
        lib().orc_stream_push(self._h, _p(x, C.c_float), len(x), block_len)
        out = []
        n = lib().orc_stream_num_events(self._h)
        for i in range(n):
            r = OrcResult()
            buf = np.zeros(self._cap, dtype=np.uint8)
            lib().orc_stream_event(self._h, i, C.byref(r), _p(buf, C.c_uint8), self._cap)
            b = buf[: max(0, r.n_bytes)].copy()
            if raw:
                out.append((r, b))
            else:
                d = result_to_dict(r, b)
                d["preambleIdx"] = int(r.preamble_idx)
                out.append(d)
        lib().orc_stream_clear_events(self._h)
        return out

    def counters(self):
        a = np.zeros(10, dtype=np.int64)
        lib().orc_stream_counters(self._h, _p(a, C.c_int64))
        keys = ["framesDecoded", "frameErrors", "falsePositives", "receivedCount", "crcErrors", "duplicates",
                "metaReceived", "state", "acScanPos", "totalWritten"]
        return dict(zip(keys, [int(v) for v in a]))

    def dcMean(self):
        return float(lib().orc_stream_dc_mean(self._h))


# ------------------------------------------------------------------ synthetic channel
def randomBytes(seed, n):
    out = np.zeros(n, dtype=np.uint8)
    lib().orc_random_bytes(seed, _p(out, C.c_uint8), n)
    return out


def addAwgn(x, sigma, seed):
    x = _f32(x).copy()
    lib().orc_add_awgn(_p(x, C.c_float), len(x), float(sigma), seed)
    return x


def power(x, a=0, b=None):
    x = _f32(x)
    return float(lib().orc_power(_p(x, C.c_float), a, len(x) if b is None else b))


def multipath(x, delays, gains):
    x = _f32(x)
    out = np.empty_like(x)
    d = (C.c_int * len(delays))(*delays)
    g = (C.c_double * len(gains))(*gains)
    lib().orc_multipath(_p(x, C.c_float), len(x), d, g, len(delays), _p(out, C.c_float))
    return out


def cfo(x, df, fs=44100.0):
    x = _f32(x)
    out = np.empty_like(x)
    lib().orc_cfo(_p(x, C.c_float), len(x), float(df), float(fs), _p(out, C.c_float))
    return out


def synthBytes(seed, frame, n):
    out = np.zeros(n, dtype=np.uint8)
    lib().orc_synth_bytes(seed, frame, _p(out, C.c_uint8), n)
    return out


def addAwgnCtr(x, sigma, seed, g0=0):
    x = _f32(x).copy()
    lib().orc_add_awgn_ctr(_p(x, C.c_float), len(x), float(sigma), seed, g0)
    return x


def benchStream(cfg, mod, rep, x, block_len=4096):
    x = _f32(x)
    pb = C.c_int64(0)
    ok = lib().orc_bench_stream(_cfg(cfg), _mod(mod), rep, _p(x, C.c_float), len(x), block_len, C.byref(pb))
    return int(ok), int(pb.value)
