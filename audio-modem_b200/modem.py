"""Host-side mirror of the reference's receive-path interface, over the C-ABI library.

The reference (playok/audio-modem) exposes browser-global JavaScript functions (modem.js) and the
StreamingReceiver class (app.js:706-998).  Node.js is not available in this image, so the host side above
the C-ABI is this Python mirror: same function names, argument meaning, result-object keys and error strings
as modem.js, including the hidden process-global `OFDM` config switched by setOFDMConfig (modem.js:87-98).
`Modem` is the explicit, batched form (one call = many recordings / frames / streams); the module-level
functions are the drop-in single-signal forms.  All compute happens in the CUDA library; there is no fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CFG, MOD, MEM_DEVICE, MEM_HOST, MEM_INPLACE, AmbError, FrameResult, LoopbackResult

FRAME_META = 0xFE   # modem.js:661
FRAME_DATA = 0xFF   # modem.js:662


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class PinnedArray:
    """numpy view of page-locked host memory from amb_host_alloc (freed by .free() or with the object)"""

    def __init__(self, shape, dtype):
        self._L = _lib.load()
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        if self._L.amb_host_alloc(C.byref(p), n) != 0:
            raise AmbError(f"amb_host_alloc({n}) failed")
        self.ptr, self.nbytes = p.value, n
        self.array = np.frombuffer((C.c_uint8 * max(n, 1)).from_address(self.ptr), dtype=dt, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            self._L.amb_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32):
    """page-locked host array (keeps its owner alive through `.base` chains: hold on to the returned PinnedArray)"""
    return PinnedArray(shape, dtype)


def _status_message(status, arg):
    s = _lib.load().amb_status_string(status).decode()
    return s.replace("{arg:x}", format(arg & 0xFFFFFFFF if arg < 0 else arg, "x")).replace("{arg}", str(arg))


def result_to_object(r, decoded):
    """amb_frame_result + decoded bytes -> the reference's result object (modem.js:644-653, :822-827, :843-848,
    error objects :564-:836)."""
    if r.status != 0:
        d = {"error": _status_message(r.status, r.err_arg)}
        if r.status == 9:
            d["frameType"] = r.err_arg
        if r.preamble_idx >= 0:
            d["preambleIdx"] = int(r.preamble_idx)
        return d
    out = {"crcValid": bool(r.crc_valid), "expectedCRC": int(r.expected_crc), "actualCRC": int(r.actual_crc)}
    b = decoded
    if r.frame_type == 0:
        out.update(frameType="legacy", dataLen=r.data_len, data=bytes(b[r.data_off:r.data_off + r.data_len]),
                   fileName=bytes(b[r.name_off:r.name_off + r.name_len]).decode("utf-8", "replace"))
    elif r.frame_type == FRAME_META:
        out.update(frameType=FRAME_META, totalChunks=r.total_chunks, totalFileSize=r.total_file_size,
                   chunkSize=r.chunk_size,
                   fileName=bytes(b[r.name_off:r.name_off + r.name_len]).decode("utf-8", "replace"))
    elif r.frame_type == FRAME_DATA:
        out.update(frameType=FRAME_DATA, seqNum=r.seq_num, dataLen=r.data_len,
                   data=bytes(b[r.data_off:r.data_off + r.data_len]))
    if r.preamble_idx >= 0:
        out["preambleIdx"] = int(r.preamble_idx)
    return out


def _pack_rows(rows, dtype=np.float32):
    """list of 1-D arrays (or a 2-D array) -> ([n][stride] contiguous array, lens int64 or None)"""
    if isinstance(rows, np.ndarray) and rows.ndim == 2:
        return np.ascontiguousarray(rows, dtype=dtype), None
    rows = [np.ascontiguousarray(r, dtype=dtype).ravel() for r in rows]
    lens = np.array([len(r) for r in rows], dtype=np.int64)
    stride = int(lens.max()) if len(rows) else 0
    stride = max(stride, 1)
    out = np.zeros((len(rows), stride), dtype=dtype)
    for i, r in enumerate(rows):
        out[i, : len(r)] = r
    return out, lens


class Modem:
    """One receive configuration on one GPU: (OFDM config, constellation, repetition) — the triple every
    reference decode call takes (app.js:60-66).  Wraps an `amb_modem*`."""

    def __init__(self, config="standard", modName="QPSK", repetition=1, device=0):
        L = _lib.load()
        self.config = config if config in CFG else "standard"      # modem.js:96
        self.modName = modName
        self.repetition = repetition or 1                          # `repetition || 1`
        if modName not in MOD:
            raise AmbError(f"unknown modulation {modName!r}")
        h = C.c_void_p()
        rc = L.amb_create(C.byref(h), device, CFG[self.config], MOD[modName], self.repetition)
        if rc == -6:
            raise AmbError("no CUDA device: audio_modem_b200 has no CPU path")
        if rc != 0:
            raise AmbError(f"amb_create failed ({rc})")
        self._h = h
        self._L = L
        info = (C.c_int32 * 8)()
        pil = (C.c_int32 * 16)()
        L.amb_config_info(CFG[self.config], info, pil)
        (self.FFT_SIZE, self.CP_LEN, self.SYMBOL_LEN, self.SAMPLE_RATE, self.SUB_START, self.SUB_END,
         n_pil, self.NUM_DATA_SUBS) = list(info)
        self.PILOTS = list(pil)[:n_pil]
        self.bps = {"BPSK": 1, "QPSK": 2, "QAM16": 4}[modName]
        self._stream_open = False
        self.bytes_stride = 0

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_stream_open", False):
                try:
                    self.stream_close()
                except Exception:
                    pass
            self._L.amb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, allow_overflow=False):
        if rc != 0 and not (allow_overflow and rc == _lib.E_OVERFLOW):
            raise AmbError(f"audio_modem_b200 error {rc}: {self._L.amb_last_error(self._h).decode()}")
        return rc

    @property
    def kernel_launches(self):
        return int(self._L.amb_kernel_launches(self._h))

    def debug_counter(self, which):
        """0 launches, 1 fast-demod fallback symbols, 2 link iterations, 3 link passes, 4 / 5 slotted / candidate tiles of the last push,
        6 phase-B links answered by phase A's own walk, 7 phase-B links walked again"""
        return int(self._L.amb_debug_counter(self._h, which))

    def debug_slots(self):
        """development: (slot_info [n][4] int32, slot extremes [n][4] float32) of the last push"""
        cap = 1 << 22
        info = np.zeros((cap // 16, 4), np.int32)
        ts = np.zeros((cap // 16, 4), np.float32)
        n = int(self._L.amb_debug_read(self._h, 0, _ptr(info), cap))
        self._L.amb_debug_read(self._h, 1, _ptr(ts), cap)
        n = max(0, min(n, cap // 16))
        return info[:n], ts[:n]

    def selftest(self, which=0, seed=1, n=100000):
        """device self-test (amb_selftest): returns (disagreements, commits)"""
        out = (C.c_int64 * 2)()
        self._ck(self._L.amb_selftest(self._h, which, seed, n, out))
        return int(out[0]), int(out[1])

    def profile(self, enable=True):
        """bracket every kernel launch with CUDA events on the handle's stream (device time per kernel)"""
        self._ck(self._L.amb_profile(self._h, int(enable)))

    def profile_read(self):
        """{kernel name: (total ms, launches)} since the last read"""
        cap = 64
        names = C.create_string_buffer(32 * cap)
        ms = (C.c_double * cap)()
        cnt = (C.c_int64 * cap)()
        n = C.c_int(0)
        self._ck(self._L.amb_profile_read(self._h, names, ms, cnt, cap, C.byref(n)))
        out = {}
        for k in range(min(n.value, cap)):
            out[names.raw[32 * k:32 * k + 32].split(b"\0")[0].decode()] = (ms[k], int(cnt[k]))
        return out

    def cuda_stream(self):
        return self._L.amb_cuda_stream(self._h)

    def synchronize(self):
        self._ck(self._L.amb_synchronize(self._h))

    def decoded_capacity(self, n_samples):
        """upper bound of bytes.length for a span of n_samples"""
        return (n_samples // self.SYMBOL_LEN + 1) * self.NUM_DATA_SUBS * self.bps // 8 + 16

    # ------------------------------------------------------------------ batched operators
    def crc32(self, rows):
        x, lens = _pack_rows([np.frombuffer(bytes(r), dtype=np.uint8) if isinstance(r, (bytes, bytearray)) else r for r in rows], np.uint8)
        out = np.zeros(len(x), dtype=np.uint32)
        self._ck(self._L.amb_crc32(self._h, _ptr(x), x.shape[1], _ptr(lens), len(x), _ptr(out), MEM_HOST))
        return out

    def preprocess_signals(self, rows):
        x, lens = _pack_rows(rows)
        out = np.zeros_like(x)
        self._ck(self._L.amb_preprocess_signal(self._h, _ptr(x), x.shape[1], _ptr(lens), len(x), _ptr(out), MEM_HOST))
        return [out[i, : (lens[i] if lens is not None else x.shape[1])] for i in range(len(x))]

    def _detect(self, fn, rows):
        x, lens = _pack_rows(rows)
        idx = np.zeros(len(x), dtype=np.int64)
        met = np.zeros(len(x), dtype=np.float64)
        self._ck(fn(self._h, _ptr(x), x.shape[1], _ptr(lens), len(x), _ptr(idx), _ptr(met), MEM_HOST))
        return idx, met

    def detect_preamble(self, rows):
        return self._detect(self._L.amb_detect_preamble, rows)

    def detect_preamble_xcorr(self, rows):
        return self._detect(self._L.amb_detect_preamble_xcorr, rows)

    def estimate_channel(self, ce_rows, knownRe=None, knownIm=None):
        x = np.zeros((len(ce_rows), self.SYMBOL_LEN), dtype=np.float32)
        for i, r in enumerate(ce_rows):
            r = np.asarray(r, dtype=np.float32)[: self.SYMBOL_LEN]
            x[i, : len(r)] = r
        kr = None if knownRe is None else np.ascontiguousarray(knownRe, dtype=np.float64)
        ki = None if knownIm is None else np.ascontiguousarray(knownIm, dtype=np.float64)
        hr = np.zeros((len(x), 512))
        hi = np.zeros((len(x), 512))
        self._ck(self._L.amb_estimate_channel(self._h, _ptr(x), len(x), _ptr(kr), _ptr(ki), _ptr(hr), _ptr(hi), MEM_HOST))
        return hr, hi

    def demodulate(self, rows, chRe, chIm, with_eq=False):
        x, lens = _pack_rows(rows)
        n, stride = x.shape
        max_sym = stride // self.SYMBOL_LEN
        bits_stride = max(1, max_sym * self.NUM_DATA_SUBS * self.bps)
        eq_stride = max(1, max_sym * self.NUM_DATA_SUBS)
        bits = np.zeros((n, bits_stride), dtype=np.uint8)
        eq = np.zeros((n, eq_stride, 2)) if with_eq else None
        nbits = np.zeros(n, dtype=np.int64)
        hr = np.ascontiguousarray(np.broadcast_to(np.asarray(chRe, dtype=np.float64), (n, 512)))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(chIm, dtype=np.float64), (n, 512)))
        self._ck(self._L.amb_demodulate(self._h, _ptr(x), stride, _ptr(lens), n, _ptr(hr), _ptr(hi), _ptr(bits), bits_stride,
                                        _ptr(nbits), _ptr(eq), eq_stride, MEM_HOST))
        out_bits = [bits[i, : nbits[i]] for i in range(n)]
        if with_eq:
            return out_bits, [eq[i, : nbits[i] // self.bps] for i in range(n)]
        return out_bits

    def _decode(self, fn, rows, raw):
        x, lens = _pack_rows(rows)
        n, stride = x.shape
        bs = self.decoded_capacity(stride)
        res = (FrameResult * max(n, 1))()
        dec = np.zeros((max(n, 1), bs), dtype=np.uint8)
        self._ck(fn(self._h, _ptr(x), stride, _ptr(lens), n, C.cast(res, C.c_void_p), _ptr(dec), bs, MEM_HOST))
        out = []
        for i in range(n):
            b = dec[i, : max(0, min(res[i].n_bytes, bs))]
            out.append((res[i], b) if raw else result_to_object(res[i], b))
        return out

    def decode_chunk_frames(self, frames, raw=False):
        """decodeChunkFrame (modem.js:770) over many frames"""
        return self._decode(self._L.amb_decode_chunk_frames, frames, raw)

    def decode_received_signals(self, recordings, raw=False):
        """decodeReceivedSignal (modem.js:557) over many recordings"""
        return self._decode(self._L.amb_decode_received_signals, recordings, raw)

    def analyze_loopback(self, recordings, testData=bytes(range(16))):
        """analyzeLoopback (modem.js:978-1081) over many recordings -> list of
        {detected, correlation, ber, channelMagnitude, snrEstimate, quality} (+ startIdx)."""
        x, lens = _pack_rows(recordings)
        n, stride = x.shape
        td = np.frombuffer(bytes(testData), np.uint8).copy() if not isinstance(testData, np.ndarray) else np.ascontiguousarray(testData, np.uint8)
        nb = self.SUB_END - self.SUB_START + 1
        res = (LoopbackResult * max(n, 1))()
        mags = np.zeros((max(n, 1), nb), np.float64)
        self._ck(self._L.amb_analyze_loopback(self._h, _ptr(x), stride, _ptr(lens), n, _ptr(td), len(td), C.cast(res, C.c_void_p), _ptr(mags), nb,
                                              MEM_HOST))
        return [{"detected": bool(r.detected), "correlation": r.correlation, "ber": r.ber, "channelMagnitude": mags[i, : r.n_mag].copy(),
                 "snrEstimate": r.snr_estimate, "quality": ("poor", "good", "excellent")[r.quality], "startIdx": int(r.start_idx)}
                for i, r in enumerate(res[:n])]

    # ------------------------------------------------------------------ transmit side
    def chunk_frame_samples(self, payload_bytes, first=False):
        """estimateFrameSamplesWithSilence (modem.js:876-884)"""
        return int(self._L.amb_chunk_frame_samples(self._h, payload_bytes, int(first)))

    def build_data_chunk_frames(self, chunks=None, seq0=0, n=None, data_len=None, synth_seed=0, noise_sigma=0.0, noise_seed=0,
                                noise_g0=0, out=None, return_data=False):
        """buildDataChunkFrame (modem.js:763) for many chunks.  chunks: [n][data_len] uint8 (or list of equal-length bytes),
        or None for synthetic chunk bytes.  out: optional torch.cuda float32 tensor [n][stride] to fill on the device."""
        if chunks is not None:
            data = np.ascontiguousarray(np.stack([np.frombuffer(bytes(c), np.uint8) if isinstance(c, (bytes, bytearray)) else np.asarray(c, np.uint8)
                                                  for c in chunks]))
            n, data_len = data.shape
        else:
            data = None
        flen = self.chunk_frame_samples(data_len + 11, False)
        copy = np.zeros((n, max(data_len, 1)), np.uint8)[:, :data_len] if return_data else None
        if copy is not None:
            copy = np.ascontiguousarray(copy)
        if out is None:
            frames = np.zeros((n, flen), np.float32)
            self._ck(self._L.amb_build_data_chunk_frames(self._h, _ptr(data), data_len if data is not None else 0, data_len, n, seq0, synth_seed,
                                                        _ptr(frames), flen, noise_sigma, noise_seed, noise_g0, _ptr(copy), MEM_HOST))
            return (frames, copy) if return_data else frames
        assert _is_torch(out) and out.is_cuda and data is None and not return_data
        self._ck(self._L.amb_build_data_chunk_frames(self._h, None, 0, data_len, n, seq0, synth_seed, _ptr(out), out.stride(0), noise_sigma,
                                                    noise_seed, noise_g0, None, MEM_DEVICE))
        return out

    def modulate(self, bits_rows):
        """modulateOFDM (modem.js:322-362) for rows of equal length: -> (samples [n][numSymbols][SYMBOL_LEN] float32,
        numSymbols, bitsPerSymbol, padded [n][paddedLength] uint8: the rows as the reference leaves its caller's array)"""
        rows = [np.asarray(b, dtype=np.int64) & 1 for b in bits_rows]
        n, n_bits = len(rows), (len(rows[0]) if rows else 0)
        assert all(len(r) == n_bits for r in rows), "rows of equal length"
        out3 = (C.c_int32 * 3)()
        self._ck(self._L.amb_modulate(self._h, None, n_bits, n_bits, 0, None, 0, out3, MEM_HOST))
        nsym, per, padded = int(out3[0]), int(out3[1]), int(out3[2])
        bits = np.zeros((max(n, 1), max(padded, 1)), np.uint8)
        for i, r in enumerate(rows):
            bits[i, :n_bits] = r
        samples = np.zeros((max(n, 1), max(nsym, 1) * self.SYMBOL_LEN), np.float32)
        if n and nsym:
            self._ck(self._L.amb_modulate(self._h, _ptr(bits), bits.shape[1], n_bits, n, _ptr(samples), samples.shape[1], out3, MEM_HOST))
        return samples[:n, : nsym * self.SYMBOL_LEN].reshape(n, nsym, self.SYMBOL_LEN), nsym, per, bits[:n, :padded]

    def build_transmit_signals(self, datas, fileName="file"):
        """buildTransmitSignal (modem.js:498-555) for many payloads of equal length -> [n][samples] float32"""
        data = np.ascontiguousarray(np.stack([np.frombuffer(bytes(d), np.uint8) if isinstance(d, (bytes, bytearray)) else np.asarray(d, np.uint8)
                                              for d in datas]))
        n, data_len = data.shape
        name = np.ascontiguousarray(np.frombuffer((fileName or "file").encode("utf-8"), np.uint8)[:255])
        flen = int(self._L.amb_transmit_signal_samples(self._h, data_len, len(name)))
        out = np.zeros((n, flen), np.float32)
        self._ck(self._L.amb_build_transmit_signals(self._h, _ptr(data), max(data_len, 1) if data_len else 0, data_len, n, _ptr(name), len(name),
                                                    _ptr(out), flen, MEM_HOST))
        return out

    def build_metadata_frame(self, totalChunks, totalFileSize, chunkSize, fileName="file", noise_sigma=0.0, noise_seed=0, noise_g0=0, out=None):
        """buildMetadataFrame (modem.js:758)"""
        name = np.frombuffer((fileName or "file").encode("utf-8"), np.uint8)
        name = np.ascontiguousarray(name[:255])
        flen = self.chunk_frame_samples(1 + 4 + 4 + 2 + 1 + len(name) + 4, True)
        if out is None:
            frame = np.zeros(flen, np.float32)
            self._ck(self._L.amb_build_metadata_frame(self._h, totalChunks, totalFileSize, chunkSize, _ptr(name), len(name), _ptr(frame),
                                                     noise_sigma, noise_seed, noise_g0, MEM_HOST))
            return frame
        assert _is_torch(out) and out.is_cuda and out.numel() >= flen
        self._ck(self._L.amb_build_metadata_frame(self._h, totalChunks, totalFileSize, chunkSize, _ptr(name), len(name), _ptr(out),
                                                 noise_sigma, noise_seed, noise_g0, MEM_DEVICE))
        return flen

    # ------------------------------------------------------------------ streaming
    def stream_open(self, n_streams=1, block_len=4096, max_push=4096 * 64, max_results=None, max_chunk=4096):
        self._ck(self._L.amb_stream_open(self._h, n_streams, block_len, max_push))
        self._stream_open = True
        self.n_streams, self.block_len, self.max_push = n_streams, block_len, max_push
        # decoded-byte rows are sized for `max_chunk` payload bytes: the reference's senders never exceed 4096
        # (app.js:195-199); a stream announcing a larger chunkSize needs stream_open(max_chunk=...)
        self.bytes_stride = self.decoded_capacity(int(self._L.amb_estimate_frame_samples(CFG[self.config], MOD[self.modName], max_chunk + 16, self.repetition)))
        min_frame = int(self._L.amb_estimate_frame_samples(CFG[self.config], MOD[self.modName], 11, self.repetition))
        self.max_results = max_results or n_streams * (max_push // min_frame + 3)
        # results and decoded bytes land in page-locked memory (full-rate, asynchronous device->host copies)
        self._res_pin = PinnedArray((self.max_results * C.sizeof(FrameResult),), np.uint8)
        self._dec_pin = PinnedArray((self.max_results, self.bytes_stride), np.uint8)
        self._res = (FrameResult * self.max_results).from_address(self._res_pin.ptr)
        self._dec = self._dec_pin.array

    def stream_close(self):
        if self._stream_open:
            self._ck(self._L.amb_stream_close(self._h))
            self._stream_open = False
            self._res = self._dec = None
            self._res_pin.free()
            self._dec_pin.free()

    def stream_input_window(self):
        p = C.c_void_p()
        s = C.c_int64()
        self._ck(self._L.amb_stream_input_window(self._h, C.byref(p), C.byref(s)))
        return p.value, s.value

    def stream_push(self, x=None, n_new=None, raw=False, collect=True, device_results=None):
        """x: [n_streams][n_new] float32, or int16 for 16-bit PCM (numpy = host, torch.cuda tensor = device, None = already
        written in place).
        device_results: optional (results, bytes) torch.cuda uint8 tensors ([max_results * sizeof(amb_frame_result)] and
        [max_results][bytes_stride]) that receive the frames on the device instead of the host buffers; returns the count."""
        if not self._stream_open:
            raise AmbError("stream not open")
        push = self._L.amb_stream_push
        if x is None:
            mem, ptr, xs = MEM_INPLACE, None, 0
        elif _is_torch(x):
            assert x.is_cuda and x.dim() == 2 and x.stride(1) == 1
            if x.element_size() == 2 and not x.dtype.is_floating_point:
                push = self._L.amb_stream_push_pcm16          # torch.int16: 16-bit PCM, sample = s / 32768
            else:
                assert x.dtype.is_floating_point and x.element_size() == 4
            mem, ptr, xs, n_new = MEM_DEVICE, _ptr(x), x.stride(0), x.shape[1]
        else:
            if isinstance(x, np.ndarray) and x.dtype == np.int16:
                x = np.ascontiguousarray(x)
                push = self._L.amb_stream_push_pcm16
            else:
                x = np.ascontiguousarray(x, dtype=np.float32)
            if x.ndim == 1:
                x = x[None, :]
            mem, ptr, xs, n_new = MEM_HOST, _ptr(x), x.shape[1], x.shape[1]
        n = C.c_int(0)
        if device_results is not None:
            dres, dbytes = device_results
            assert dres.is_cuda and dbytes.is_cuda and dbytes.stride(0) == self.bytes_stride
            cap = min(dres.numel() // C.sizeof(FrameResult), dbytes.shape[0])
            rc = push(self._h, ptr, xs, n_new, mem, _ptr(dres), cap, C.byref(n), _ptr(dbytes), self.bytes_stride, MEM_DEVICE)
            self._ck(rc, allow_overflow=True)
            self.last_push_frames = n.value
            return n.value
        rc = push(self._h, ptr, xs, n_new, mem, C.cast(self._res, C.c_void_p), self.max_results, C.byref(n),
                  _ptr(self._dec), self.bytes_stride, MEM_HOST)
        self._ck(rc, allow_overflow=True)
        self.last_push_frames = n.value
        if not collect:
            return n.value
        out = []
        for i in range(min(n.value, self.max_results)):
            r = self._res[i]
            b = self._dec[i, : max(0, min(r.n_bytes, self.bytes_stride))].copy()
            if raw:
                rr = FrameResult()
                C.memmove(C.byref(rr), C.byref(r), C.sizeof(FrameResult))
                out.append((rr, b))
            else:
                o = result_to_object(r, b)
                o["stream"] = r.stream
                out.append(o)
        return out

    def stream_enable_assembly(self, max_file_bytes):
        """ChunkAssembler without IndexedDB: keep the received chunks of every stream on the device (app.js:597-704)"""
        self._ck(self._L.amb_stream_enable_assembly(self._h, int(max_file_bytes)))
        self._file_cap = int(max_file_bytes)

    def stream_assemble(self, stream=0):
        """assembleFile (app.js:666-682): the file of the metadata frame in force, missing chunks as zeros"""
        out = np.zeros(max(1, self._file_cap), dtype=np.uint8)
        size = C.c_int64(0)
        self._ck(self._L.amb_stream_assemble(self._h, stream, _ptr(out), len(out), C.byref(size), MEM_HOST))
        return bytes(out[: size.value])

    def stream_counters(self, stream=-1):
        a = (C.c_int64 * 8)()
        self._ck(self._L.amb_stream_counters(self._h, stream, a))
        keys = ["framesDecoded", "frameErrors", "falsePositives", "receivedCount", "crcErrors", "duplicates", "metaReceived", "samplesWritten"]
        return dict(zip(keys, [int(v) for v in a]))


# ---------------------------------------------------------------------------------------------------
# Drop-in single-signal forms with the reference's names and its hidden global config
# ---------------------------------------------------------------------------------------------------
_state = {"config": "standard", "device": 0}
_handles = {}


def setDevice(device):
    _state["device"] = device


def setOFDMConfig(name):
    """modem.js:95-98 — unknown names select `standard`"""
    _state["config"] = name if name in CFG else "standard"


def _modem(modName="QPSK", repetition=1):
    key = (_state["config"], modName, repetition or 1, _state["device"])
    if key not in _handles:
        _handles[key] = Modem(key[0], modName, key[2], key[3])
    return _handles[key]


def estimateFrameSamples(payloadBytes, modName, repetition=1):
    return int(_lib.load().amb_estimate_frame_samples(CFG[_state["config"]], MOD[modName], payloadBytes, repetition or 1))


def generatePreambleSymbol1():
    m = _modem()
    out = np.zeros(m.SYMBOL_LEN, dtype=np.float32)
    _lib.load().amb_preamble_symbol1(CFG[_state["config"]], _ptr(out))
    return out


def generateChannelEstSymbol():
    m = _modem()
    out = np.zeros(m.SYMBOL_LEN, dtype=np.float32)
    known = np.zeros(512)
    _lib.load().amb_channel_est_symbol(CFG[_state["config"]], _ptr(out), _ptr(known))
    return {"samples": out, "knownRe": known, "knownIm": np.zeros(512)}


def crc32(data):
    return int(_modem().crc32([data])[0])


def preprocessSignal(signal):
    return _modem().preprocess_signals([signal])[0]


def detectPreamble(signal):
    return int(_modem().detect_preamble([signal])[0][0])


def detectPreambleCrossCorr(signal):
    return int(_modem().detect_preamble_xcorr([signal])[0][0])


def estimateChannel(receivedSamples, knownRe, knownIm):
    hr, hi = _modem().estimate_channel([receivedSamples], knownRe, knownIm)
    return [hr[0], hi[0]]


def demodulateOFDM(signal, modName, channelRe, channelIm):
    return _modem(modName).demodulate([signal], channelRe, channelIm)[0]


def modulateOFDM(bits, modName):
    """modem.js:322-362 -> {samples: [Float32Array per symbol], numSymbols, bitsPerSymbol}.  Like the reference it pads the
    CALLER's bit array in place with zeros up to a whole number of symbols (modem.js:329) when that array can grow (a list)."""
    samples, nsym, per, padded = _modem(modName).modulate([bits])
    if isinstance(bits, list) and len(bits) < padded.shape[1]:
        bits.extend([0] * (padded.shape[1] - len(bits)))
    return {"samples": [samples[0, s] for s in range(nsym)], "numSymbols": nsym, "bitsPerSymbol": per}


def decodeReceivedSignal(signal, modName, repetition=1):
    return _modem(modName, repetition).decode_received_signals([signal])[0]


def analyzeLoopback(recorded, modName, repetition=1, testData=bytes(range(16))):
    """modem.js:978-1081"""
    return _modem(modName, repetition).analyze_loopback([recorded], testData)[0]


def decodeChunkFrame(frameSamples, modName, repetition=1):
    return _modem(modName, repetition).decode_chunk_frames([frameSamples])[0]


def buildTransmitSignal(fileData, modName, fileName="file", repetition=1):
    """modem.js:498-555 -> {signal, numSymbols, bitsPerSymbol, totalBits, dataLen}"""
    m = _modem(modName, repetition)
    data = bytes(fileData)
    sig = m.build_transmit_signals([data], fileName)[0]
    name_len = min(len((fileName or "file").encode("utf-8")), 255)
    bps = m.NUM_DATA_SUBS * m.bps
    total_bits = (1 + name_len + 4 + len(data) + 4) * 8 * m.repetition
    return {"signal": sig, "numSymbols": -(-total_bits // bps), "bitsPerSymbol": bps, "totalBits": total_bits, "dataLen": len(data)}


def generateTestSignal(modName, repetition=1):
    """modem.js:913-975"""
    td = bytes(range(16))
    return {"signal": buildTransmitSignal(td, modName, "test", repetition)["signal"], "testData": np.frombuffer(td, np.uint8).copy()}


def buildDataChunkFrame(chunkData, seqNum, modName, rep=1):
    """modem.js:763-766"""
    return _modem(modName, rep).build_data_chunk_frames([chunkData], seq0=seqNum)[0]


def buildMetadataFrame(totalChunks, totalFileSize, chunkSize, fileName, modName, rep=1):
    """modem.js:758-761"""
    return _modem(modName, rep).build_metadata_frame(totalChunks, totalFileSize, chunkSize, fileName)


class StreamingReceiver:
    """app.js:706-981 without DOM/IndexedDB: processAudioBlock returns the frames it completed."""

    def __init__(self, modName, repetition=1, max_push=4096 * 16):
        self.modem = Modem(_state["config"], modName, repetition, _state["device"])
        self.modem.stream_open(1, 4096, max_push)

    def processAudioBlock(self, inputSamples):
        return self.modem.stream_push(np.asarray(inputSamples, dtype=np.float32)[None, :])

    def counters(self):
        return self.modem.stream_counters(0)

    def cleanup(self):
        self.modem.stream_close()
        self.modem.close()
