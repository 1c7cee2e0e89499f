"""ctypes binding of include/audio_modem_b200.h.  No CPU path: a missing library or GPU is an error."""
import ctypes as C
import os

from . import _build

CFG = {"standard": 0, "acoustic": 1, "narrowband": 2}
MOD = {"BPSK": 0, "QPSK": 1, "QAM16": 2}
MEM_HOST, MEM_DEVICE, MEM_INPLACE = 0, 1, 2
E_OVERFLOW = -5


class FrameResult(C.Structure):
    """amb_frame_result"""
    _fields_ = [
        ("status", C.c_int32), ("frame_type", C.c_int32), ("preamble_idx", C.c_int64),
        ("seq_num", C.c_int32), ("data_len", C.c_int32), ("total_chunks", C.c_int32),
        ("total_file_size", C.c_int32), ("chunk_size", C.c_int32), ("name_len", C.c_int32),
        ("name_off", C.c_int32), ("data_off", C.c_int32), ("expected_crc", C.c_uint32),
        ("actual_crc", C.c_uint32), ("crc_valid", C.c_int32), ("n_bytes", C.c_int32),
        ("err_arg", C.c_int32), ("stream", C.c_int32), ("coarse_metric", C.c_double),
        ("fine_metric", C.c_double),
    ]


class LoopbackResult(C.Structure):
    """amb_loopback_result"""
    _fields_ = [("detected", C.c_int32), ("quality", C.c_int32), ("start_idx", C.c_int64), ("n_mag", C.c_int32), ("status", C.c_int32),
                ("correlation", C.c_double), ("ber", C.c_double), ("snr_estimate", C.c_double), ("pilot_mag", C.c_double)]


class AmbError(RuntimeError):
    pass


EXPORTS = {
    # name: (restype, argtypes)
    "amb_version": (C.c_int, []),
    "amb_status_string": (C.c_char_p, [C.c_int]),
    "amb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]),
    "amb_destroy": (None, [C.c_void_p]),
    "amb_last_error": (C.c_char_p, [C.c_void_p]),
    "amb_synchronize": (C.c_int, [C.c_void_p]),
    "amb_cuda_stream": (C.c_void_p, [C.c_void_p]),
    "amb_kernel_launches": (C.c_int64, [C.c_void_p]),
    "amb_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "amb_host_free": (C.c_int, [C.c_void_p]),
    "amb_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "amb_profile_read": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int)]),
    "amb_estimate_frame_samples": (C.c_int64, [C.c_int, C.c_int, C.c_int64, C.c_int]),
    "amb_config_info": (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "amb_modem_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "amb_preamble_symbol1": (C.c_int, [C.c_int, C.c_void_p]),
    "amb_channel_est_symbol": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    "amb_crc32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "amb_preprocess_signal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "amb_detect_preamble": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "amb_detect_preamble_xcorr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "amb_estimate_channel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "amb_demodulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "amb_decode_chunk_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "amb_decode_received_signals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "amb_analyze_loopback": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_int]),
    "amb_stream_open": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64]),
    "amb_stream_close": (C.c_int, [C.c_void_p]),
    "amb_stream_input_window": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "amb_stream_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                  C.c_void_p, C.c_int64, C.c_int]),
    "amb_stream_push_pcm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                        C.c_void_p, C.c_int64, C.c_int]),
    "amb_stream_counters": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "amb_stream_enable_assembly": (C.c_int, [C.c_void_p, C.c_int64]),
    "amb_stream_assemble": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_int]),
    "amb_transmit_signal_samples": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int]),
    "amb_build_transmit_signals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int]),
    "amb_modulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.c_int]),
    "amb_chunk_frame_samples": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int]),
    "amb_build_data_chunk_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int32, C.c_uint64, C.c_void_p,
                                              C.c_int64, C.c_double, C.c_uint64, C.c_int64, C.c_void_p, C.c_int]),
    "amb_build_metadata_frame": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int, C.c_void_p, C.c_double,
                                           C.c_uint64, C.c_int64, C.c_int]),
}

# development entry points (audio-modem_b200/csrc/amb_private.h): exported by the library, not part of the boundary header
PRIVATE_EXPORTS = {
    "amb_debug_read": (C.c_int64, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "amb_debug_counter": (C.c_int64, [C.c_void_p, C.c_int]),
    "amb_selftest": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_int64)]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load the C-ABI library.  Raises if it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(_build.LIB):
            raise AmbError(f"{_build.LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). The package has no CPU or PyTorch fallback.")
        L = C.CDLL(_build.LIB)
        for name, (res, args) in list(EXPORTS.items()) + list(PRIVATE_EXPORTS.items()):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
