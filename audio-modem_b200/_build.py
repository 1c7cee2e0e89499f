"""Builds libaudiomodem_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# AMB_LIB selects another build of the library (development: A/B comparisons of compile-time variants)
LIB = os.environ.get("AMB_LIB") or os.path.join(HERE, "libaudiomodem_b200.so")
SOURCES = ["amb_api.cu", "amb_tables.cpp"]
HEADERS = ["amb_tables.h", "amb_device.cuh", "amb_walker.cuh", "amb_kernels.cuh", "amb_tx.cuh", "amb_private.h", "../../include/audio_modem_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # -ffp-contract=off: the host table generators must round like the reference's float64 expressions
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, extra=(), out=None):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd)
    return out or LIB
