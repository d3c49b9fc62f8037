"""ctypes wrapper of tools/libcusparse_fair.so: the kernel-fair cuSPARSE comparator (cusparseSpMM CSR_ALG2 on row-major
B / C as the reference calls it, src/baseline_mm.cu:186-213, with descriptors, buffer and cusparseSpMM_preprocess
hoisted out of the timed call).  Comparator only: benchmarks and bench.py's `comparators` block use it."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "libcusparse_fair.so"
SRC = HERE / "cusparse_fair.cu"
ALGS = {"alg2": 0, "alg1": 1, "default": 2, "alg3": 3}
_lib = None


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        subprocess.run(["nvcc", "-O2", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets", "-o", str(LIB),
                        str(SRC), "-lcusparse"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        L.fair_create.restype = C.c_void_p
        L.fair_create.argtypes = [C.c_void_p] * 3 + [C.c_int64] * 3 + [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                                                     C.c_int64, C.c_int, C.c_void_p]
        L.fair_run.restype = C.c_int
        L.fair_run.argtypes = [C.c_void_p]
        L.fair_destroy.argtypes = [C.c_void_p]
        L.fair_buffer_bytes.restype = C.c_size_t
        L.fair_buffer_bytes.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class FairSpMM:
    """C = A_csr @ B with cuSPARSE, everything but the SpMM call hoisted.  Tensors are torch CUDA tensors."""

    def __init__(self, crow, col, val, shape, B, Cout, alg="alg2", stream=0):
        self._keep = (crow, col, val, B, Cout)
        M, K = shape
        self._h = lib().fair_create(crow.data_ptr(), col.data_ptr(), val.data_ptr(), M, K, val.numel(), B.data_ptr(),
                                    B.stride(0), B.shape[1], Cout.data_ptr(), Cout.stride(0), ALGS[alg], stream)
        if not self._h:
            raise RuntimeError("cusparse_fair: setup failed (see stderr)")
        self.buffer_bytes = lib().fair_buffer_bytes(self._h)

    def run(self):
        rc = lib().fair_run(self._h)
        if rc != 0:
            raise RuntimeError(f"cusparseSpMM failed with status {rc}")

    def close(self):
        if self._h:
            lib().fair_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
