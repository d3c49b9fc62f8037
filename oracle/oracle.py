"""oracle.py — numpy/ctypes front end of the CPU oracle (oracle/spmm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product never imports it.

Parity status: PINNED — tests/test_oracle.py checks every function here against
  (i)  tests/golden/ref_*.npz: outputs of the REAL reference (oracle/_ref = its own custom_mm + matmuls, nothing
       stubbed) recorded on a B200 by tests/golden/make_ref_golden.py — cuSPARSE forward on the reference's five
       test shapes, its naive kernel (reproduced bit for bit by spmm_f32), its grad_m2, and the CSR arrays its
       get_sparse_tensor_properties built;
  (ii) tests/golden/{fwd,naive,csr,bwd}_*.npz: generated in the build container by importing the reference's
       matmuls.py against a stub custom_mm (tests/golden/make_golden.py) — shape dispatch and edge-case CSRs.

Each function names the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
_lib = None

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i64 = C.c_int64


def build(force: bool = False) -> Path:
    src = HERE / "spmm_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared", "-o", str(LIB),
                        str(src), "-lm"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.oracle_spmm_f64.argtypes = [_i32p, _i32p, _f32p, _i64, _i64, _f32p, _i64, _f64p]
        L.oracle_spmm_f32.argtypes = [_i32p, _i32p, _f32p, _i64, _i64, _f32p, _i64, _f32p, C.c_int]
        L.oracle_dense_to_csr_count.argtypes = [_f32p, _i64, _i64, _i64, _i32p]
        L.oracle_dense_to_csr_count.restype = _i64
        L.oracle_dense_to_csr_fill.argtypes = [_f32p, _i64, _i64, _i64, _i32p, _f32p]
        L.oracle_csr_transpose.argtypes = [_i32p, _i32p, _f32p, _i64, _i64, _i32p, _i32p, _f32p]
        L.oracle_sddmm_f64.argtypes = [_i32p, _i32p, _i64, _i64, _f32p, _i64, _f32p, _i64, _f64p]
        L.oracle_spmm_t_f64.argtypes = [_i32p, _i32p, _f32p, _i64, _i64, _i64, _f32p, _i64, _f64p]
        for f in (L.oracle_spmm_f64, L.oracle_spmm_f32, L.oracle_dense_to_csr_fill, L.oracle_csr_transpose,
                  L.oracle_sddmm_f64, L.oracle_spmm_t_f64):
            f.restype = None
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def spmm_f64(row_ptr, col, val, shape, B) -> np.ndarray:
    """fp64 CSR SpMM, src/naive_sparse_mm.cu:52-99 (the tolerance oracle)."""
    M, K = shape
    rp, ci, va, Bc = _i32(row_ptr), _i32(col), _f32(val), _f32(B)
    assert Bc.shape[0] == K
    N = Bc.shape[1]
    out = np.empty((M, N), dtype=np.float64)
    lib().oracle_spmm_f64(_p(rp, _i32p), _p(ci, _i32p), _p(va, _f32p), M, N, _p(Bc, _f32p), N, _p(out, _f64p))
    return out


def spmm_f32(row_ptr, col, val, shape, B, fused: bool = True) -> np.ndarray:
    """fp32 CSR-order SpMM: the naive kernel's exact arithmetic, src/naive_sparse_mm.cu:84-90."""
    M, K = shape
    rp, ci, va, Bc = _i32(row_ptr), _i32(col), _f32(val), _f32(B)
    N = Bc.shape[1]
    out = np.empty((M, N), dtype=np.float32)
    lib().oracle_spmm_f32(_p(rp, _i32p), _p(ci, _i32p), _p(va, _f32p), M, N, _p(Bc, _f32p), N,
                          _p(out, _f32p), int(fused))
    return out


def dense_to_csr(dense):
    """x != 0, row-major, int32 — matmuls.py:215,228 + :185-186; src/baseline_mm.cu:241-258."""
    d = _f32(dense)
    M, K = d.shape
    rp = np.empty(M + 1, dtype=np.int32)
    nnz = lib().oracle_dense_to_csr_count(_p(d, _f32p), M, K, K, _p(rp, _i32p))
    ci = np.empty(nnz, dtype=np.int32)
    va = np.empty(nnz, dtype=np.float32)
    lib().oracle_dense_to_csr_fill(_p(d, _f32p), M, K, K, _p(ci, _i32p), _p(va, _f32p))
    return rp, ci, va


def dense_to_csr_numpy(dense):
    """Same contract in pure numpy (cross-check of the C restatement)."""
    d = np.asarray(dense, dtype=np.float32)
    mask = d != 0
    rp = np.zeros(d.shape[0] + 1, dtype=np.int32)
    rp[1:] = np.cumsum(mask.sum(axis=1))
    r, c = np.nonzero(mask)
    return rp, c.astype(np.int32), d[r, c]


def csr_transpose(row_ptr, col, val, shape):
    """Canonical CSR of A^T, matmuls.py:253 (`m1.transpose(-1,-2)` -> `.to_sparse_csr()`)."""
    M, K = shape
    rp, ci, va = _i32(row_ptr), _i32(col), _f32(val)
    nnz = len(va)
    t_rp = np.empty(K + 1, dtype=np.int32)
    t_ci = np.empty(nnz, dtype=np.int32)
    t_va = np.empty(nnz, dtype=np.float32)
    lib().oracle_csr_transpose(_p(rp, _i32p), _p(ci, _i32p), _p(va, _f32p), M, K, _p(t_rp, _i32p),
                               _p(t_ci, _i32p), _p(t_va, _f32p))
    return t_rp, t_ci, t_va


def sddmm_f64(row_ptr, col, shape, G, B) -> np.ndarray:
    """grad_A on A's pattern in fp64: matmuls.py:250 sampled at A's non-zeros."""
    M, K = shape
    rp, ci, Gc, Bc = _i32(row_ptr), _i32(col), _f32(G), _f32(B)
    N = Gc.shape[1]
    out = np.empty(len(ci), dtype=np.float64)
    lib().oracle_sddmm_f64(_p(rp, _i32p), _p(ci, _i32p), M, N, _p(Gc, _f32p), N, _p(Bc, _f32p), N,
                           _p(out, _f64p))
    return out


def spmm_t_f64(row_ptr, col, val, shape, G) -> np.ndarray:
    """grad_B = A^T G in fp64: matmuls.py:253."""
    M, K = shape
    rp, ci, va, Gc = _i32(row_ptr), _i32(col), _f32(val), _f32(G)
    N = Gc.shape[1]
    out = np.empty((K, N), dtype=np.float64)
    lib().oracle_spmm_t_f64(_p(rp, _i32p), _p(ci, _i32p), _p(va, _f32p), M, K, N, _p(Gc, _f32p), N,
                            _p(out, _f64p))
    return out


def rel_err(x, ref64) -> float:
    """The tolerance norm of BASELINE.json: max|x - ref| / max|ref| (per output matrix)."""
    ref64 = np.asarray(ref64, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    if ref64.size == 0:
        return 0.0
    denom = np.max(np.abs(ref64))
    num = np.max(np.abs(x - ref64))
    return float(num / denom) if denom > 0 else float(num)
