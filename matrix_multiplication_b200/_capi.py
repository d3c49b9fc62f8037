"""ctypes binding of the C ABI in include/spmm_b200.h (libspmm_b200.so).

This is the stub a maintainer of a non-torch host would write; the
parity tests call the engine through it with plain numpy host buffers.  There is no CPU
fallback: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libspmm_b200.so"
HEADER = PKG.parent / "include" / "spmm_b200.h"

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC = range(5)
FLAG_NO_STAGE = 1
OP_B_TRANSPOSED, OP_C_TRANSPOSED, OP_RELU, OP_NO_TC, OP_FORCE_TC = 1, 2, 4, 8, 16
TC_KIND_TF32, TC_KIND_F16 = 0, 1
TC_KIND_DEFAULT = TC_KIND_F16


class B200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"spmm_b200 status {status}: {message}")
        self.status = status


_lib = None

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p
_i64 = C.c_int64

# name -> (restype, argtypes); every symbol include/spmm_b200.h declares must be listed here
SIGNATURES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_abi_version": (C.c_int, []),
    "b200_build_info": (C.c_char_p, []),
    "b200_kernel_launch_count": (C.c_uint64, []),
    "b200_smoke_kernel": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "b200_set_option": (C.c_int, [C.c_char_p, _i64]),
    "b200_get_option": (_i64, [C.c_char_p]),
    "b200_spmm_workspace_bytes": (C.c_size_t, [_i64, _i64, _i64]),
    "b200_spmm_csr_forward": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp,
                                        C.c_size_t, _vp]),
    "b200_spmm_csr_forward_flags": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64,
                                              _vp, C.c_size_t, _vp, C.c_uint32]),
    "b200_sddmm_csr": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _i64, _vp, _vp]),
    "b200_csr_transpose": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "b200_gather_f32": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "b200_dense_to_csr_count": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i64p, _vp]),
    "b200_dense_to_csr_fill": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "b200_dense_to_csr_onepass_scratch_bytes": (C.c_size_t, [_i64, _i64]),
    "b200_dense_to_csr_onepass": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, C.c_size_t, _i64p, _vp]),
    "b200_dense_to_csr_alloc": (C.c_int, [_vp, _i64, _i64, _i64, C.POINTER(_vp), C.POINTER(_vp),
                                          C.POINTER(_vp), _i64p, _vp]),
    "b200_free_csr": (C.c_int, [_vp, _vp, _vp]),
    "b200_csr_block_diag_cols": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "b200_plan_create": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, C.POINTER(_vp), _vp]),
    "b200_plan_destroy": (C.c_int, [_vp]),
    "b200_plan_values_updated": (C.c_int, [_vp]),
    "b200_plan_stats": (C.c_int, [_vp, _i64p, _i64p]),
    "b200_plan_forward": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "b200_plan_forward_host": (C.c_int, [_vp, _vp, _i64, _vp, C.c_int]),
    "b200_plan_forward_peers": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, C.POINTER(_vp), C.c_int, _vp, _vp]),
    "b200_plan_forward_ex": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, C.c_uint32, _vp]),
    "b200_plan_backward_B_ex": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, C.c_uint32, _vp]),
    "b200_plan_backward_A_ex": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, C.c_uint32, _vp]),
    "b200_plan_would_use_tc": (C.c_int, [_vp, C.c_int, _i64, C.c_uint32]),
    "b200_plan_tc_info": (C.c_int, [_vp, C.c_int, _i64p, _i64p, C.POINTER(C.c_int)]),
    "b200_tc_last_timeout": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "b200_plan_backward_B": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "b200_plan_backward_A": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp]),
    "b200_spmm_csr_forward_host": (C.c_int, [_i32p, _i32p, _f32p, _i64, _i64, _i64, _f32p, _i64, _f32p]),
    "b200_sddmm_csr_host": (C.c_int, [_i32p, _i32p, _i64, _i64, _i64, _f32p, _f32p, _i64, _f32p]),
    "b200_spmm_csr_backward_B_host": (C.c_int, [_i32p, _i32p, _f32p, _i64, _i64, _i64, _f32p, _i64, _f32p]),
    "b200_csr_transpose_host": (C.c_int, [_i32p, _i32p, _f32p, _i64, _i64, _i64, _i32p, _i32p, _f32p]),
    "b200_dense_to_csr_host": (C.c_int, [_f32p, _i64, _i64, _i32p, _i32p, _f32p, _i64, _i64p]),
    "b200_merge_path_boundary": (C.c_int, [_i32p, _i64, _i64, _i64, _i64, _i64p, _i64p]),
    "b200_shard_rows": (C.c_int, [_i32p, _i64, C.c_int, C.c_double, C.c_double, _i64p]),
}


def declared_symbols() -> list[str]:
    """Every function name declared in include/spmm_b200.h."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[A-Za-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m matrix_multiplication_b200.build` "
                "(the engine has no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != OK:
        raise B200Error(rc, lib().b200_last_error().decode())


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


# ---- host-buffer calls (numpy in, numpy out) ---------------------------------------------------
def spmm_host(row_ptr, col_ind, val, shape, B) -> np.ndarray:
    """C = A_csr @ B with host buffers (H2D, kernels, D2H inside the call)."""
    M, K = shape
    rp, rp_p = _i32(row_ptr)
    ci, ci_p = _i32(col_ind)
    va, va_p = _f32(val)
    Bc, B_p = _f32(B)
    assert Bc.ndim == 2 and Bc.shape[0] == K, (Bc.shape, K)
    N = Bc.shape[1]
    out = np.empty((M, N), dtype=np.float32)
    check(lib().b200_spmm_csr_forward_host(rp_p, ci_p, va_p, M, K, len(va), B_p, N, out.ctypes.data_as(_f32p)))
    return out


def sddmm_host(row_ptr, col_ind, shape, G, B) -> np.ndarray:
    M, K = shape
    rp, rp_p = _i32(row_ptr)
    ci, ci_p = _i32(col_ind)
    Gc, G_p = _f32(G)
    Bc, B_p = _f32(B)
    N = Gc.shape[1]
    assert Gc.shape == (M, N) and Bc.shape == (K, N)
    out = np.empty(len(ci), dtype=np.float32)
    check(lib().b200_sddmm_csr_host(rp_p, ci_p, M, K, len(ci), G_p, B_p, N, out.ctypes.data_as(_f32p)))
    return out


def backward_B_host(row_ptr, col_ind, val, shape, G) -> np.ndarray:
    M, K = shape
    rp, rp_p = _i32(row_ptr)
    ci, ci_p = _i32(col_ind)
    va, va_p = _f32(val)
    Gc, G_p = _f32(G)
    N = Gc.shape[1]
    assert Gc.shape[0] == M
    out = np.empty((K, N), dtype=np.float32)
    check(lib().b200_spmm_csr_backward_B_host(rp_p, ci_p, va_p, M, K, len(va), G_p, N,
                                              out.ctypes.data_as(_f32p)))
    return out


def transpose_host(row_ptr, col_ind, val, shape):
    M, K = shape
    rp, rp_p = _i32(row_ptr)
    ci, ci_p = _i32(col_ind)
    va, va_p = _f32(val)
    nnz = len(va)
    t_rp = np.empty(K + 1, dtype=np.int32)
    t_ci = np.empty(nnz, dtype=np.int32)
    t_va = np.empty(nnz, dtype=np.float32)
    check(lib().b200_csr_transpose_host(rp_p, ci_p, va_p, M, K, nnz, t_rp.ctypes.data_as(_i32p),
                                        t_ci.ctypes.data_as(_i32p), t_va.ctypes.data_as(_f32p)))
    return t_rp, t_ci, t_va


def dense_to_csr_host(dense):
    d, d_p = _f32(dense)
    assert d.ndim == 2
    M, K = d.shape
    rp = np.empty(M + 1, dtype=np.int32)
    nnz = C.c_int64(0)
    rc = lib().b200_dense_to_csr_host(d_p, M, K, rp.ctypes.data_as(_i32p), None, None, 0, C.byref(nnz))
    if rc != OK and not (rc == ERR_INVALID and nnz.value > 0):
        check(rc)
    n = nnz.value
    ci = np.empty(n, dtype=np.int32)
    va = np.empty(n, dtype=np.float32)
    if n:
        check(lib().b200_dense_to_csr_host(d_p, M, K, rp.ctypes.data_as(_i32p), ci.ctypes.data_as(_i32p),
                                           va.ctypes.data_as(_f32p), n, C.byref(nnz)))
    return rp, ci, va


# ---- host-side partition helpers -----------------------------------------------------------------
def merge_path_boundary(row_ptr, d: int, heavy: int):
    rp, rp_p = _i32(row_ptr)
    M = len(rp) - 1
    r, z = C.c_int64(), C.c_int64()
    check(lib().b200_merge_path_boundary(rp_p, M, int(rp[-1]), d, heavy, C.byref(r), C.byref(z)))
    return r.value, z.value


def shard_rows(row_ptr, parts: int, nnz_weight: float = 1.0, row_weight: float = 0.0) -> np.ndarray:
    rp, rp_p = _i32(row_ptr)
    M = len(rp) - 1
    out = np.zeros(parts + 1, dtype=np.int64)
    check(lib().b200_shard_rows(rp_p, M, parts, nnz_weight, row_weight, out.ctypes.data_as(_i64p)))
    return out


def set_option(name: str, value: int):
    check(lib().b200_set_option(name.encode(), value))


def get_option(name: str) -> int:
    return lib().b200_get_option(name.encode())


def kernel_launch_count() -> int:
    return lib().b200_kernel_launch_count()


def smoke() -> int:
    n = C.c_int(0)
    check(lib().b200_smoke_kernel(None, C.byref(n)))
    return n.value


# ---- device-pointer calls (anything with .data_ptr(): torch CUDA tensors) -------------------------
def _dp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def spmm_device(crow, col, val, shape, B, Cout, ldb=None, ldc=None, workspace=None, stream=0, flags=0):
    """b200_spmm_csr_forward_flags on device buffers; Cout is written in place."""
    M, K = shape
    N = B.shape[1]
    ws_p, ws_n = (None, 0) if workspace is None else (_dp(workspace), workspace.numel() * workspace.element_size())
    check(lib().b200_spmm_csr_forward_flags(_dp(crow), _dp(col), _dp(val), M, K, val.numel(), _dp(B),
                                            N if ldb is None else ldb, N, _dp(Cout),
                                            N if ldc is None else ldc, ws_p, ws_n,
                                            C.c_void_p(stream), flags))
    return Cout


def dense_to_csr_onepass_device(dense, capacity=None, stream=0):
    """b200_dense_to_csr_onepass on a 2-D torch CUDA tensor (unit stride in the last dimension).  Returns
    (crow, col, val, nnz) with col / val holding `capacity` slots (default M*K); raises B200Error (ERR_INVALID, nnz
    in the message, crow complete) when capacity is too small."""
    import torch
    M, K = dense.shape
    cap = M * K if capacity is None else int(capacity)
    crow = torch.empty(M + 1, dtype=torch.int32, device=dense.device)
    col = torch.empty(max(cap, 1), dtype=torch.int32, device=dense.device)
    val = torch.empty(max(cap, 1), dtype=torch.float32, device=dense.device)
    nnz = C.c_int64(0)
    sb = lib().b200_dense_to_csr_onepass_scratch_bytes(M, K)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dense.device)
    rc = lib().b200_dense_to_csr_onepass(_dp(dense), M, K, dense.stride(0) if M > 1 else max(K, 1), _dp(crow), _dp(col),
                                         _dp(val), cap, _dp(scratch), sb, C.byref(nnz), C.c_void_p(stream))
    if rc != OK:
        err = B200Error(rc, lib().b200_last_error().decode())
        err.nnz, err.crow = nnz.value, crow
        raise err
    return crow, col, val, nnz.value


def sddmm_device(crow, col, shape, G, B, out, stream=0):
    M, K = shape
    N = G.shape[1]
    check(lib().b200_sddmm_csr(_dp(crow), _dp(col), M, K, col.numel(), _dp(G), N, _dp(B), N, N, _dp(out),
                               C.c_void_p(stream)))
    return out


class Plan:
    """inspect / execute / clean on device buffers (b200_plan_*); borrows crow/col/val."""

    def __init__(self, crow, col, val, shape, stream=0):
        self._keep = (crow, col, val)
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = val.numel()
        h = C.c_void_p()
        check(lib().b200_plan_create(_dp(crow), _dp(col), _dp(val), self.shape[0], self.shape[1], self.nnz,
                                     C.byref(h), C.c_void_p(stream)))
        self._h = h

    def forward(self, B, Cout, stream=0):
        N = B.shape[1]
        check(lib().b200_plan_forward(self._h, _dp(B), N, N, _dp(Cout), N, C.c_void_p(stream)))
        return Cout

    def forward_ex(self, B, Cout, bias=None, flags=0, stream=0):
        """Extended forward: B is [K, N] (or [N, K] with OP_B_TRANSPOSED), Cout [M, N] (or [N, M] with
        OP_C_TRANSPOSED); bias [M] and ReLU are fused into the epilogue."""
        N = B.shape[0] if flags & OP_B_TRANSPOSED else B.shape[1]
        check(lib().b200_plan_forward_ex(self._h, _dp(B), B.stride(0), N, _dp(Cout), Cout.stride(0), _dp(bias),
                                         flags, C.c_void_p(stream)))
        return Cout

    def backward_B_ex(self, G, dB, flags=0, stream=0):
        N = G.shape[0] if flags & OP_B_TRANSPOSED else G.shape[1]
        check(lib().b200_plan_backward_B_ex(self._h, _dp(G), G.stride(0), N, _dp(dB), dB.stride(0), flags,
                                            C.c_void_p(stream)))
        return dB

    def would_use_tc(self, N, side=0, flags=0):
        return bool(lib().b200_plan_would_use_tc(self._h, side, N, flags))

    def tc_info(self, side=0):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int()
        check(lib().b200_plan_tc_info(self._h, side, C.byref(a), C.byref(b), C.byref(c)))
        return {"tiles": a.value, "bytes": b.value, "last_path": c.value}

    def tc_timeout(self):
        c = C.c_int()
        check(lib().b200_tc_last_timeout(self._h, C.byref(c)))
        return c.value

    def forward_host(self, B_host, C_host, panels=0):
        """End-to-end forward with (pinned) HOST buffers — numpy arrays or CPU torch tensors; uploads, the product
        and downloads are pipelined over column panels inside the library.  Returns when C_host is complete."""
        N = B_host.shape[1]
        pb = B_host.data_ptr() if hasattr(B_host, "data_ptr") else B_host.ctypes.data
        pc = C_host.data_ptr() if hasattr(C_host, "data_ptr") else C_host.ctypes.data
        check(lib().b200_plan_forward_host(self._h, C.c_void_p(pb), N, C.c_void_p(pc), panels))
        return C_host

    def forward_peers(self, B, Cout, peer_ptrs=(), mc_ptr=0, stream=0):
        """Forward whose epilogue also stores every row into peers' buffers (raw device addresses of THIS
        rank's row block inside each peer's C) or to an NVLS multicast address (mc_ptr)."""
        N = B.shape[1]
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*[C.c_void_p(int(x)) for x in peer_ptrs])
        check(lib().b200_plan_forward_peers(self._h, _dp(B), N, N, _dp(Cout), N, arr, len(peer_ptrs),
                                            C.c_void_p(int(mc_ptr)) if mc_ptr else None, C.c_void_p(stream)))
        return Cout

    def backward_B(self, G, dB, stream=0):
        N = G.shape[1]
        check(lib().b200_plan_backward_B(self._h, _dp(G), N, N, _dp(dB), N, C.c_void_p(stream)))
        return dB

    def backward_A(self, G, B, dval, stream=0):
        N = G.shape[1]
        check(lib().b200_plan_backward_A(self._h, _dp(G), N, _dp(B), N, N, _dp(dval), C.c_void_p(stream)))
        return dval

    def backward_A_ex(self, G, B, dval, flags=0, stream=0):
        """grad of the CSR values; with OP_B_TRANSPOSED, G is [N, M] and B is [N, K]."""
        N = G.shape[0] if flags & OP_B_TRANSPOSED else G.shape[1]
        check(lib().b200_plan_backward_A_ex(self._h, _dp(G), G.stride(0), _dp(B), B.stride(0), N, _dp(dval), flags,
                                            C.c_void_p(stream)))
        return dval

    def values_updated(self):
        check(lib().b200_plan_values_updated(self._h))

    def stats(self):
        a, b = C.c_int64(), C.c_int64()
        check(lib().b200_plan_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if self._h is not None:
            lib().b200_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
