"""Generates tests/golden/*.npz by running the REFERENCE's own Python layer.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What is real reference code here: `/root/reference/matmuls.py` is imported unmodified, so the
fixtures record what its `get_sparse_tensor_properties` (torch `to_sparse_csr()` + the int32 casts,
matmuls.py:178-187), its `sparse_matmul` / `naive_matmul` shape dispatch (matmuls.py:189-303) and its
`cusparseMM` forward / grad_m2 (matmuls.py:238-256) hand to and get back from `custom_mm`.

What is NOT available here: the container has no GPU, so the compiled `custom_mm` (cuSPARSE) cannot
run.  A stub `custom_mm` stands in for it with the C++ layer's raw-pointer semantics
(src/custom_mm.cpp:203-217: contiguous row-major buffers, strides ignored, C written in place) and
evaluates the SpMM contract in fp64 with scipy — outputs are stored rounded to fp32.  The inputs use
the generators of the reference's tests (tests/cusparse_kernel_test.py:13-30,32-38 and
tests/naive_kernel_test.py:62-64), seeded.
"""
import random
import sys
import types
from pathlib import Path

import numpy as np
import scipy.sparse as sp
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


# ---- stub custom_mm with the C++ layer's semantics ------------------------------------------------
def _raw(t):
    """What data_ptr<T>() sees: the storage read as a contiguous buffer (strides ignored)."""
    return t.detach().cpu().contiguous().numpy() if t.is_contiguous() else \
        np.asarray(t.detach().cpu().untyped_storage().tolist(), dtype=np.uint8).view(
            {torch.float32: np.float32, torch.int32: np.int32}[t.dtype])[: t.numel()]


CALLS = []


def _spmm_stub(values, cols, offs, nnz, rows, ncols, B, C):
    assert values.dtype == torch.float32 and cols.dtype == torch.int32 and offs.dtype == torch.int32
    v, c, o = _raw(values)[:nnz], _raw(cols)[:nnz], _raw(offs)[: rows + 1]
    b_rows, b_cols = B.shape[0], B.shape[1]
    batch = B.numel() // (b_rows * b_cols)
    assert batch == 1
    Bm = _raw(B).reshape(b_rows, b_cols).astype(np.float64)
    A = sp.csr_matrix((v.astype(np.float64), c, o), shape=(rows, ncols))
    out = (A @ Bm).astype(np.float32)
    CALLS.append(dict(values=v.copy(), cols=c.copy(), offs=o.copy(), shape=(rows, ncols)))
    C.copy_(torch.from_numpy(out).reshape(C.shape))
    return C


stub = types.ModuleType("custom_mm")
stub.cusparse_mmul = _spmm_stub
stub.naive_spmm = _spmm_stub
stub.cublas_mmul = lambda a, b, c, ta, tb: c.copy_((a.t() if ta else a) @ (b.t() if tb else b))
stub.cublas_bmm = lambda a, b, c, d, ta, tb: c.copy_(
    (a.transpose(-1, -2) if ta else a) @ (b.transpose(-1, -2) if tb else b))
stub.init_cusparse = stub.destroy_cusparse = lambda: None
sys.modules["custom_mm"] = stub

# the reference calls .cuda() and device=torch.device('cuda'); route both to the CPU
torch.Tensor.cuda = lambda self, *a, **k: self
_real_zeros = torch.zeros
torch.zeros = lambda *a, **k: _real_zeros(*a, **{x: y for x, y in k.items() if x != "device"})
_real_device = torch.device
sys.path.insert(0, str(REF))
import matmuls as ref_matmuls  # noqa: E402  (the unmodified reference module)


# ---- the reference tests' own generators (tests/cusparse_kernel_test.py:13-30) ---------------------
def gen_coords(num_vals, dim1, dim2):
    coords = set()
    while len(coords) < num_vals:
        x, y = 0, 0
        while (x, y) == (0, 0):
            x, y = random.randint(0, dim1 - 1), random.randint(0, dim2 - 1)
        coords.add((x, y))
    return coords


def sparsify(coords, dim1, dim2):
    mat = torch.zeros(dim1, dim2)
    for (x, y) in coords:
        mat[x][y] = random.random()
    return mat


def csr_of(a_csr):
    vals, cols, offs, nnz, M, K = ref_matmuls.get_sparse_tensor_properties(a_csr)
    return dict(val=vals.numpy().copy(), col=cols.numpy().copy(), row_ptr=offs.numpy().copy(),
                shape=np.array([M, K], dtype=np.int64), nnz=np.int64(nnz))


def main():
    random.seed(1234)
    torch.manual_seed(1234)
    np.random.seed(1234)
    n, n_vals = 10, 10
    fixtures = {}

    # 1. forward, the 5 shape pairs of tests/cusparse_kernel_test.py:32-38
    shapes = [((n, n), (n, n)), ((n, 2 * n), (2 * n, n)), ((n, n), (n, n // 2)),
              ((2 * n, n), (n, n // 2)), ((512, 1024), (1024, 256))]
    for i, ((ar, ac), bshape) in enumerate(shapes):
        coords = gen_coords((ar * ac) / n_vals, ar, ac)
        a = sparsify(coords, ar, ac)
        b = torch.rand(bshape)
        a_csr = a.to_sparse_csr()
        out = ref_matmuls.cusparseMM.apply(a_csr, b)
        d = csr_of(a_csr)
        d.update(B=b.numpy().copy(), C=out.numpy().copy())
        if ar * ac <= 4096:
            d["A_dense"] = a.numpy().copy()
        fixtures[f"fwd_{i}"] = d

    # 2. naiveSpMM on fully dense inputs, tests/naive_kernel_test.py:62-64 (small shapes)
    for i, (ashape, bshape, transb) in enumerate([((4, 2), (2, 3), False), ((2, 4, 2), (2, 2, 3), False),
                                                  ((2, 4, 2), (2, 4, 2), True)]):
        a = torch.rand(ashape)
        b = torch.rand(bshape)
        _b = b.transpose(-1, -2).contiguous() if transb else b
        out = ref_matmuls.naiveSpMM.apply(a.contiguous(), _b)
        fixtures[f"naive_{i}"] = dict(A_dense=a.numpy().copy(), B=_b.numpy().copy(), C=out.numpy().copy())

    # 3. CSR construction on special values: what to_sparse_csr + int32 casts keep and drop
    sv = torch.tensor([[0.0, -0.0, 1.0, float("nan"), 0.0],
                       [float("inf"), 0.0, -float("inf"), 1e-42, -1e-45],
                       [0.0, 0.0, 0.0, 0.0, 0.0],
                       [-2.5, 0.0, 0.0, 0.0, 3.0]], dtype=torch.float32)
    d = csr_of(sv.to_sparse_csr())
    d["A_dense"] = sv.numpy().copy()
    fixtures["csr_special"] = d

    # 4. empty rows / ragged rows / one long row, seeded Bernoulli masks
    g = torch.Generator().manual_seed(7)
    for name, (M, K, dens) in {"csr_ragged": (37, 53, 0.15), "csr_sparse": (64, 129, 0.02),
                               "csr_dense": (9, 33, 1.0)}.items():
        a = torch.rand(M, K, generator=g) * (torch.rand(M, K, generator=g) < dens)
        if name == "csr_ragged":
            a[5] = 0
            a[6] = torch.rand(K, generator=g) + 0.5  # one full row
            a[M - 1] = 0
        d = csr_of(a.to_sparse_csr())
        d["A_dense"] = a.numpy().copy()
        fixtures[name] = d
    z = torch.zeros(6, 11)
    d = csr_of(z.to_sparse_csr())
    d["A_dense"] = z.numpy().copy()
    fixtures["csr_allzero"] = d

    # 5. backward: grad_m2 of cusparseMM (matmuls.py:253; grad_m1 is the reference's known-bad path,
    #    SURVEY.md §9 item 1, and is not recorded) and the transposed CSR it builds on the way
    a = torch.rand(24, 40, generator=g) * (torch.rand(24, 40, generator=g) < 0.2)
    b = torch.rand(40, 12, generator=g).requires_grad_(True)
    gout = torch.rand(24, 12, generator=g)
    a_csr = a.to_sparse_csr()
    CALLS.clear()
    out = ref_matmuls.cusparseMM.apply(a_csr, b)
    out.backward(gout)
    t = CALLS[-1]  # the SpMM of backward: CSR of a^T against grad_output
    d = csr_of(a_csr)
    d.update(A_dense=a.numpy().copy(), B=b.detach().numpy().copy(), G=gout.numpy().copy(),
             C=out.detach().numpy().copy(), grad_B=b.grad.numpy().copy(),
             t_row_ptr=t["offs"], t_col=t["cols"], t_val=t["values"])
    # intended grad_m1 (matmuls.py:250) in fp64 on a's pattern, for the SDDMM parity test
    full = gout.double() @ b.detach().double().t()
    r = np.repeat(np.arange(24), np.diff(d["row_ptr"]))
    d["grad_A_val"] = full.numpy()[r, d["col"]].astype(np.float32)
    fixtures["bwd_0"] = d

    for name, d in fixtures.items():
        np.savez_compressed(OUT / f"{name}.npz", **d)
        print(name, {k: getattr(v, "shape", v) for k, v in d.items()})


if __name__ == "__main__":
    main()
