"""Generates tests/golden/ref_*.npz by running the REAL reference on a GPU box.

    gpurun -- python tests/golden/make_ref_golden.py        (writes gpurun_out/ref_golden/*.npz; copy them here)

`oracle/_ref` holds the reference's own extension (src/ compiled unmodified for sm_100a, see oracle/build_ref.py)
and its byte-compiled matmuls.py.  This script drives them exactly like the reference's test scripts do
(tests/cusparse_kernel_test.py:13-58, tests/naive_kernel_test.py:62-64) with seeded inputs and records
inputs + what the reference returned:

  ref_fwd_<i>    cusparseMM.apply(a_csr, b)            -> custom_mm.cusparse_mmul -> cusparseSpMM (CSR_ALG2)
                 + the CSR arrays get_sparse_tensor_properties produced (torch to_sparse_csr + int32 casts)
  ref_naive_<i>  naiveSpMM.apply(a, b) on dense inputs -> custom_mm.naive_spmm    -> spmm_kernel (the open kernel)
  ref_bwd_0      grad_m2 of cusparseMM (matmuls.py:253: to_sparse_csr of m1^T, then cuSPARSE)

Unlike make_golden.py (which had to stub custom_mm because the build container has no GPU), nothing is
stubbed here: these are outputs of the reference itself.
"""
import random
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import build_ref  # noqa: E402

OUT = ROOT / "gpurun_out" / "ref_golden"


def generate_sparse_dense(shape, n_vals, seed):
    """tests/cusparse_kernel_test.py:13-30, seeded."""
    rnd = random.Random(seed)
    a = torch.zeros(shape)
    seen = {(0, 0)}
    for _ in range(n_vals):
        x, y = 0, 0
        while (x, y) in seen:
            x, y = rnd.randrange(shape[0]), rnd.randrange(shape[1])
        seen.add((x, y))
        a[x, y] = rnd.random()
    return a


def main():
    cm, mm = build_ref.load()
    cm.init_cusparse()
    cm.init_cublas()
    OUT.mkdir(parents=True, exist_ok=True)
    shapes = [((10, 10), (10, 10)), ((10, 20), (20, 10)), ((10, 10), (10, 5)), ((20, 10), (10, 5)),
              ((512, 1024), (1024, 256))]  # tests/cusparse_kernel_test.py:32-38
    for i, (sa, sb) in enumerate(shapes):
        torch.manual_seed(100 + i)
        a = generate_sparse_dense(sa, sa[0] * sa[1] // 10, 100 + i).cuda()
        b = torch.rand(sb).cuda()
        a_csr = a.to_sparse_csr()
        vals, cols, offs, nnz, rows, ncols = mm.get_sparse_tensor_properties(a_csr)
        c = mm.cusparseMM.apply(a_csr, b)
        torch.cuda.synchronize()
        np.savez(OUT / f"ref_fwd_{i}.npz", A_dense=a.cpu().numpy(), B=b.cpu().numpy(), C=c.cpu().numpy(),
                 row_ptr=offs.cpu().numpy(), col=cols.cpu().numpy(), val=vals.cpu().numpy(),
                 shape=np.array([rows, ncols]), nnz=np.array(nnz))
        print("ref_fwd", i, sa, sb, "nnz", nnz, "max|C - A@B| =", float((c - a @ b).abs().max()))
    # naive kernel on dense inputs (tests/naive_kernel_test.py:62-64: (4,2)x(2,3) and the batched 3-D shapes)
    for i, (sa, sb) in enumerate([((4, 2), (2, 3)), ((2, 4, 2), (2, 2, 3)), ((3, 16, 24), (3, 24, 8))]):
        torch.manual_seed(200 + i)
        a, b = torch.rand(sa).cuda(), torch.rand(sb).cuda()
        c = mm.naiveSpMM.apply(a, b)
        torch.cuda.synchronize()
        np.savez(OUT / f"ref_naive_{i}.npz", A_dense=a.cpu().numpy(), B=b.cpu().numpy(), C=c.cpu().numpy())
        print("ref_naive", i, sa, sb, "max|C - A@B| =", float((c - a @ b).abs().max()))
    # backward wrt the dense operand (the well-defined half of cusparseMM.backward)
    torch.manual_seed(300)
    a = generate_sparse_dense((64, 48), 64 * 48 // 10, 300).cuda()
    b = torch.rand(48, 16).cuda().requires_grad_(True)
    g = torch.rand(64, 16).cuda()
    a_csr = a.to_sparse_csr()
    c = mm.cusparseMM.apply(a_csr, b)
    c.backward(g)
    torch.cuda.synchronize()
    vals, cols, offs, nnz, rows, ncols = mm.get_sparse_tensor_properties(a_csr)
    np.savez(OUT / "ref_bwd_0.npz", A_dense=a.cpu().numpy(), B=b.detach().cpu().numpy(), G=g.cpu().numpy(),
             C=c.detach().cpu().numpy(), grad_B=b.grad.cpu().numpy(), row_ptr=offs.cpu().numpy(),
             col=cols.cpu().numpy(), val=vals.cpu().numpy(), shape=np.array([rows, ncols]))
    print("ref_bwd_0 max|grad_B - A^T G| =", float((b.grad - a.t() @ g).abs().max()))
    cm.destroy_cusparse()


if __name__ == "__main__":
    main()
