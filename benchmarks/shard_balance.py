#!/usr/bin/env python
"""Per-shard forward time of the row-sharded config-4 product, measured on ONE GPU (each shard's plan run by
itself with B resident): the data behind the cost weights of sharding.shard_bounds.  Fits t = a*nnz + b*rows + c."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from matrix_multiplication_b200 import capi, sharding, synthetic  # noqa: E402


def main():
    parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    miss = float(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None
    if len(sys.argv) > 3:
        capi.set_option("items_per_task", int(sys.argv[3]))
    rp, ci, va, (M, K) = synthetic.rmat_csr(20, 16, seed=0)
    N = 128
    dev = torch.device("cuda:0")
    B = torch.rand(K, N, device=dev)
    C = torch.empty(M, N, device=dev)
    rows_all = []
    for name, bounds in (("cost", sharding.shard_bounds(rp, parts, N) if miss is None else sharding.shard_bounds(rp, parts, N, miss)),
                         ("nnz", capi.shard_rows(rp, parts, 1.0, 0.0)), ("rows", capi.shard_rows(rp, parts, 0.0, 1.0))):
        ts = []
        only = int(sys.argv[4]) if len(sys.argv) > 4 else None
        for g in range(parts):
            if only is not None and (g != only or name != "cost"):
                continue
            l_rp, l_ci, l_va, (r0, r1) = sharding.shard_csr(rp, ci, va, bounds, g)
            crow, col, val = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (l_rp, l_ci, l_va))
            plan = capi.Plan(crow, col, val, (r1 - r0, K))
            out = C[r0:r1]
            for _ in range(3):
                plan.forward(B, out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                plan.forward(B, out)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 10)
            rows_all.append((len(l_va), r1 - r0, ts[-1]))
            plan.close()
        if not ts:
            continue
        print(name, "bounds", [int(b) for b in bounds], "ms", [round(t, 4) for t in ts], "max", round(max(ts), 4), flush=True)
    if len(rows_all) < 4:
        return
    A = np.array([[n, r, 1.0] for n, r, _ in rows_all]); y = np.array([t for _, _, t in rows_all])
    coef, *_ = np.linalg.lstsq(A, y, rcond=None)
    print("fit: ms = %.3e * nnz + %.3e * rows + %.4f   (ratio row/nnz cost = %.2f)" % (coef[0], coef[1], coef[2], coef[1] / coef[0]))


if __name__ == "__main__":
    main()
