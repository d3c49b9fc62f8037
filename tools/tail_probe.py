"""Effect of the short-task tail on the forward: whole config 4 and its 8 cost-balanced shards (one GPU, each shard run by
itself), for several tail_tasks / items_per_task settings."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi, sharding, synthetic

rp, ci, va, (M, K) = synthetic.rmat_csr(20, 16, seed=0)
N = 128
dev = torch.device("cuda:0")
B = torch.rand(K, N, device=dev)
C = torch.empty(M, N, device=dev)


def time_plan(crow, col, val, shape, out):
    plan = capi.Plan(crow, col, val, shape)
    for _ in range(3):
        plan.forward(B, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        plan.forward(B, out)
    e1.record(); torch.cuda.synchronize()
    plan.close()
    return e0.elapsed_time(e1) / 20


full = tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (rp, ci, va))
bounds = sharding.shard_bounds(rp, 8, N)
shards = []
for g in range(8):
    l_rp, l_ci, l_va, (r0, r1) = sharding.shard_csr(rp, ci, va, bounds, g)
    shards.append((tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (l_rp, l_ci, l_va)), (r0, r1)))
for tail in (0, 2368, 4736, 9472, 18944):
    for ipt in (0, 160, 256):
        capi.set_option("tail_tasks", tail)
        capi.set_option("items_per_task", ipt)
        t_full = time_plan(*full, (M, K), C)
        ts = [time_plan(*s, (r1 - r0, K), C[r0:r1]) for s, (r0, r1) in shards]
        print(f"tail_tasks {tail:6d} items_per_task {ipt:3d}: full {t_full:.4f} ms   shards max {max(ts):.4f} mean {np.mean(ts):.4f}"
              f"  -> {t_full / max(ts):.2f}x of this full / {0.5327 / max(ts):.2f}x of 0.5327   {[round(t, 4) for t in ts]}", flush=True)
