"""Narrow-N forward on config 5 (16384^2 x 16): shared-memory slab kernel vs lane-group gather kernel vs the tiles."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi, synthetic
from oracle import oracle
dev = torch.device("cuda:0")
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for sp in (0.9, 0.99):
    rp, ci, va = synthetic.bernoulli_csr(16384, 16384, sp, seed=0)
    crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, ci, va))
    for N in (16, 8):
        B = torch.rand(16384, N, device=dev)
        C = torch.empty(16384, N, device=dev)
        plan = capi.Plan(crow, col, val, (16384, 16384))
        res = {}
        for name, opts, flags in (("slab2", {"slab_mode": 3}, capi.OP_NO_TC), ("slab1", {"slab_mode": 2}, capi.OP_NO_TC),
                                  ("gather", {"slab_mode": 0}, capi.OP_NO_TC),
                                  ("tiles", {"slab_mode": 0}, capi.OP_FORCE_TC), ("auto", {"slab_mode": 1}, 0)):
            for k, v in opts.items(): capi.set_option(k, v)
            plan.forward_ex(B, C, flags=flags); torch.cuda.synchronize()
            path = plan.tc_info()["last_path"]
            if name == "slab2":
                ref = oracle.spmm_f64(rp, ci, va, (16384, 16384), B.cpu().numpy()) if sp >= 0.9 else None
                if ref is not None:
                    print(f"   slab2 rel err vs fp64 oracle: {oracle.rel_err(C.cpu().numpy(), ref):.2e}", flush=True)
            res[name] = (round(timeit(lambda: plan.forward_ex(B, C, flags=flags)), 4), path)
        capi.set_option("slab_mode", 1)
        by = synthetic.spmm_algorithmic_bytes(16384, 16384, len(va), N)
        best = res["auto"][0]
        print(f"sparsity {sp} N {N}: {res}  auto = {by / best / 1e6 / 6549.1:.3f} of roof", flush=True)
        plan.close()
