"""e2e host pipeline variants on config 4: copy engine vs SM copy kernel, panel counts."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi, synthetic
rp, ci, va, (M, K) = synthetic.rmat_csr(20, 16, seed=0)
N = 128
dev = torch.device("cuda:0")
crow, col, val = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (rp, ci, va))
plan = capi.Plan(crow, col, val, (M, K))
B = torch.rand(K, N).pin_memory()
C = torch.empty(M, N).pin_memory()
ref = None
for split in (1, 2, 3, 4):
    for panels in (2, 4, 8):
        capi.set_option("host_copy_mode", 0); capi.set_option("host_copy_split", split)
        for _ in range(2): plan.forward_host(B, C, panels)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): plan.forward_host(B, C, panels)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
        if ref is None: ref = C.clone()
        print(f"split {split} panels {panels}: {ms:.3f} ms  same={bool(torch.equal(ref, C))}", flush=True)
