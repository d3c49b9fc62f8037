import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from matrix_multiplication_b200 import capi, synthetic
z = np.load('/tmp/b200_bench_cache/c4_None.npz', allow_pickle=True) if False else None
rp, ci, va, (M, K) = synthetic.rmat_csr(int(sys.argv[1]) if len(sys.argv) > 1 else 18, 16, seed=0)
N = 128
dev = torch.device('cuda:0')
crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, ci, va))
plan = capi.Plan(crow, col, val, (M, K))
B = torch.rand(K, N).pin_memory(); C = torch.empty(M, N).pin_memory()
ref = None
for panels in (1, 2, 3, 4, 8, 16):
    for _ in range(2): plan.forward_host(B, C, panels)
    if ref is None: ref = C.clone()
    else: assert torch.equal(ref, C), f"panels={panels}: result differs from the single-panel run"
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): plan.forward_host(B, C, panels)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5 * 1e3
    print(f"panels {panels:2d}: {dt:.3f} ms   ({2 * K * N * 4 / dt / 1e6:.1f} GB/s both directions summed)")
# raw copies
Bd = torch.empty(K, N, device=dev)
for name, fn in (("H2D contiguous", lambda: Bd.copy_(B, non_blocking=True)), ("D2H contiguous", lambda: C.copy_(Bd, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5 * 1e3
    print(f"{name}: {dt:.3f} ms  {K * N * 4 / dt / 1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): Bd.copy_(B, non_blocking=True)
    with torch.cuda.stream(s2): C.copy_(Bd, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5 * 1e3
print(f"H2D || D2H on two streams: {dt:.3f} ms")
