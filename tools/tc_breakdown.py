"""Where does the tensor-core kernel's time go?  Config 3 fc1, kernel-only timing (packing excluded is not possible through the
API, so the pack time is measured separately with terms that skip everything else)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi
from matrix_multiplication_b200._capi import OP_FORCE_TC
dev = torch.device("cuda")
M, K, N = 3072, 768, 4096
g = torch.Generator(device=dev).manual_seed(0)
A = (torch.rand(M, K, device=dev, generator=g) * (torch.rand(M, K, device=dev, generator=g) < 0.1)).float()
B = torch.rand(K, N, device=dev, generator=g)
sp = A.to_sparse_csr()
crow, col, val = sp.crow_indices().int().contiguous(), sp.col_indices().int().contiguous(), sp.values().float().contiguous()
plan = capi.Plan(crow, col, val, (M, K))
C = torch.empty(M, N, device=dev)
def t(iters=30):
    for _ in range(3): plan.forward_ex(B, C, flags=OP_FORCE_TC)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): plan.forward_ex(B, C, flags=OP_FORCE_TC)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for pair in (0, 1):
    capi.set_option("tc_pair", pair)
    for name, terms in (("all", 7), ("hi only", 1), ("no MMA", 0), ("no MMA, no store", 8), ("no MMA, no store, no B copy", 24),
                        ("no MMA, no store, no copies", 56), ("all, no store", 15), ("all, no copies", 7 + 48)):
        capi.set_option("tc_terms", terms)
        print(f"pair={pair} {name:32s} {t():7.1f} us", flush=True)
capi.set_option("tc_terms", 7)
for pair in (0, 1):
    capi.set_option("tc_pair", pair)
    for nt in (128, 256):
        for ch in (2, 4, 8, 12):
            capi.set_option("tc_nt", nt); capi.set_option("tc_chunk", ch)
            print(f"pair={pair} nt={nt} chunk={ch}: {t():7.1f} us", flush=True)
