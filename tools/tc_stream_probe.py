"""Narrow-N forward on the tensor-core tiles (config 5, 16384^2 @90 % x 64): what limits one CTA's stream of A tiles?
Timing variants through the debug bits of tc_terms, tile width, chunk length; also an 8192-row shard (64 tiles) of it."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi
from matrix_multiplication_b200._capi import OP_FORCE_TC
dev = torch.device("cuda")
K, N = 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 64
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def build(M):
    g = torch.Generator(device=dev).manual_seed(0)
    A = (torch.rand(M, K, device=dev, generator=g) * (torch.rand(M, K, device=dev, generator=g) < 0.1)).float()
    sp = A.to_sparse_csr()
    del A
    return capi.Plan(sp.crow_indices().int().contiguous(), sp.col_indices().int().contiguous(),
                     sp.values().float().contiguous(), (M, K)), torch.empty(M, N, device=dev)


B = torch.rand(K, N, device=dev)


def t(plan, C, iters=10):
    for _ in range(3):
        plan.forward_ex(B, C, flags=OP_FORCE_TC)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.forward_ex(B, C, flags=OP_FORCE_TC); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


for M in (16384, 8192):
    plan, C = build(M)
    for name, terms in (("all", 7), ("hi only", 1), ("no MMA", 0), ("no MMA, no store", 8), ("no MMA, no store, no B copy", 24),
                        ("no MMA, no store, no A copy", 40), ("no MMA, no store, no copies", 56), ("all, no copies", 7 + 48)):
        capi.set_option("tc_terms", terms)
        print(f"M={M} N={N} {name:32s} {t(plan, C):7.1f} us", flush=True)
    capi.set_option("tc_terms", 7)
    for nt in (64, 128):
        for ch in (4, 8, 16):
            capi.set_option("tc_nt", nt); capi.set_option("tc_chunk", ch)
            print(f"M={M} N={N} nt={nt} chunk={ch}: {t(plan, C):7.1f} us", flush=True)
    capi.set_option("tc_nt", 0); capi.set_option("tc_chunk", 8)
    plan.close()
