"""Times the block-dense forward (row-major C) and the tensor-core SDDMM on a few BASELINE shapes (CUDA events, L2 flushed):
python tools/tc_epi_probe.py [--opt name=value ...]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi
from matrix_multiplication_b200._capi import OP_FORCE_TC

only = None
for a in sys.argv[1:]:
    if a.startswith("only="):
        only = a[5:]
    elif "=" in a:
        k, v = a.split("=")
        capi.set_option(k, int(v))
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=8):
    fn(); fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


shapes = [("c3 fc1", 3072, 768, 4096, 0.1), ("c3 fc2", 768, 3072, 4096, 0.1), ("c5 @50% x64", 16384, 16384, 64, 0.5),
          ("c5 @90% x64", 16384, 16384, 64, 0.1), ("c5 @90% x256", 16384, 16384, 256, 0.1), ("c5 @50% x16", 16384, 16384, 16, 0.5)]
for name, M, K, N, d in shapes:
    if only is not None and only not in name:
        continue
    g = torch.Generator(device=dev).manual_seed(0)
    A = (torch.randn(M, K, device=dev, generator=g) * (torch.rand(M, K, device=dev, generator=g) < d)).float()
    sp = A.to_sparse_csr()
    crow, col, val = sp.crow_indices().int().contiguous(), sp.col_indices().int().contiguous(), sp.values().float().contiguous()
    del sp
    B = torch.randn(K, N, device=dev, generator=g)
    G = torch.randn(M, N, device=dev, generator=g)
    plan = capi.Plan(crow, col, val, (M, K))
    C = torch.empty(M, N, device=dev)
    dv = torch.empty(val.numel(), device=dev)
    t_f = timed(lambda: plan.forward_ex(B, C, flags=OP_FORCE_TC))
    t_s = timed(lambda: plan.backward_A_ex(G, B, dv, flags=OP_FORCE_TC))
    # parity on sampled rows (fp64)
    rows = torch.randint(0, M, (64,), device=dev, generator=g)
    ref = A[rows].double() @ B.double()
    e_f = float((C[rows].double() - ref).abs().max() / ref.abs().max())
    full = G[rows].double() @ B.double().t()
    e_s = 0.0
    cr = crow.long()
    for i, r in enumerate(rows.tolist()):
        s, e = int(cr[r]), int(cr[r + 1])
        if e > s:
            want = full[i, col[s:e].long()]
            e_s = max(e_s, float((dv[s:e].double() - want).abs().max() / full.abs().max()))
    print(f"{name:14s} nnz {val.numel():10d}  fwd {t_f:.4f} ms  sddmm {t_s:.4f} ms  rel_err fwd {e_f:.2e} sddmm {e_s:.2e}  timeout {plan.tc_timeout()}", flush=True)
    del plan, A, B, G, C, dv
    torch.cuda.empty_cache()
