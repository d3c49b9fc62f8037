"""Tensor-core SDDMM, short contractions: the 16-epilogue-warp kernel (tc_sddmm_short=1) against the general kernel
(=0) and an fp64 reference:   python tools/sddmm_short_probe.py [quick]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi
from matrix_multiplication_b200._capi import OP_FORCE_TC

dev = torch.device("cuda")
MODES = (2, 3, 1, 0)  # 128-row tiles, 256-row tiles, auto, general kernel
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def case(M, K, N, density, kind=1, time_it=True):
    g = torch.Generator(device=dev).manual_seed(M * 7 + K * 3 + N)
    mask = torch.rand(M, K, device=dev, generator=g) < density
    A = (torch.rand(M, K, device=dev, generator=g) + 0.5) * mask
    sp = A.to_sparse_csr()
    crow, col, val = sp.crow_indices().int().contiguous(), sp.col_indices().int().contiguous(), sp.values().float().contiguous()
    del A, mask
    B = torch.rand(K, N, device=dev, generator=g) - 0.5
    G = torch.rand(M, N, device=dev, generator=g) - 0.5
    plan = capi.Plan(crow, col, val, (M, K))
    capi.set_option("tc_kind", kind)
    rows = torch.repeat_interleave(torch.arange(M, device=dev), crow.diff().long())
    ref = (G.double()[rows] * B.double()[col.long()]).sum(1) if val.numel() * N < 4e9 else None
    out = {}
    for short in MODES:
        capi.set_option("tc_sddmm_short", short)
        dv = torch.full((val.numel(),), float("nan"), device=dev)
        plan.backward_A_ex(G, B, dv, flags=OP_FORCE_TC)
        torch.cuda.synchronize()
        lk = capi.get_option("tc_last_kernel")
        err = float((dv.double() - ref).abs().max() / ref.abs().max()) if ref is not None else -1.0
        ms = timed(lambda: plan.backward_A_ex(G, B, dv, flags=OP_FORCE_TC)) if time_it else 0.0
        out[short] = (lk, err, ms, plan.tc_timeout(), dv.clone())
    msg = f"M {M} K {K} N {N} density {density} kind {kind}:"
    for short in MODES:
        same = float((out[short][4].double() - out[0][4].double()).abs().max() / out[0][4].double().abs().max())
        msg += f" | mode {short}: kernel {out[short][0]} err {out[short][1]:.2e} vs general {same:.1e} {out[short][2]:.4f} ms timeout {out[short][3]}"
    print(msg, flush=True)
    capi.set_option("tc_sddmm_short", 1)
    plan.close()


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for nt in (128, 256):  # (the 16-warp kernel needs a packing width that is a multiple of its 128-column tiles)
    capi.set_option("tc_nt", nt)
    for shape in [(300, 500, 40, 0.2), (128, 128, 8, 0.5), (1000, 130, 64, 0.1), (4096, 4096, 256, 0.1),
                  (777, 3001, 100, 0.05), (1, 33, 512, 1.0), (5000, 31, 9, 0.7)]:
        case(*shape, time_it=False)
    case(777, 3001, 100, 0.05, kind=0, time_it=False)
capi.set_option("tc_nt", 0)
if not quick:
    for N in (16, 64, 128, 256):
        for d in (0.1, 0.5):
            case(16384, 16384, N, d)
    case(16384, 16384, 64, 0.1, kind=0)
