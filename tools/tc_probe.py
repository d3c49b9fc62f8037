"""GPU probe of the block-dense tensor-core path (tc_gemm.cu): layout checks, accuracy of the split-precision
products against fp64, accumulate-rounding study, and timings against the gather kernels and torch.mm.

    python tools/tc_probe.py [--quick] > gpurun_out/tc_probe.txt
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi  # noqa: E402
from matrix_multiplication_b200._capi import (OP_B_TRANSPOSED, OP_C_TRANSPOSED, OP_FORCE_TC, OP_NO_TC,  # noqa: E402
                                              OP_RELU)

dev = torch.device("cuda")


def out(**kw):
    print(json.dumps(kw), flush=True)


def make_csr(dense):
    sp = dense.to_sparse_csr()
    return (sp.crow_indices().to(torch.int32).contiguous(), sp.col_indices().to(torch.int32).contiguous(),
            sp.values().to(torch.float32).contiguous())


def rand_sparse(M, K, density, signed=True, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    mask = torch.rand(M, K, device=dev, generator=g) < density
    v = torch.randn(M, K, device=dev, generator=g) if signed else torch.rand(M, K, device=dev, generator=g)
    v = torch.where(v == 0, torch.ones_like(v), v)
    return (v * mask).to(torch.float32)


def rel_err(x, ref64):
    return float((x.double() - ref64).abs().max() / ref64.abs().max().clamp_min(1e-300))


def time_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        e0[i].record()
        fn()
        e1[i].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
    return ts[len(ts) // 2]


def run_case(name, A, B, flags=OP_FORCE_TC, bias=None, opts=None, timing=False, ref64=None, bt=False, ct=False):
    M, K = A.shape
    N = B.shape[1]
    crow, col, val = make_csr(A)
    plan = capi.Plan(crow, col, val, (M, K))
    for k, v in (opts or {}).items():
        capi.set_option(k, v)
    Bin = B.t().contiguous() if bt else B
    Cout = torch.full((N, M) if ct else (M, N), float("nan"), device=dev, dtype=torch.float32)
    f = flags | (OP_B_TRANSPOSED if bt else 0) | (OP_C_TRANSPOSED if ct else 0)
    rec = {"case": name, "M": M, "K": K, "N": N, "nnz": int(val.numel()), "opts": opts or {}, "flags": f}
    try:
        plan.forward_ex(Bin, Cout, bias=bias, flags=f)
        torch.cuda.synchronize()
        rec["timeout_code"] = plan.tc_timeout()
        rec["info"] = plan.tc_info()
        C = Cout.t() if ct else Cout
        if ref64 is None:
            ref64 = A.double() @ B.double()
            if bias is not None:
                ref64 = ref64 + bias.double()[:, None]
            if f & OP_RELU:
                ref64 = ref64.clamp_min(0)
        rec["rel_err"] = rel_err(C, ref64)
        rec["nan"] = int(torch.isnan(C).sum())
        if rec["rel_err"] > 1e-4 or rec["nan"]:
            bad = ((C.double() - ref64).abs() > 1e-4 * ref64.abs().max()) | torch.isnan(C)
            rows = bad.any(1).nonzero().flatten()[:8].tolist()
            cols = bad.any(0).nonzero().flatten()[:8].tolist()
            rec["bad_count"] = int(bad.sum())
            rec["bad_rows"] = rows
            rec["bad_cols"] = cols
            rec["sample_got"] = C[:2, :6].tolist()
            rec["sample_ref"] = ref64[:2, :6].tolist()
        if timing:
            rec["ms"] = time_ms(lambda: plan.forward_ex(Bin, Cout, bias=bias, flags=f))
            rec["dense_tflops_x3"] = 3 * 2.0 * M * K * N / rec["ms"] / 1e9
            rec["sparse_gflops"] = 2.0 * val.numel() * N / rec["ms"] / 1e6
    except Exception as e:  # noqa: BLE001
        rec["error"] = repr(e)
    for k in (opts or {}):
        capi.set_option(k, {"tc_terms": 7, "tc_chunk": 8, "tc_nt": 0, "tc_mode": 1, "tc_kind": 1}.get(k, 0))
    plan.close()
    out(**rec)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    out(device=torch.cuda.get_device_name(0), sms=capi.smoke())

    kinds = (0, 1)
    # 1. layout: A = first 32 columns identity, B small exact values -> C[r] = B[r] for r < 32, 0 elsewhere
    A = torch.zeros(128, 32, device=dev)
    A[torch.arange(32), torch.arange(32)] = 1.0
    B = (torch.arange(32, device=dev)[:, None] + torch.arange(32, device=dev)[None, :] / 64.0).float()
    for kind in kinds:
        run_case(f"layout_identity_k{kind}", A, B, opts={"tc_terms": 1, "tc_kind": kind})
    # permutation-like A exercising every (row, k) swizzle position
    A = torch.zeros(128, 128, device=dev)
    r = torch.arange(128, device=dev)
    A[r, (r * 7 + 3) % 128] = ((r % 5) + 1).float()
    B = (torch.arange(128, device=dev)[:, None] * 2 + torch.arange(64, device=dev)[None, :] / 32.0).float()
    for kind in kinds:
        run_case(f"layout_perm_k{kind}", A, B, opts={"tc_kind": kind})

    # 2. accuracy of the split
    A = rand_sparse(128, 64, 0.5)
    B = torch.randn(64, 64, device=dev)
    for kind in kinds:
        run_case(f"rand_hi_only_k{kind}", A, B, opts={"tc_terms": 1, "tc_kind": kind})
        run_case(f"rand_3x_k{kind}", A, B, opts={"tc_kind": kind})
    # 3. ragged edges, several row blocks / k tiles / n blocks, every tile width
    A = rand_sparse(300, 200, 0.3, seed=1)
    B = torch.randn(200, 100, device=dev)
    for kind in kinds:
        for nt in (64, 128, 192, 256):
            run_case(f"ragged_nt{nt}_k{kind}", A, B, opts={"tc_nt": nt, "tc_kind": kind})
    A = rand_sparse(1000, 1000, 0.1, seed=2)
    B = torch.randn(1000, 777, device=dev)
    bias = torch.randn(1000, device=dev)
    A2 = A.clone()
    A2[128:384] = 0
    A2[:, 64:256] = 0
    B2 = B.clone()
    B2[5, 7] = float("inf")
    B2[9, 3] = float("nan")
    # wide dynamic range: rows of A and columns of B spanning 2^+-40, and a tiny-but-relevant element pairing
    rsc = torch.exp2(torch.linspace(-40, 40, 1000, device=dev)).float()
    csc = torch.exp2(torch.linspace(30, -30, 777, device=dev)).float()
    A3 = A * rsc[:, None]
    B3 = B * csc[None, :]
    A4 = A.clone()
    B4 = B.clone()
    A4[:, 17] = A4[:, 17] * 1e6
    B4[17, :] = B4[17, :] * 1e-6
    for kind in kinds:
        o = {"tc_kind": kind}
        run_case(f"ragged_1000_k{kind}", A, B, opts=o)
        run_case(f"bias_relu_k{kind}", A, B, flags=OP_FORCE_TC | OP_RELU, bias=bias, opts=o)
        run_case(f"bt_k{kind}", A, B, bt=True, opts=o)
        run_case(f"ct_k{kind}", A, B, ct=True, opts=o)
        run_case(f"bt_ct_bias_relu_k{kind}", A, B, flags=OP_FORCE_TC | OP_RELU, bias=bias, bt=True, ct=True, opts=o)
        run_case(f"empty_blocks_k{kind}", A2, B, opts=o)
        run_case(f"range_rows_cols_k{kind}", A3, B3, opts=o)
        run_case(f"range_pairing_k{kind}", A4, B4, opts=o)
        # non-finite dense operand -> CSR-order fallback on the device: must equal the gather kernels' result
        crow, col, val = make_csr(A)
        pl = capi.Plan(crow, col, val, A.shape)
        Cg = torch.empty(1000, 777, device=dev)
        pl.forward_ex(B2, Cg, flags=OP_NO_TC)
        capi.set_option("tc_kind", kind)
        Ct = torch.empty(1000, 777, device=dev)
        pl.forward_ex(B2, Ct, flags=OP_FORCE_TC)
        torch.cuda.synchronize()
        capi.set_option("tc_kind", 1)
        same_nan = bool((torch.isnan(Cg) == torch.isnan(Ct)).all())
        fin = torch.isfinite(Cg)
        out(case=f"nonfinite_fallback_k{kind}", same_nan_pattern=same_nan, nan=int(torch.isnan(Ct).sum()),
            inf=int(torch.isinf(Ct).sum()), max_diff_finite=float((Cg[fin] - Ct[fin]).abs().max()))
        pl.close()

    # 4. accumulate rounding: long K, positive data
    A = rand_sparse(256, 16384, 0.5, signed=False, seed=3)
    B = torch.rand(16384, 256, device=dev)
    ref = A.double() @ B.double()
    for kind in kinds:
        for ch in (1, 2, 4, 8, 16, 1000):
            run_case(f"longk_pos_k{kind}_chunk{ch}", A, B, opts={"tc_chunk": ch, "tc_kind": kind}, ref64=ref, timing=True)
    run_case("longk_pos_gather", A, B, flags=OP_NO_TC, ref64=ref, timing=True)
    out(case="longk_pos_torch_fp32", rel_err=rel_err(A @ B, ref))
    As = rand_sparse(256, 16384, 0.5, signed=True, seed=3)
    Bs = torch.randn(16384, 256, device=dev)
    ref = As.double() @ Bs.double()
    for kind in kinds:
        for ch in (4, 1000):
            run_case(f"longk_signed_k{kind}_chunk{ch}", As, Bs, opts={"tc_chunk": ch, "tc_kind": kind}, ref64=ref)
    if args.quick:
        return

    # 5. config 3 (BERT FC layers, 90 % sparse weights, 4096 tokens)
    for (M, K) in ((3072, 768), (768, 3072)):
        A = rand_sparse(M, K, 0.1, signed=False, seed=4)
        B = torch.rand(K, 4096, device=dev)
        ref = A.double() @ B.double()
        run_case(f"c3_{M}x{K}_gather", A, B, flags=OP_NO_TC, timing=True, ref64=ref)
        for kind in kinds:
            for nt in (0, 128, 192, 256):
                run_case(f"c3_{M}x{K}_k{kind}_nt{nt}", A, B, opts={"tc_nt": nt, "tc_kind": kind}, timing=True, ref64=ref)
            for ch in (2, 8):
                run_case(f"c3_{M}x{K}_k{kind}_chunk{ch}", A, B, opts={"tc_chunk": ch, "tc_kind": kind}, timing=True, ref64=ref)
            run_case(f"c3_{M}x{K}_k{kind}_hi_only", A, B, opts={"tc_terms": 1, "tc_kind": kind}, timing=True, ref64=ref)
            run_case(f"c3_{M}x{K}_k{kind}_bt_ct", A, B, opts={"tc_kind": kind}, timing=True, ref64=ref, bt=True, ct=True)
        torch.backends.cuda.matmul.allow_tf32 = True
        out(case=f"c3_{M}x{K}_torch_tf32", ms=time_ms(lambda: torch.mm(A, B)), rel_err=rel_err(torch.mm(A, B), ref))
        torch.backends.cuda.matmul.allow_tf32 = False
        Ah, Bh = A.half(), B.half()
        out(case=f"c3_{M}x{K}_torch_fp16", ms=time_ms(lambda: torch.mm(Ah, Bh)))

    # 6. config 5 points
    if args.big:
        for sparsity in (0.9, 0.5):
            A = rand_sparse(16384, 16384, 1 - sparsity, signed=False, seed=5)
            for N in (64, 256):
                B = torch.rand(16384, N, device=dev)
                ref = A.double() @ B.double()
                run_case(f"c5_{sparsity}_{N}_gather", A, B, flags=OP_NO_TC, timing=True, ref64=ref)
                for kind in kinds:
                    run_case(f"c5_{sparsity}_{N}_tc_k{kind}", A, B, opts={"tc_kind": kind}, timing=True, ref64=ref)
                torch.backends.cuda.matmul.allow_tf32 = False
                out(case=f"c5_{sparsity}_{N}_torch_fp32", ms=time_ms(lambda: torch.mm(A, B), iters=5))
            del A
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
