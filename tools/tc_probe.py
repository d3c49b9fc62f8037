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
        capi.set_option(k, {"tc_terms": 7, "tc_split_acc": 0, "tc_nt": 0, "tc_mode": 1}.get(k, 0))
    plan.close()
    out(**rec)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    out(device=torch.cuda.get_device_name(0), sms=capi.smoke())

    # 1. layout: A = first 32 columns identity, B small exact values -> C[r] = B[r] for r < 32, 0 elsewhere
    A = torch.zeros(128, 32, device=dev)
    A[torch.arange(32), torch.arange(32)] = 1.0
    B = (torch.arange(32, device=dev)[:, None] + torch.arange(32, device=dev)[None, :] / 64.0).float()
    run_case("layout_identity", A, B, opts={"tc_terms": 1})
    run_case("layout_identity_all_terms", A, B)
    # permutation-like A exercising every (row, k) swizzle position
    A = torch.zeros(128, 64, device=dev)
    r = torch.arange(128, device=dev)
    A[r, (r * 7 + 3) % 64] = ((r % 5) + 1).float()
    B = (torch.arange(64, device=dev)[:, None] * 2 + torch.arange(64, device=dev)[None, :] / 32.0).float()
    run_case("layout_perm", A, B, opts={"tc_terms": 1})

    # 2. accuracy of the split
    A = rand_sparse(128, 64, 0.5)
    B = torch.randn(64, 64, device=dev)
    run_case("rand_hi_only", A, B, opts={"tc_terms": 1})
    run_case("rand_3x", A, B)
    run_case("rand_3x_split", A, B, opts={"tc_split_acc": 1})
    # 3. ragged edges, several row blocks / k tiles / n blocks, every tile width
    A = rand_sparse(300, 200, 0.3, seed=1)
    B = torch.randn(200, 100, device=dev)
    for nt in (32, 64, 96, 128, 160, 192, 224, 256):
        run_case(f"ragged_nt{nt}", A, B, opts={"tc_nt": nt})
    A = rand_sparse(1000, 1000, 0.1, seed=2)
    B = torch.randn(1000, 777, device=dev)
    run_case("ragged_1000", A, B)
    bias = torch.randn(1000, device=dev)
    run_case("bias_relu", A, B, flags=OP_FORCE_TC | OP_RELU, bias=bias)
    run_case("bt", A, B, bt=True)
    run_case("ct", A, B, ct=True)
    run_case("bt_ct_bias_relu", A, B, flags=OP_FORCE_TC | OP_RELU, bias=bias, bt=True, ct=True)
    # empty row blocks and empty k tiles
    A2 = A.clone()
    A2[128:384] = 0
    A2[:, 64:256] = 0
    run_case("empty_blocks", A2, B)
    # non-finite dense operand -> CSR-order fallback on the device
    B2 = B.clone()
    B2[5, 7] = float("inf")
    B2[9, 3] = float("nan")
    ref = torch.zeros(1000, 777, device=dev, dtype=torch.float64)
    Ad = A.double()
    Bd = B2.double()
    Bd0 = torch.where(torch.isfinite(Bd), Bd, torch.zeros_like(Bd))
    ref = Ad @ Bd0
    rec = run_case("nonfinite_fallback", A, B2, ref64=ref)

    # 4. accumulate rounding: long K, positive data
    A = rand_sparse(256, 16384, 0.5, signed=False, seed=3)
    B = torch.rand(16384, 256, device=dev)
    ref = A.double() @ B.double()
    run_case("longk_pos", A, B, ref64=ref)
    run_case("longk_pos_split", A, B, opts={"tc_split_acc": 1, "tc_nt": 128}, ref64=ref)
    run_case("longk_pos_hi", A, B, opts={"tc_terms": 1}, ref64=ref)
    run_case("longk_pos_gather", A, B, flags=OP_NO_TC, ref64=ref)
    out(case="longk_pos_torch_fp32", rel_err=rel_err(A @ B, ref))
    if args.quick:
        return

    # 5. config 3 (BERT FC layers, 90 % sparse weights, 4096 tokens)
    for (M, K) in ((3072, 768), (768, 3072)):
        A = rand_sparse(M, K, 0.1, signed=False, seed=4)
        B = torch.rand(K, 4096, device=dev)
        ref = A.double() @ B.double()
        run_case(f"c3_{M}x{K}_gather", A, B, flags=OP_NO_TC, timing=True, ref64=ref)
        for nt in (0, 64, 128, 192, 256):
            run_case(f"c3_{M}x{K}_nt{nt}", A, B, opts={"tc_nt": nt}, timing=True, ref64=ref)
        run_case(f"c3_{M}x{K}_split128", A, B, opts={"tc_nt": 128, "tc_split_acc": 1}, timing=True, ref64=ref)
        run_case(f"c3_{M}x{K}_hi_only", A, B, opts={"tc_terms": 1}, timing=True, ref64=ref)
        run_case(f"c3_{M}x{K}_bt_ct", A, B, timing=True, ref64=ref, bt=True, ct=True)
        Ad = A
        torch.backends.cuda.matmul.allow_tf32 = False
        out(case=f"c3_{M}x{K}_torch_fp32", ms=time_ms(lambda: torch.mm(Ad, B)))
        torch.backends.cuda.matmul.allow_tf32 = True
        out(case=f"c3_{M}x{K}_torch_tf32", ms=time_ms(lambda: torch.mm(Ad, B)), rel_err=rel_err(torch.mm(Ad, B), ref))
        torch.backends.cuda.matmul.allow_tf32 = False

    # 6. config 5 points
    if args.big:
        for sparsity in (0.9, 0.5):
            A = rand_sparse(16384, 16384, 1 - sparsity, signed=False, seed=5)
            for N in (64, 256):
                B = torch.rand(16384, N, device=dev)
                ref = None
                # row-sampled fp64 reference (full fp64 product of 16k^2 x 256 is fine on the GPU too)
                ref = A.double() @ B.double()
                run_case(f"c5_{sparsity}_{N}_gather", A, B, flags=OP_NO_TC, timing=True, ref64=ref)
                run_case(f"c5_{sparsity}_{N}_tc", A, B, timing=True, ref64=ref)
                if N == 256:
                    run_case(f"c5_{sparsity}_{N}_tc_split", A, B, opts={"tc_nt": 128, "tc_split_acc": 1}, timing=True,
                             ref64=ref)
                torch.backends.cuda.matmul.allow_tf32 = False
                out(case=f"c5_{sparsity}_{N}_torch_fp32", ms=time_ms(lambda: torch.mm(A, B), iters=5))
            del A
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
