#!/usr/bin/env python
"""All BASELINE.json configs on one B200: forward SpMM, grad_B (SpMM over the cached transposed CSR),
grad_A (SDDMM) and dense->CSR of the engine, next to
  * cuSPARSE, kernel-fair: cusparseSpMM CSR_ALG2 with descriptors / buffer / preprocess hoisted (tools/cusparse_fair.cu),
  * cuSPARSE as torch calls it (`torch.sparse.mm`, descriptors per call) and as the reference build calls it
    (`cusparse_mmul` of oracle/_ref, per-call cudaMalloc), when present,
  * the dense crossover: cuBLAS SGEMM (fp32 SIMT), TF32 and BF16 tensor-core GEMMs on the densified A.
Cold = L2 flushed before every timed launch; warm = back-to-back launches (operands L2-resident where they fit).

    python benchmarks/run_configs.py [--configs c1,c2,c3,c4,c5] [--out profiles/r1_configs.json]

Timing: CUDA events on the current stream, 3 warm-ups, L2 flushed (256 MB fill) before every timed launch,
median of `--iters` launches.  Bytes are the algorithmic bytes of SURVEY.md §8(d)."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from matrix_multiplication_b200 import capi, synthetic  # noqa: E402
from tools import cusparse_fair  # noqa: E402

PEAK = 6549.1
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def timeit_warm(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def configs(which):
    out = []
    if "c1" in which:
        out.append(("c1 1024^2 @90% x64", lambda: synthetic.bernoulli_csr(1024, 1024, 0.9, 0), (1024, 1024), 64))
    if "c2" in which:
        out.append(("c2 fc1 512x784 @90% x256", lambda: synthetic.bernoulli_csr(512, 784, 0.9, 0), (512, 784), 256))
        out.append(("c2 fc2 10x512 @90% x256", lambda: synthetic.bernoulli_csr(10, 512, 0.9, 1), (10, 512), 256))
    if "c3" in which:
        out.append(("c3 fc1 3072x768 @90% x4096", lambda: synthetic.bernoulli_csr(3072, 768, 0.9, 0), (3072, 768), 4096))
        out.append(("c3 fc2 768x3072 @90% x4096", lambda: synthetic.bernoulli_csr(768, 3072, 0.9, 1), (768, 3072), 4096))
    if "c4" in which:
        def rmat():
            rp, ci, va, _ = synthetic.rmat_csr(20, 16, seed=0)
            return rp, ci, va
        out.append(("c4 R-MAT 2^20 nnz16M x128", rmat, (1 << 20, 1 << 20), 128))
    if "c5" in which:
        for sp in (0.5, 0.9, 0.99, 0.999):
            for N in (16, 64, 256):
                out.append((f"c5 16384^2 @{sp * 100:g}% x{N}",
                            (lambda sp=sp: synthetic.bernoulli_csr(16384, 16384, sp, 0)), (16384, 16384), N))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "configs.json"))
    ap.add_argument("--no-dense", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    for o in args.opt:
        capi.set_option(o.split("=")[0], int(o.split("=")[1]))
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush = lambda: flush_buf.fill_(1)  # noqa: E731
    ref_cm = None
    try:
        from oracle import build_ref
        if build_ref.available():
            ref_cm, _ = build_ref.load()
            ref_cm.init_cusparse()
    except Exception as e:  # comparator only
        print("reference build not loadable:", e, file=sys.stderr)

    rows = []
    cache = {}
    for name, gen, (M, K), N in configs(args.configs.split(",")):
        t0 = time.time()
        key = name.split(" x")[0]
        if key not in cache:
            cache.clear()
            cache[key] = gen()
        rp, ci, va = cache[key]
        nnz = len(va)
        crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, ci, va))
        g = torch.Generator(device="cpu").manual_seed(1)
        B = torch.rand(K, N, generator=g).to(dev)
        G = torch.rand(M, N, generator=g).to(dev)
        C = torch.empty(M, N, device=dev)
        dB = torch.empty(K, N, device=dev)
        dA = torch.empty(nnz, device=dev)
        plan = capi.Plan(crow, col, val, (M, K))
        stream = torch.cuda.current_stream().cuda_stream
        r = {"config": name, "M": M, "K": K, "N": N, "nnz": nnz}
        fl = 2.0 * nnz * N
        by_f = synthetic.spmm_algorithmic_bytes(M, K, nnz, N)
        by_b = synthetic.spmm_algorithmic_bytes(K, M, nnz, N)
        by_a = synthetic.sddmm_algorithmic_bytes(M, K, nnz, N)
        t = timeit(lambda: plan.forward(B, C, stream=stream), args.iters, flush)
        r["fwd_ms"], r["fwd_gflops"], r["fwd_gbs"], r["fwd_frac"] = t, fl / t / 1e6, by_f / t / 1e6, by_f / t / 1e6 / PEAK
        r["fwd_path"] = ["gather", "tensor-core tiles", "strided", "smem slabs"][plan.tc_info()["last_path"]]
        r["fwd_warm_ms"] = timeit_warm(lambda: plan.forward(B, C, stream=stream), args.iters * 2)
        if r["fwd_path"] != "gather":  # the gather kernels on the same operand, for the crossover table
            from matrix_multiplication_b200._capi import OP_NO_TC
            r["fwd_gather_ms"] = timeit(lambda: plan.forward_ex(B, C, flags=OP_NO_TC, stream=stream), args.iters, flush)
            plan.forward(B, C, stream=stream)
        plan.backward_B(G, dB, stream=stream)  # builds + caches the transposed CSR (inspector time, not timed)
        t = timeit(lambda: plan.backward_B(G, dB, stream=stream), args.iters, flush)
        r["gradB_ms"], r["gradB_gbs"], r["gradB_frac"] = t, by_b / t / 1e6, by_b / t / 1e6 / PEAK
        t = timeit(lambda: plan.backward_A(G, B, dA, stream=stream), args.iters, flush)
        r["gradA_ms"], r["gradA_gbs"], r["gradA_frac"] = t, by_a / t / 1e6, by_a / t / 1e6 / PEAK
        # comparators
        try:
            Cf = torch.empty(M, N, device=dev)
            fair = cusparse_fair.FairSpMM(crow, col, val, (M, K), B, Cf, stream=stream)
            fair.run()
            r["cusparse_fair_ms"] = timeit(fair.run, args.iters, flush)
            r["cusparse_fair_warm_ms"] = timeit_warm(fair.run, args.iters * 2)
            r["rel_diff_vs_cusparse_fair"] = float((Cf - C).abs().max() / Cf.abs().max().clamp_min(1e-30))
            fair.close()
            del Cf
        except Exception as e:  # comparator only
            r["cusparse_fair_ms"] = None
            r["cusparse_fair_error"] = str(e)[:80]
        a_t = torch.sparse_csr_tensor(crow.long(), col.long(), val, size=(M, K))
        try:
            Cs = torch.sparse.mm(a_t, B)
            r["cusparse_ms"] = timeit(lambda: torch.sparse.mm(a_t, B), args.iters, flush)
            r["rel_diff_vs_cusparse"] = float((Cs - C).abs().max() / Cs.abs().max().clamp_min(1e-30))
            del Cs
        except Exception as e:
            r["cusparse_ms"] = None
            r["cusparse_error"] = str(e)[:80]
        if ref_cm is not None:
            try:
                Cr = torch.zeros(M, N, device=dev)
                r["ref_cusparse_mmul_ms"] = timeit(lambda: ref_cm.cusparse_mmul(val, col, crow, nnz, M, K, B, Cr),
                                                   max(3, args.iters // 2), flush)
            except Exception as e:
                r["ref_cusparse_mmul_ms"] = None
        if not args.no_dense and M * K <= (1 << 28):
            Ad = a_t.to_dense()
            torch.backends.cuda.matmul.allow_tf32 = False
            r["cublas_sgemm_ms"] = timeit(lambda: torch.mm(Ad, B), max(3, args.iters // 2), flush)
            torch.backends.cuda.matmul.allow_tf32 = True
            r["cublas_tf32_ms"] = timeit(lambda: torch.mm(Ad, B), max(3, args.iters // 2), flush)
            torch.backends.cuda.matmul.allow_tf32 = False
            Ah, Bh = Ad.bfloat16(), B.bfloat16()
            r["cublas_bf16_ms"] = timeit(lambda: torch.mm(Ah, Bh), max(3, args.iters // 2), flush)
            del Ah, Bh
            # dense -> CSR of the engine on the densified A (bit-exact check against the generator's CSR)
            import custom_mm
            crow2, col2, val2 = custom_mm.dense_to_csr(Ad)
            assert torch.equal(crow2, crow) and torch.equal(col2, col) and torch.equal(val2.view(torch.int32), val.view(torch.int32))
            t = timeit(lambda: custom_mm.dense_to_csr(Ad), max(3, args.iters // 2), flush)
            by_d = 4.0 * M * K + 4.0 * (M + 1) + 8.0 * nnz  # the dense matrix counted ONCE (algorithmic bytes)
            r["dense2csr_ms"], r["dense2csr_gbs"], r["dense2csr_frac"] = t, by_d / t / 1e6, by_d / t / 1e6 / PEAK
            t = timeit(lambda: Ad.to_sparse_csr(), max(3, args.iters // 2), flush)
            r["torch_to_sparse_csr_ms"] = t
            del Ad
        plan.close()
        rows.append(r)
        print(json.dumps(r), flush=True)
        print(f"[{name}] {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
    Path(args.out).parent.mkdir(exist_ok=True)
    Path(args.out).write_text(json.dumps(rows, indent=1))

    def f(x, d=3):
        return "—" if x is None else f"{x:.{d}f}"
    print("\n| config | nnz | path | fwd ms (warm) | fwd GFLOP/s | fwd GB/s (of roof) | cuSPARSE kernel-fair ms (warm) | speed-up | torch.sparse.mm ms | gather ms | SGEMM / TF32 / BF16 GEMM ms | grad_B ms (of roof) | grad_A ms (of roof) | dense→CSR ms (of roof) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        su = (r["cusparse_fair_ms"] / r["fwd_ms"]) if r.get("cusparse_fair_ms") else None
        print(f"| {r['config']} | {r['nnz']} | {r.get('fwd_path', '')} | {f(r['fwd_ms'], 4)} ({f(r.get('fwd_warm_ms'), 4)}) | {r['fwd_gflops']:.0f} | {r['fwd_gbs']:.0f} ({100 * r['fwd_frac']:.1f} %) | "
              f"{f(r.get('cusparse_fair_ms'), 4)} ({f(r.get('cusparse_fair_warm_ms'), 4)}) | {f(su, 2)}× | {f(r.get('cusparse_ms'), 4)} | {f(r.get('fwd_gather_ms'), 4)} | "
              f"{f(r.get('cublas_sgemm_ms'), 4)} / {f(r.get('cublas_tf32_ms'), 4)} / {f(r.get('cublas_bf16_ms'), 4)} | "
              f"{f(r['gradB_ms'], 4)} ({100 * r['gradB_frac']:.1f} %) | {f(r['gradA_ms'], 4)} ({100 * r['gradA_frac']:.1f} %) | "
              + (f"{f(r['dense2csr_ms'], 4)} ({100 * r['dense2csr_frac']:.1f} %) |" if "dense2csr_ms" in r else "— |"))


if __name__ == "__main__":
    main()
