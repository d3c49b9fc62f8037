#!/usr/bin/env python
"""bench.py — headline benchmark of the SpMM hot path (contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c1|c3|c5]

Metric (BASELINE.json): CSR SpMM forward throughput in GFLOP/s (2 * nnz * N per product), with the
effective GB/s on the algorithmic bytes and its fraction of the measured HBM roof in `roofline`.
Default workload (`c4`): R-MAT 2^20 x 2^20, ~16.1 M non-zeros, times a dense 2^20 x 128 operand.

A step = one C = A_csr * B over the whole matrix (all ranks together when N > 1: A is row-sharded by an
nnz-balanced cost, B is resident on every rank, every rank ends the step holding all of C).
  value : CUDA-event time, operands resident in HBM (B and C are 512 MB each, > the 126 MB L2).
  e2e   : same product through the public API (`matmuls.cusparseMM.apply`) with B coming from pinned host
          memory and C copied back to pinned host memory inside the timed region.
`--impl reference` runs the unmodified reference (oracle/_ref: its own custom_mm.so + matmuls) on the
same workload: cusparseMM.apply -> cusparse_mmul -> cusparseSpMM, with all of its per-call overheads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import warnings  # noqa: E402

import numpy as np  # noqa: E402

warnings.filterwarnings("ignore", message="Sparse CSR tensor support is in beta state")
warnings.filterwarnings("ignore", message="Sparse invariant checks are implicitly disabled")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
def make_workload(name: str, scale: int | None):
    from matrix_multiplication_b200 import synthetic
    t0 = time.time()
    cache = Path("/tmp/b200_bench_cache") / (f"{name}_{scale}".replace(":", "_").replace(".", "p") + ".npz")
    if cache.exists():  # same seeded generator, same arrays: only saves the generation time
        z = np.load(cache, allow_pickle=True)
        log(f"[bench] workload {z['desc']} loaded from {cache} in {time.time() - t0:.1f}s")
        return dict(name=name.split(":")[0], desc=str(z["desc"]), row_ptr=z["row_ptr"], col=z["col"], val=z["val"],
                    shape=(int(z["shape"][0]), int(z["shape"][1])), N=int(z["N"]), B=z["B"])
    if name == "c4":
        s = 20 if scale is None else scale
        rp, ci, va, shape = synthetic.rmat_csr(scale=s, edge_factor=16, seed=0)
        N = 128
        desc = f"c4: R-MAT 2^{s} x 2^{s} (a,b,c,d)=(.57,.19,.19,.05) seed 0, nnz {len(va)}, x dense 2^{s} x 128"
    elif name == "c1":
        rp, ci, va = synthetic.bernoulli_csr(1024, 1024, 0.9, seed=0)
        shape, N = (1024, 1024), 64
        desc = f"c1: 1024 x 1024 at 90% sparsity (nnz {len(va)}) x dense 1024 x 64"
    elif name == "c3":
        rp, ci, va = synthetic.bernoulli_csr(3072, 768, 0.9, seed=0)
        shape, N = (3072, 768), 4096
        desc = f"c3: BERT fc1 weight 3072 x 768 at 90% sparsity (nnz {len(va)}) x 768 x 4096 tokens"
    elif name.startswith("c5"):
        # c5[:sparsity[:N]]  e.g. c5:0.9:64
        parts = name.split(":")
        sp = float(parts[1]) if len(parts) > 1 else 0.9
        N = int(parts[2]) if len(parts) > 2 else 64
        dim = 16384 if scale is None else scale
        rp, ci, va = synthetic.bernoulli_csr(dim, dim, sp, seed=0)
        shape = (dim, dim)
        desc = f"c5: {dim} x {dim} at {sp * 100:g}% sparsity (nnz {len(va)}) x dense {dim} x {N}"
    else:
        raise SystemExit(f"unknown workload {name}")
    B = synthetic.dense_operand(shape[1], N, seed=1)
    log(f"[bench] workload {desc} generated in {time.time() - t0:.1f}s")
    try:
        cache.parent.mkdir(exist_ok=True)
        tmp = cache.with_suffix(f".{os.getpid()}.tmp.npz")
        np.savez(tmp, desc=desc, row_ptr=rp, col=ci, val=va, shape=np.array(shape), N=N, B=B)
        os.replace(tmp, cache)
    except OSError:
        pass
    return dict(name=name.split(":")[0], desc=desc, row_ptr=rp, col=ci, val=va, shape=shape, N=N, B=B)


def algorithmic_bytes(w) -> int:
    from matrix_multiplication_b200 import synthetic
    M, K = w["shape"]
    return synthetic.spmm_algorithmic_bytes(M, K, len(w["val"]), w["N"])


def flops(w) -> int:
    return 2 * len(w["val"]) * w["N"]


# ---------------------------------------------------------------------------------------------------
# clocks (nvidia-smi sampled during the timed region)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled every few ms during the timed region (NVML; nvidia-smi fallback)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.sm, self.mx, self.reasons = index, [], None, set()
        self._stop, self._t, self.power = threading.Event(), None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                        self.h = h
                        break
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _sample(self):
        if self.nv is not None:
            nv = self.nv
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            for n, bit in self.BAD.items():
                if r & bit:
                    self.reasons.add(n)
            try:
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
        else:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                  str(self.index)], capture_output=True, text=True, timeout=5).stdout
            r = [x.strip() for x in out.strip().split(",")]
            self.sm.append(float(r[0])); self.mx = float(r[1])
            for n, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        self._stop.clear()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": 0}
        out = {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
               "samples": len(self.sm)}
        if self.power:
            out["power_w_max"] = round(max(self.power), 1)
        return out


# ---------------------------------------------------------------------------------------------------
# CPU baseline: torch.sparse.mm on the host cores (north_star) + the oracle's OpenMP restatement
# ---------------------------------------------------------------------------------------------------
def cpu_baseline(w, budget_s: float = 20.0):
    import torch
    from oracle import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    M, K = w["shape"]
    nnz = len(w["val"])
    # bounded sample: every stride-th row of the workload (uniform over the rows, so light and heavy rows appear in
    # the matrix's own proportions), ~4M non-zeros; the whole matrix if it is smaller
    cap = 4_000_000
    stride = max(1, -(-nnz // cap))
    rp_full = w["row_ptr"].astype(np.int64)
    pick = np.arange(0, M, stride, dtype=np.int64)
    lens = rp_full[pick + 1] - rp_full[pick]
    s_rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rows, n = len(pick), int(s_rp[-1])
    src = np.repeat(rp_full[pick] - s_rp[:-1], lens) + np.arange(n, dtype=np.int64)
    s_col, s_val = np.ascontiguousarray(w["col"][src]), np.ascontiguousarray(w["val"][src])
    a = torch.sparse_csr_tensor(torch.from_numpy(s_rp), torch.from_numpy(s_col.astype(np.int64)),
                                torch.from_numpy(s_val), size=(rows, K))
    b = torch.from_numpy(w["B"])
    torch.sparse.mm(a, b)
    best, t_end, it = float("inf"), time.time() + budget_s / 2, 0
    while it < 3 or (time.time() < t_end and it < 20):
        t0 = time.perf_counter(); torch.sparse.mm(a, b); best = min(best, time.perf_counter() - t0); it += 1
    gf_torch = 2.0 * n * w["N"] / best / 1e9
    # the oracle port (C, OpenMP, fp32 CSR-order) on the same sample
    rp = s_rp.astype(np.int32)
    oracle.spmm_f32(rp, s_col, s_val, (rows, K), w["B"])
    best_o, t_end, it = float("inf"), time.time() + budget_s / 2, 0
    while it < 2 or (time.time() < t_end and it < 10):
        t0 = time.perf_counter(); oracle.spmm_f32(rp, s_col, s_val, (rows, K), w["B"])
        best_o = min(best_o, time.perf_counter() - t0); it += 1
    gf_oracle = 2.0 * n * w["N"] / best_o / 1e9
    cpu_model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                cpu_model = line.split(":", 1)[1].strip(); break
    except OSError:
        pass
    return {"value": round(gf_torch, 3), "unit": "GFLOP/s", "cores": cores, "kind": "port",
            "impl": "torch.sparse.mm (fp32 CSR, all host threads)",
            "sample": f"every {stride}th row of the workload ({rows} rows / {n} nnz, uniform over the rows), N={w['N']}, "
                      "best of >=3",
            "oracle_c_openmp_gflops": round(gf_oracle, 3), "cpu": cpu_model}


# ---------------------------------------------------------------------------------------------------
def dist_setup(n_gpus: int):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return rank, world, local, dist
    torch.cuda.set_device(0)
    return 0, 1, 0, None


def timed_region(fn_step, steps, warmup, dist, flush=None):
    """warm-up, barrier + sync, EXACTLY `steps` steps bracketed by CUDA events on the current stream,
    sync + barrier; returns (ms total by device events, ms total by wall clock)."""
    import torch
    for _ in range(warmup):
        if flush is not None:
            flush()
        fn_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    if flush is None:
        # operands exceed L2: the K steps run back to back inside ONE event bracket (a bracket per step would expose
        # ~3 us of launch latency per step, 3-4 % of a 0.09 ms multi-GPU step)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn_step()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            dist.barrier()
        return e0.elapsed_time(e1), wall
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.perf_counter()
    for i in range(steps):
        flush()  # small operands: L2 flushed before every step, each step in its own event bracket
        ev[i][0].record()
        fn_step()
        ev[i][1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    return dev_ms, wall


def max_over_ranks(x: float, dist) -> float:
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_ranks(x: float, dist):
    """x of every rank, as a list (rank order)."""
    if dist is None:
        return [x]
    import torch
    t = torch.zeros(dist.get_world_size(), dtype=torch.float64, device="cuda")
    t[dist.get_rank()] = x
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def committed_traffic():
    """dram bytes per launch of the dominant kernel, parsed from the committed ncu --set full capture summary
    (profiles/*_ncu_fwd_c4.json, written by tools/ncu_traffic.py from the .ncu-rep)."""
    best = None
    for f in sorted((ROOT / "profiles").glob("r*_ncu_fwd_c4.json")):
        try:
            d = json.loads(f.read_text())
            best = (float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), f.name)
        except Exception:
            pass
    return best


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
# the other BASELINE configs, compact (parity cases, not the headline): forward on the path the plan picks, next to the
# kernel-fair cuSPARSE call on the same operands.  Cold = 256 MB L2 flush before every timed launch.
# ---------------------------------------------------------------------------------------------------
def run_config_block(dev, peak):
    import torch
    from matrix_multiplication_b200 import capi, synthetic
    from tools import cusparse_fair
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush_buf.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    cases = [("c1 1024^2 @90% x64", "c1", None), ("c3 fc1 3072x768 @90% x4096", "c3", None),
             ("c5 16384^2 @90% x16", "c5:0.9:16", None), ("c5 16384^2 @90% x64", "c5:0.9:64", None),
             ("c5 16384^2 @90% x256", "c5:0.9:256", None), ("c5 16384^2 @99% x64", "c5:0.99:64", None)]
    rows = []
    t_start = time.time()
    for label, name, scale in cases:
        if time.time() - t_start > 75:  # keep the default run short
            rows.append({"config": label, "skipped": "time budget"})
            continue
        w = make_workload(name, scale)
        M, K = w["shape"]
        N, nnz = w["N"], len(w["val"])
        crow, col, val = (torch.from_numpy(x).to(dev) for x in (w["row_ptr"], w["col"], w["val"]))
        B = torch.from_numpy(w["B"]).to(dev)
        C = torch.empty(M, N, device=dev)
        plan = capi.Plan(crow, col, val, (M, K))
        plan.forward(B, C)
        t = timeit(lambda: plan.forward(B, C))
        by = synthetic.spmm_algorithmic_bytes(M, K, nnz, N)
        r = {"config": label, "nnz": nnz, "path": ["gather kernels", "tensor-core tiles", "strided", "smem slabs"][plan.tc_info()["last_path"]],
             "fwd_ms": round(t, 5), "gflops": round(2.0 * nnz * N / t / 1e6, 1),
             "frac_of_hbm_roof": round(by / t / 1e6 / peak, 4)}
        try:
            Cf = torch.empty(M, N, device=dev)
            fair = cusparse_fair.FairSpMM(crow, col, val, (M, K), B, Cf)
            fair.run()
            r["cusparse_kernel_fair_ms"] = round(timeit(fair.run), 5)
            r["speedup"] = round(r["cusparse_kernel_fair_ms"] / t, 2)
            r["rel_diff_vs_cusparse"] = float((Cf - C).abs().max() / Cf.abs().max().clamp_min(1e-30))
            fair.close()
        except Exception as e:
            r["cusparse_kernel_fair_ms"] = f"failed: {e}"
        plan.close()
        rows.append(r)
        del crow, col, val, B, C
        torch.cuda.empty_cache()
    return rows


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    rank, world, local, dist = dist_setup(args.gpus)
    import custom_mm
    import matmuls
    from matrix_multiplication_b200 import capi
    from matrix_multiplication_b200 import sharding

    w = make_workload(args.workload, args.scale)
    M, K = w["shape"]
    N, nnz = w["N"], len(w["val"])
    dev = torch.device("cuda", local)
    custom_mm.init_cusparse()
    for o in args.opt:
        name, value = o.split("=")
        capi.set_option(name, int(value))

    # ---- shard A's rows (nnz-balanced cost; one shard when world == 1) -----------------------------
    B_host = torch.from_numpy(w["B"]).pin_memory()
    B = B_host.to(dev)
    bounds = sharding.shard_bounds(w["row_ptr"], world, N)
    rebalance_log = []

    def upload_shard(bounds):
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        p0, p1 = int(w["row_ptr"][r0]), int(w["row_ptr"][r1])
        return (r0, r1, p0, p1, torch.from_numpy((w["row_ptr"][r0:r1 + 1] - p0).astype(np.int32)).to(dev),
                torch.from_numpy(w["col"][p0:p1]).to(dev), torch.from_numpy(w["val"][p0:p1]).to(dev))

    r0, r1, p0, p1, crow, col, val = upload_shard(bounds)
    if world > 1 and not args.no_rebalance:
        # inspector step of the multi-GPU plan: every rank times its own shard, the rows are re-cut from the measured
        # times (sharding.rebalance_bounds); six probes, every re-cut starts from the best boundaries seen so far and
        # moves a fraction of the way (halved after a move that did not help); the best set is kept.
        # Outside the timed region, like plan creation.
        best = None
        n_probes = 6
        step = 0.5
        for it in range(n_probes):
            probe = capi.Plan(crow, col, val, (r1 - r0, K))
            out_probe = torch.empty(r1 - r0, N, device=dev)
            for _ in range(3):
                probe.forward(B, out_probe)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                probe.forward(B, out_probe)
            e1.record()
            torch.cuda.synchronize()
            ts = all_ranks(e0.elapsed_time(e1) / 20, dist)
            probe.close()
            del out_probe
            rebalance_log.append({"shard_rows": [int(b) for b in bounds], "ms": [round(t, 5) for t in ts]})
            if best is None or max(ts) < best[0]:
                best = (max(ts), np.asarray(bounds).copy(), list(ts))
            else:
                step *= 0.5  # the last move made the slowest shard slower: go back to the best set, move less
            if it + 1 < n_probes:
                # a line search: the measured response to moving rows is 2-3x what the row model predicts near the
                # shard edges (full re-cuts oscillate between 1.03 and 1.05 at 8 GPUs, profiles/r2_scaling.md)
                prop = sharding.rebalance_bounds(w["row_ptr"], best[1], best[2], N)
                bounds = np.rint(best[1] + step * (prop - best[1])).astype(np.int64)
            else:
                bounds = best[1]  # (every rank holds the same all-gathered times: the same choice everywhere)
            r0, r1, p0, p1, crow, col, val = upload_shard(bounds)
    gather = sharding.RowGather(bounds, N, dev, dist) if world > 1 else None
    C_full = gather.buffer if gather is not None else torch.empty(M, N, device=dev)
    C_local = C_full[r0:r1]
    plan = capi.Plan(crow, col, val, (r1 - r0, K))
    stream = torch.cuda.current_stream().cuda_stream

    small = (algorithmic_bytes(w) < 400e6)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None
    flush = (lambda: flush_buf.fill_(1)) if small else None

    def step_resident():  # every rank computes its row block of C (C stays row-sharded: no data-path collective)
        plan.forward(B, C_local, stream=stream)

    def step_gathered():  # ... and then every rank collects all of C over NVLink (reported separately)
        plan.forward(B, C_local, stream=stream)
        gather.all_gather()

    clk = ClockSampler(local)
    n0 = capi.kernel_launch_count()
    with clk:  # clocks are sampled across all timed regions
        dev_ms, wall_ms = timed_region(step_resident, args.steps, args.warmup, dist, flush)
    launches = capi.kernel_launch_count() - n0
    kern_ms = dev_ms  # this rank's own device time = its spmm_flat_kernel launches (one per step)
    per_rank_ms = all_ranks(dev_ms / args.steps, dist)
    dev_ms = max_over_ranks(dev_ms, dist)
    ms_per_step = dev_ms / args.steps
    gathered_ms = None
    fused = {}
    if gather is not None:
        with clk:
            g_ms, _ = timed_region(step_gathered, args.steps, args.warmup, dist, flush)
        gathered_ms = max_over_ranks(g_ms, dist) / args.steps
        # the same gather fused into the SpMM epilogue (peer stores / NVLS multicast stores), no NCCL call
        for mode in ("p2p", "multicast"):
            try:
                sym = sharding.SymmetricRowGather(bounds, N, dev, dist, use_multicast=(mode == "multicast"))
                if mode == "multicast" and not sym.mc_ptr:
                    fused[mode] = "no multicast support"
                    continue

                def step_fused(sym=sym):
                    sym.forward(plan, B, stream=stream)
                    sym.barrier()

                with clk:
                    f_ms, _ = timed_region(step_fused, args.steps, args.warmup, dist, flush)
                f_ms = max_over_ranks(f_ms, dist) / args.steps
                torch.cuda.synchronize()
                dist.barrier()
                same = bool(torch.equal(sym.buffer, C_full)) if mode else None
                ok = torch.tensor([1.0 if same else 0.0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                fused[mode] = {"ms_per_step": round(f_ms, 5), "value": round(flops(w) / (f_ms * 1e-3) / 1e9, 2),
                               "identical_to_nccl_gather_on_all_ranks": bool(ok.item() == 1.0)}
                del sym
            except Exception as e:  # symmetric memory is plumbing that may be unavailable on a box
                fused[mode] = f"unavailable: {type(e).__name__}: {str(e)[:120]}"

    # ---- multi-GPU phases the north_star names (all ranks take part): B broadcast, C gathered to one rank, and the
    # row-sharded backward (grad_A local, grad_B = reduce-scatter of the ranks' partial A_g^T G_g) ----------------
    mg_phases = None
    if world > 1:
        mg_phases = {}
        Bb = torch.empty_like(B)
        if rank == 0:
            Bb.copy_(B)
        ms, _ = timed_region(lambda: dist.broadcast(Bb, src=0), args.steps, args.warmup, dist, None)
        ms = max_over_ranks(ms, dist) / args.steps
        mg_phases["B_broadcast_from_rank0_ms"] = round(ms, 5)
        mg_phases["B_broadcast_GBs_per_rank_ingress"] = round(K * N * 4 / (ms * 1e-3) / 1e9, 1)
        del Bb
        ms, _ = timed_region(lambda: gather.gather_to(0), args.steps, args.warmup, dist, None)
        ms = max_over_ranks(ms, dist) / args.steps
        mg_phases["C_gather_to_rank0_ms"] = round(ms, 5)
        mg_phases["C_gather_rank0_ingress_GBs"] = round((M - (r1 - r0 if rank == 0 else int(bounds[1] - bounds[0]))) * N * 4
                                                        / (ms * 1e-3) / 1e9, 1)
        ms, _ = timed_region(lambda: (plan.forward(B, C_local, stream=stream), gather.gather_to(0)), args.steps,
                             args.warmup, dist, None)
        mg_phases["forward_plus_gather_to_rank0_ms"] = round(max_over_ranks(ms, dist) / args.steps, 5)
        if K % world == 0:
            G_local = torch.rand(r1 - r0, N, device=dev)
            dA_local = torch.empty(p1 - p0, device=dev)
            dB_part = torch.empty(K, N, device=dev)
            dB_mine = torch.empty(K // world, N, device=dev)
            plan.backward_B(G_local, dB_part, stream=stream)  # builds the transposed shard (inspector time)
            ga, _ = timed_region(lambda: plan.backward_A(G_local, B, dA_local, stream=stream), args.steps, args.warmup,
                                 dist, None)
            gb, _ = timed_region(lambda: plan.backward_B(G_local, dB_part, stream=stream), args.steps, args.warmup, dist,
                                 None)
            rs, _ = timed_region(lambda: sharding.reduce_scatter_rows(dB_part, dB_mine, dist), args.steps, args.warmup,
                                 dist, None)

            def bwd_step():
                plan.backward_A(G_local, B, dA_local, stream=stream)
                plan.backward_B(G_local, dB_part, stream=stream)
                sharding.reduce_scatter_rows(dB_part, dB_mine, dist)
            bw, _ = timed_region(bwd_step, args.steps, args.warmup, dist, None)
            ga, gb, rs, bw = (max_over_ranks(x, dist) / args.steps for x in (ga, gb, rs, bw))
            mg_phases.update({
                "grad_A_local_sddmm_ms": round(ga, 5), "grad_B_partial_spmm_ms": round(gb, 5),
                "grad_B_reduce_scatter_ms": round(rs, 5), "backward_step_ms": round(bw, 5),
                "fwd_bwd_ms": round(ms_per_step + bw, 5),
                "fwd_bwd_gflops": round(3 * flops(w) / ((ms_per_step + bw) * 1e-3) / 1e9, 1),
                "what": "grad_A: SDDMM on the shard's pattern (no exchange); grad_B: every rank's [K, N] partial A_g^T G_g "
                        "then one NCCL reduce-scatter (rank g keeps K/G rows); link bound of the exchange = "
                        f"{(world - 1) / world * K * N * 4 / 770e9 * 1e3:.3f} ms at 770 GB/s per direction"})
            del G_local, dA_local, dB_part, dB_mine

    # ---- e2e through the public API with host buffers -------------------------------------------------
    a_t = torch.sparse_csr_tensor(crow.long(), col.long(), val, size=(r1 - r0, K))
    C_host = torch.empty(r1 - r0, N).pin_memory()  # each rank returns its own row block to the host
    B_in = torch.empty_like(B)

    def step_e2e():
        B_in.copy_(B_host, non_blocking=True)
        c = matmuls.cusparseMM.apply(a_t, B_in)
        C_host.copy_(c, non_blocking=True)

    e2e_steps = max(3, min(args.steps, 10))
    with clk:
        e2e_ms, _ = timed_region(step_e2e, e2e_steps, 3, dist, None)
    e2e_torch_ms = max_over_ranks(e2e_ms, dist) / e2e_steps

    # ... and through the C ABI's host-buffer entry point (b200_plan_forward_host): the library pipelines the upload
    # of B, the product and the download of C over column panels on its own streams; the call returns when C is in
    # host memory, so it is timed as a blocking call (barrier + synchronize on both sides)
    if world > 1 and K % world == 0:
        # N > 1: the dense operand exists once in host memory; every rank uploads 1/N of its rows over its own
        # PCIe link and the slices are all-gathered over NVLink (NCCL), then each rank multiplies its row shard of
        # A and returns its row block of C to the host.  Total PCIe traffic per step = B once + C once.
        ks = K // world
        B_slice_host = B_host[rank * ks:(rank + 1) * ks]

        def step_host():
            B_in[rank * ks:(rank + 1) * ks].copy_(B_slice_host, non_blocking=True)
            dist.all_gather_into_tensor(B_in, B_in[rank * ks:(rank + 1) * ks])
            plan.forward(B_in, C_local, stream=stream)
            C_host.copy_(C_local, non_blocking=True)
        e2e_api = ("per-rank upload of K/N rows of B (pinned) + NCCL all-gather over NVLink + b200_plan_forward on the "
                   "row shard + download of the rank's rows of C")
        e2e_h2d = int(B_host.numel() * 4)
    else:
        def step_host():
            plan.forward_host(B_host, C_host)
        e2e_api = "b200_plan_forward_host (C ABI, pinned host buffers)"
        e2e_h2d = int(B_host.numel() * 4) * world

    for _ in range(3):
        step_host()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    with clk:
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        torch.cuda.synchronize()
        host_ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        dist.barrier()
    e2e_ms = max_over_ranks(host_ms, dist) / e2e_steps
    e2e_host_check = float((C_host.to(dev) - C_local).abs().max())  # same bits as the resident product

    if gather is not None:  # the parity check below reads all of C on rank 0
        step_gathered()
        torch.cuda.synchronize()

    # ---- parity spot check of what was just timed (oracle as checker, not measured) ---------------------
    parity = None
    if rank == 0:
        from oracle import oracle
        rows = np.unique(np.concatenate([np.argsort(np.diff(w["row_ptr"]))[-4:], np.arange(0, M, max(1, M // 32))]))
        lens = np.diff(w["row_ptr"])[rows]
        sub_rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        idx = np.concatenate([np.arange(w["row_ptr"][r], w["row_ptr"][r + 1]) for r in rows]) if lens.sum() else np.zeros(0, np.int64)
        ref = oracle.spmm_f64(sub_rp, w["col"][idx], w["val"][idx], (len(rows), K), w["B"])
        got = C_full[torch.from_numpy(rows).to(dev)].cpu().numpy()
        parity = oracle.rel_err(got, ref)

    if rank != 0:
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- backward of the same product (north_star: forward + backward), single GPU only ---------------------
    phases = None
    if world == 1:
        from matrix_multiplication_b200 import synthetic
        Gd = torch.rand(M, N, device=dev)
        dB = torch.empty(K, N, device=dev)
        dA = torch.empty(nnz, device=dev)
        plan.backward_B(Gd, dB, stream=stream)  # builds + caches the transposed CSR (inspector time, not timed)
        gb_ms, _ = timed_region(lambda: plan.backward_B(Gd, dB, stream=stream), args.steps, args.warmup, None, flush)
        ga_ms, _ = timed_region(lambda: plan.backward_A(Gd, B, dA, stream=stream), args.steps, args.warmup, None, flush)
        gb_ms, ga_ms = gb_ms / args.steps, ga_ms / args.steps
        by = (synthetic.spmm_algorithmic_bytes(M, K, nnz, N) + synthetic.spmm_algorithmic_bytes(K, M, nnz, N)
              + synthetic.sddmm_algorithmic_bytes(M, K, nnz, N))
        tot = ms_per_step + gb_ms + ga_ms
        phases = {"forward_ms": round(ms_per_step, 5), "grad_B_spmm_over_cached_transpose_ms": round(gb_ms, 5),
                  "grad_A_sddmm_ms": round(ga_ms, 5), "fwd_bwd_ms": round(tot, 5),
                  "fwd_bwd_gflops": round(3 * flops(w) / (tot * 1e-3) / 1e9, 1),
                  "fwd_bwd_effective_gbs": round(by / (tot * 1e-3) / 1e9, 1),
                  "fwd_bwd_frac_of_hbm_roof": round(by / (tot * 1e-3) / 1e9 / peaks()[0], 4)}
        del Gd, dB, dA

    # ---- comparators on the same GPU (single-GPU only; informational) ------------------------------------
    comp = {}
    if world == 1 and not args.no_comparators:
        try:
            a_full = torch.sparse_csr_tensor(crow.long(), col.long(), val, size=(M, K))
            torch.sparse.mm(a_full, B)
            ms, _ = timed_region(lambda: torch.sparse.mm(a_full, B), max(3, args.steps // 2), 2, None, flush)
            comp["torch_sparse_mm_cusparse_ms"] = round(ms / max(3, args.steps // 2), 4)
        except Exception as e:
            comp["torch_sparse_mm_cusparse_ms"] = f"failed: {e}"
        try:
            from oracle import build_ref
            if build_ref.available():
                rcm, _ = build_ref.load()
                rcm.init_cusparse()
                Cc = torch.zeros(M, N, device=dev)
                f = lambda: rcm.cusparse_mmul(val, col, crow, nnz, M, K, B, Cc)  # noqa: E731
                f()
                ms, _ = timed_region(f, max(3, args.steps // 2), 2, None, flush)
                comp["reference_cusparse_mmul_ms"] = round(ms / max(3, args.steps // 2), 4)
                comp["reference_cusparse_rel_diff"] = float((Cc - C_full).abs().max() / C_full.abs().max())
        except Exception as e:
            comp["reference_cusparse_mmul_ms"] = f"failed: {e}"

        try:
            from tools import cusparse_fair
            Cf = torch.empty(M, N, device=dev)
            fair = cusparse_fair.FairSpMM(crow, col, val, (M, K), B, Cf, stream=stream)
            fair.run()
            ms, _ = timed_region(fair.run, max(3, args.steps // 2), 2, None, flush)
            comp["cusparse_spmm_alg2_kernel_fair_ms"] = round(ms / max(3, args.steps // 2), 4)
            comp["cusparse_kernel_fair_what"] = ("cusparseSpMM CSR_ALG2, row-major B/C as the reference calls it "
                                                 "(src/baseline_mm.cu:186-213), descriptors + buffer + preprocess hoisted")
            comp["speedup_vs_cusparse_kernel_fair"] = round(comp["cusparse_spmm_alg2_kernel_fair_ms"] / ms_per_step, 3)
            fair.close()
            del Cf
        except Exception as e:
            comp["cusparse_spmm_alg2_kernel_fair_ms"] = f"failed: {e}"

    configs_block = None
    if world == 1 and not args.no_configs:
        try:
            configs_block = run_config_block(dev, peaks()[0])
        except Exception as e:  # informational block: never lose the headline line over it
            configs_block = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    cpu = cpu_baseline(w) if world == 1 and not args.no_cpu else None
    peak, peak_src = peaks()
    # the dominant kernel: spmm_fwd_kernel over this rank's shard (rank 0's shard for the roofline)
    sh_bytes = 4 * (r1 - r0 + 1) + 8 * (p1 - p0) + 4 * K * N + 4 * (r1 - r0) * N
    k_ms = kern_ms / args.steps
    achieved = sh_bytes / (k_ms * 1e-3) / 1e9
    # SURVEY 8(d)'s formula charges all K rows of B to the launch; a shard (and an R-MAT with empty columns) touches
    # fewer: the same figure with only the distinct B rows this rank's non-zeros name, reported beside it
    b_rows_touched = int(torch.unique(col).numel()) if col.numel() else 0
    sh_bytes_touched = sh_bytes - 4 * (K - b_rows_touched) * N
    out = {
        "metric": "CSR SpMM forward GFLOP/s (2*nnz*N)", "value": round(flops(w) / (ms_per_step * 1e-3) / 1e9, 2),
        "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "M": M, "K": K, "N": N, "nnz": nnz,
                   "cache": "L2 flushed between steps (256 MB fill)" if small else
                   f"operands exceed L2 (B {K * N * 4 >> 20} MB, C {M * N * 4 >> 20} MB)",
                   "parallelism": (f"row-shard x{world} (cost-balanced on nnz and rows), B resident on every rank, C left "
                                   "row-sharded; 'gathered' = the same step + NCCL all-gather of C") if world > 1 else "single GPU",
                   "shard_rows": [int(b) for b in bounds],
                   "shard_rebalance": rebalance_log or None},
        "effective_gbs": round(algorithmic_bytes(w) / (ms_per_step * 1e-3) / 1e9, 1),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": args.traffic,
                     "traffic_source": args.traffic_source,
                     "kernel": "spmm_flat_kernel (single launch: split rows are summed by the last arriving task)",
                     "kernel_ms": round(k_ms, 5), "kernel_ms_per_rank": [round(x, 5) for x in per_rank_ms],
                     "algorithmic_bytes": sh_bytes,
                     "b_rows_touched": b_rows_touched, "algorithmic_bytes_touched_rows_of_B_only": sh_bytes_touched,
                     "frac_touched_rows_of_B_only": round(sh_bytes_touched / (k_ms * 1e-3) / 1e9 / peak, 4),
                     "scope": "rank 0's launch over its own shard against ONE GPU's peak",
                     "peak_source": peak_src, "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)},
        "e2e": {"value": round(flops(w) / (e2e_ms * 1e-3) / 1e9, 2), "unit": "GFLOP/s",
                "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": int(M * N * 4),
                "ms_per_step": round(e2e_ms, 4), "api": e2e_api,
                "max_abs_diff_vs_resident": e2e_host_check,
                "torch_api": {"api": "matmuls.cusparseMM.apply(a_csr, b) + pinned copies on one stream",
                              "ms_per_step": round(e2e_torch_ms, 4),
                              "value": round(flops(w) / (e2e_torch_ms * 1e-3) / 1e9, 2)}},
        # what actually limits this workload: every non-zero pulls an N*4-byte B row L2 -> SM;
        # ceiling = best rate of tools/l2bw.cu for 512-byte rows with the same column skew (profiles/r1_variant_sweep.md)
        "binding_roof": {"name": "L2->SM gather of B rows", "bytes": int((p1 - p0)) * N * 4,
                         "achieved_TBs": round((p1 - p0) * N * 4 / (k_ms * 1e-3) / 1e12, 2), "measured_ceiling_TBs": 20.25},
        "gpu_launches": int(launches), "clocks": clk.summary(), "parity_rel_err_vs_fp64_oracle": parity,
        "wall_ms_per_step": round(wall_ms / args.steps, 5),
        "comparators": comp,
    }
    if phases is not None:
        out["phases"] = phases
    if mg_phases is not None:
        out["phases"] = mg_phases
    if configs_block is not None:
        out["configs"] = configs_block
    if gathered_ms is not None:
        out["gathered"] = {"value": round(flops(w) / (gathered_ms * 1e-3) / 1e9, 2), "unit": "GFLOP/s",
                           "ms_per_step": round(gathered_ms, 5),
                           "what": "same step + every rank receives all of C (one grouped NCCL all-gather of unequal row blocks)",
                           "fused_epilogue": fused}
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out), flush=True)
    plan.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference (oracle/_ref) on the same workload
# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import build_ref
    w = make_workload(args.workload, args.scale)
    M, K = w["shape"]
    N, nnz = w["N"], len(w["val"])
    base = {"impl": "reference", "metric": "CSR SpMM forward GFLOP/s (2*nnz*N)", "unit": "GFLOP/s",
            "n_gpus": args.gpus, "gpus_used": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "M": M, "K": K, "N": N, "nnz": nnz}}
    cpu = cpu_baseline(w)
    if not (torch.cuda.is_available() and build_ref.available()):
        # no GPU or no reference build: the reference's path restated on the CPU (torch.sparse.mm)
        base.update({"value": cpu["value"], "ms_per_step": None, "cpu_baseline": cpu,
                     "e2e": {"value": cpu["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0},
                     "note": "oracle/_ref not available: CPU port timed instead"})
        print(json.dumps(base), flush=True)
        return
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    rcm, rmm = build_ref.load()
    rcm.init_cusparse()
    a = torch.sparse_csr_tensor(torch.from_numpy(w["row_ptr"].astype(np.int64)), torch.from_numpy(w["col"].astype(np.int64)),
                                torch.from_numpy(w["val"]), size=(M, K)).to(dev)
    B_host = torch.from_numpy(w["B"]).pin_memory()
    B = B_host.to(dev)
    C_host = torch.empty(M, N).pin_memory()
    B_in = torch.empty_like(B)
    out_keep = {}

    def step():
        out_keep["c"] = rmm.cusparseMM.apply(a, B)

    def step_e2e():
        B_in.copy_(B_host, non_blocking=True)
        c = rmm.cusparseMM.apply(a, B_in)
        C_host.copy_(c, non_blocking=True)

    small = (algorithmic_bytes(w) < 400e6)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None
    flush = (lambda: flush_buf.fill_(1)) if small else None
    with ClockSampler(0) as clk:
        dev_ms, wall_ms = timed_region(step, args.steps, args.warmup, None, flush)
    ms = dev_ms / args.steps
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms, _ = timed_region(step_e2e, e2e_steps, 3, None, None)
    e2e_ms /= e2e_steps
    base.update({"value": round(flops(w) / (ms * 1e-3) / 1e9, 2), "ms_per_step": round(ms, 5),
                 "effective_gbs": round(algorithmic_bytes(w) / (ms * 1e-3) / 1e9, 1),
                 "e2e": {"value": round(flops(w) / (e2e_ms * 1e-3) / 1e9, 2), "unit": "GFLOP/s",
                         "h2d_bytes_per_step": int(B_host.numel() * 4), "d2h_bytes_per_step": int(C_host.numel() * 4),
                         "ms_per_step": round(e2e_ms, 4), "api": "reference matmuls.cusparseMM.apply(a_csr, b)"},
                 "cpu_baseline": cpu, "clocks": clk.summary(), "wall_ms_per_step": round(wall_ms / args.steps, 5),
                 "reference_build": "oracle/_ref (reference src/ compiled unmodified for sm_100a; "
                                    "-DCUSPARSE_MM_ALG_DEFAULT=CUSPARSE_SPMM_ALG_DEFAULT)"})
    print(json.dumps(base), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--scale", type=int, default=None, help="R-MAT scale (c4) or matrix dimension (c5) override")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-comparators", action="store_true")
    ap.add_argument("--no-rebalance", action="store_true", help="keep the static cost-model shards at N > 1")
    ap.add_argument("--no-configs", action="store_true", help="skip the compact block of the other BASELINE configs")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (b200_set_option)")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu --set full capture "
                         "(default: the committed figure for the default workload on 1 GPU)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.traffic_source = "--traffic" if args.traffic is not None else None
    if args.traffic is None and args.workload == "c4" and args.scale is None and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        # dram__bytes_read.sum + dram__bytes_write.sum of spmm_flat_kernel from the committed capture of this workload
        t = committed_traffic()
        if t is not None:
            args.traffic, args.traffic_source = t
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
