#!/usr/bin/env python
"""BASELINE configs 2 and 3: sparse-weight FC layers, forward + backward, synthetic batch, one B200.

  config 2  MNIST-shaped MLP 784-512-10, 90 %-sparse weights, batch 256       (benchmarks/cusparse_fc_layer.py shape)
  config 3  BERT-base FC block 768 -> 3072 -> 768, 90 %-sparse weights, 8 x 512 tokens

Arms (same weights, same inputs, CUDA-event timing of forward + loss.sum().backward(), median of --iters):
  engine        matrix_multiplication_b200.layers.cusparseLinear (persistent CSR weight, cached transpose, SDDMM grad)
  engine_fused  same with the ReLU of the first layer fused into its product (cusparseLinear(activation='relu'));
                also replayed from a CUDA graph (launch-bound configs: one launch per step)
  torch_sparse  torch.sparse.mm on a CSR weight (cuSPARSE forward; torch's own sparse backward)
  dense         nn.Linear on the densified weights (cuBLAS SGEMM fp32)
  reference     the reference's cusparse FC layer call, `cusparseMM.apply(inp, weight.t())`
                (benchmarks/cusparse_fc_layer.py:41) from oracle/_ref, forward only — its backward densifies
                grad_output into a CSR and multiplies mis-strided views, so only its forward is timed
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn  # noqa: E402

from matrix_multiplication_b200.layers import cusparseLinear  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


class TorchSparseLinear(nn.Module):
    def __init__(self, lin: cusparseLinear):
        super().__init__()
        self.crow, self.col = lin.crow.long(), lin.col.long()
        self.values = nn.Parameter(lin.values.detach().clone())
        self.bias = nn.Parameter(lin.bias.detach().clone())
        self.shape = (lin.out_features, lin.in_features)

    def forward(self, x):
        w = torch.sparse_csr_tensor(self.crow, self.col, self.values, size=self.shape)
        return torch.sparse.mm(w, x.t()).t() + self.bias


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "fc_layer.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    for name, dims, tokens in (("c2 MLP 784-512-10, batch 256", (784, 512, 10), 256),
                               ("c3 BERT FC 768-3072-768, 8x512 tokens", (768, 3072, 768), 4096)):
        torch.manual_seed(0)
        l1 = cusparseLinear(dims[0], dims[1], sparsity=0.9, seed=1).to(dev)
        l2 = cusparseLinear(dims[1], dims[2], sparsity=0.9, seed=2).to(dev)
        f1 = cusparseLinear(dims[0], dims[1], sparsity=0.9, seed=1, activation="relu").to(dev)  # same weights as l1
        x = torch.rand(tokens, dims[0], device=dev)
        xT = x.t().contiguous()
        nnz = l1.values.numel() + l2.values.numel()
        flops = 3 * 2.0 * nnz * tokens  # forward + grad_x + grad_w per layer

        def step_engine():
            for p in (l1.values, l2.values, l1.bias, l2.bias):
                p.grad = None
            l2(torch.relu(l1(x))).sum().backward()

        def step_engine_t():
            for p in (f1.values, l2.values, f1.bias, l2.bias):
                p.grad = None
            l2(f1(x)).sum().backward()

        t1, t2 = TorchSparseLinear(l1).to(dev), TorchSparseLinear(l2).to(dev)

        def step_torch_sparse():
            for p in (t1.values, t2.values, t1.bias, t2.bias):
                p.grad = None
            t2(torch.relu(t1(x))).sum().backward()

        d1, d2 = nn.Linear(dims[0], dims[1]).to(dev), nn.Linear(dims[1], dims[2]).to(dev)
        with torch.no_grad():
            d1.weight.copy_(l1.dense_weight()); d1.bias.copy_(l1.bias)
            d2.weight.copy_(l2.dense_weight()); d2.bias.copy_(l2.bias)

        def step_dense():
            d1.zero_grad(set_to_none=True); d2.zero_grad(set_to_none=True)
            d2(torch.relu(d1(x))).sum().backward()

        # parity of the arms before timing
        with torch.no_grad():  # (no autograd graph may outlive this line: CUDA-graph capture below)
            y_e = l2(torch.relu(l1(x)))
            y_d = d2(torch.relu(d1(x)))
        rel = float((y_e - y_d).abs().max() / y_d.abs().max())
        r = {"config": name, "nnz": nnz, "tokens": tokens, "rel_diff_vs_dense": rel}
        # the same engine step captured once in a CUDA graph and replayed (launch-bound configs: one launch per step)
        try:
            for p in (f1.values, l2.values, f1.bias, l2.bias):
                p.grad = None
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step_engine_t()
            torch.cuda.current_stream().wait_stream(side)
            for p in (f1.values, l2.values, f1.bias, l2.bias):
                p.grad = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                l2(f1(x)).sum().backward()
            graph.replay()
            torch.cuda.synchronize()
            g_graph = f1.values.grad.clone()
            step_engine_t()
            r["graph_grad_rel_diff"] = float((g_graph - f1.values.grad).abs().max() / f1.values.grad.abs().max())
            r["engine_fused_graph_ms"] = timeit(graph.replay, args.iters)
        except Exception as e:
            r["engine_fused_graph_error"] = f"{type(e).__name__}: {str(e)[:160]}"
        r["engine_ms"] = timeit(step_engine, args.iters)
        r["engine_fused_ms"] = timeit(step_engine_t, args.iters)
        r["paths"] = {"fc1": l1.csr().plan().last_path(), "fc2": l2.csr().plan().last_path()}
        try:
            r["torch_sparse_ms"] = timeit(step_torch_sparse, args.iters)
        except Exception as e:
            r["torch_sparse_ms"] = None
            r["torch_sparse_error"] = str(e)[:100]
        r["dense_ms"] = timeit(step_dense, args.iters)
        r["engine_gflops"] = flops / r["engine_fused_ms"] / 1e6
        try:
            from oracle import build_ref
            rcm, rmm = build_ref.load()
            rcm.init_cusparse()
            w1t, w2t = l1.dense_weight().t().contiguous(), l2.dense_weight().t().contiguous()

            def ref_forward():  # benchmarks/cusparse_fc_layer.py:41 — the INPUT is what gets CSR-ified, per call
                h = torch.relu(rmm.cusparseMM.apply(x, w1t) + l1.bias)
                return rmm.cusparseMM.apply(h, w2t) + l2.bias
            yr = ref_forward()
            r["reference_fwd_rel_diff"] = float((yr - y_d).abs().max() / y_d.abs().max())
            r["reference_fwd_only_ms"] = timeit(ref_forward, max(3, args.iters // 4))
            with torch.no_grad():
                r["engine_fwd_only_ms"] = timeit(lambda: l2(torch.relu(l1(x))), args.iters)
        except Exception as e:
            r["reference_error"] = f"{type(e).__name__}: {str(e)[:120]}"
        rows.append(r)
        print(json.dumps(r), flush=True)
    Path(args.out).parent.mkdir(exist_ok=True)
    Path(args.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
