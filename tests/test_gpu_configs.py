"""Parity on the BASELINE.json configurations themselves (the sizes the bench numbers are quoted on), through the
C ABI's plan entry points, against the fp64 oracle.  Runs on the B200 box only (`-m gpu`).

  config 3  BERT FC layers 3072x768 / 768x3072 at 90 % sparsity x 4096 tokens: forward, grad_B, grad_A — on the
            path the plan picks by itself (block-dense tensor-core tiles) and on the gather kernels
  config 5  16384 x 16384 at sparsity 0.9 / 0.99 / 0.5 with N in {16, 64, 256}: forward (full fp64 oracle while
            nnz * N <= 1e10, row-sampled beyond)
  config 4  R-MAT scale 20 (16.09 M nnz) x 128: every split row (longer than one task), its neighbours, and 4096
            uniformly sampled rows

Tolerance (north_star): max|x - ref64| / max|ref64| <= 1e-5."""
import numpy as np
import pytest

from matrix_multiplication_b200 import capi, synthetic
from matrix_multiplication_b200._capi import OP_NO_TC
from oracle import oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev(*arrays):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0") for a in arrays]


def _sub_csr(rp, ci, va, rows):
    lens = np.diff(rp)[rows]
    sub_rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([np.arange(rp[r], rp[r + 1]) for r in rows]) if len(rows) else np.zeros(0, np.int64)
    return sub_rp.astype(np.int32), ci[idx], va[idx]


@pytest.mark.parametrize("layer", ["fc1_3072x768", "fc2_768x3072"])
@pytest.mark.parametrize("path", ["auto", "gather"])
def test_config3_bert_fc_layers(layer, path):
    import torch
    M, K = (3072, 768) if layer.startswith("fc1") else (768, 3072)
    N = 4096
    rp, ci, va = synthetic.bernoulli_csr(M, K, 0.9, seed=3, signed=True)
    B = synthetic.dense_operand(K, N, seed=4, signed=True)
    G = synthetic.dense_operand(M, N, seed=5, signed=True)
    crow, col, val, Bd, Gd = _dev(rp, ci, va, B, G)
    plan = capi.Plan(crow, col, val, (M, K))
    flags = OP_NO_TC if path == "gather" else 0
    C = plan.forward_ex(Bd, torch.empty(M, N, device="cuda:0"), flags=flags)
    assert plan.tc_info()["last_path"] == (0 if path == "gather" else 1), "config 3 is dense enough for the tiles"
    dB = plan.backward_B_ex(Gd, torch.empty(K, N, device="cuda:0"), flags=flags)
    dA = plan.backward_A_ex(Gd, Bd, torch.empty(len(va), device="cuda:0"), flags=flags)
    assert plan.tc_info()["last_path"] == (0 if path == "gather" else 1)
    assert plan.would_use_tc(N, side=2) == (path != "gather") or path == "gather"
    torch.cuda.synchronize()
    assert plan.tc_timeout() == 0
    assert oracle.rel_err(C.cpu().numpy(), oracle.spmm_f64(rp, ci, va, (M, K), B)) <= TOL
    assert oracle.rel_err(dB.cpu().numpy(), oracle.spmm_t_f64(rp, ci, va, (M, K), G)) <= TOL
    assert oracle.rel_err(dA.cpu().numpy(), oracle.sddmm_f64(rp, ci, (M, K), G, B)) <= TOL
    plan.close()


@pytest.fixture(scope="module")
def c5_matrices():
    cache = {}

    def get(sparsity):
        if sparsity not in cache:
            cache.clear()  # one 16k x 16k matrix at a time in host memory
            cache[sparsity] = synthetic.bernoulli_csr(16384, 16384, sparsity, seed=0)
        return cache[sparsity]
    return get


@pytest.mark.parametrize("sparsity,N", [(0.9, 16), (0.9, 256), (0.99, 64), (0.5, 16), (0.5, 64)])
def test_config5_sweep_points(c5_matrices, sparsity, N):
    import torch
    M = K = 16384
    rp, ci, va = c5_matrices(sparsity)
    B = synthetic.dense_operand(K, N, seed=1)
    crow, col, val, Bd = _dev(rp, ci, va, B)
    plan = capi.Plan(crow, col, val, (M, K))
    C = plan.forward(Bd, torch.empty(M, N, device="cuda:0"))
    torch.cuda.synchronize()
    assert plan.tc_timeout() == 0
    if not (sparsity == 0.9 and N == 16):  # (that point is a wash between the two paths: either is fine)
        want_tc = sparsity <= 0.9  # the cost model: the tiles win from ~5 % density on an operand of this size
        assert plan.tc_info()["last_path"] == (1 if want_tc else 0)
    Ch = C.cpu().numpy()
    if len(va) * N <= 1e10:
        ref = oracle.spmm_f64(rp, ci, va, (M, K), B)
        assert oracle.rel_err(Ch, ref) <= TOL
    else:
        rows = np.arange(0, M, 8)
        srp, sci, sva = _sub_csr(rp, ci, va, rows)
        ref = oracle.spmm_f64(srp, sci, sva, (len(rows), K), B)
        assert oracle.rel_err(Ch[rows], ref) <= TOL
    plan.close()


@pytest.mark.parametrize("sparsity,N", [(0.9, 64), (0.99, 64)])
def test_config5_backward_points(c5_matrices, sparsity, N):
    """grad_B over the cached transpose and grad_A (SDDMM on the pattern) at the sweep's own size, on whatever path the
    plan's cost models pick (90 % sparse x 64: tiles for both; 99 %: gather kernels)."""
    import torch
    M = K = 16384
    rp, ci, va = c5_matrices(sparsity)
    B = synthetic.dense_operand(K, N, seed=1, signed=True)
    G = synthetic.dense_operand(M, N, seed=2, signed=True)
    crow, col, val, Bd, Gd = _dev(rp, ci, va, B, G)
    plan = capi.Plan(crow, col, val, (M, K))
    dB = plan.backward_B(Gd, torch.empty(K, N, device="cuda:0"))
    dA = plan.backward_A_ex(Gd, Bd, torch.empty(len(va), device="cuda:0"))
    torch.cuda.synchronize()
    assert plan.tc_timeout() == 0
    assert plan.tc_info()["last_path"] == (1 if plan.would_use_tc(N, side=2) else 0)
    assert oracle.rel_err(dB.cpu().numpy(), oracle.spmm_t_f64(rp, ci, va, (M, K), G)) <= TOL
    assert oracle.rel_err(dA.cpu().numpy(), oracle.sddmm_f64(rp, ci, (M, K), G, B)) <= TOL
    # the other SDDMM path on the same operands agrees too
    dA2 = plan.backward_A_ex(Gd, Bd, torch.empty(len(va), device="cuda:0"), flags=OP_NO_TC)
    torch.cuda.synchronize()
    assert plan.tc_info()["last_path"] == 0
    assert oracle.rel_err(dA2.cpu().numpy(), dA.cpu().numpy().astype(np.float64)) <= 2 * TOL
    plan.close()


def test_config4_rmat_scale20_split_rows_and_sample():
    import torch
    rp, ci, va, shape = synthetic.rmat_csr(scale=20, edge_factor=16, seed=0)
    M, K = shape
    N = 128
    B = synthetic.dense_operand(K, N, seed=1)
    crow, col, val, Bd = _dev(rp, ci, va, B)
    plan = capi.Plan(crow, col, val, shape)
    C = plan.forward(Bd, torch.empty(M, N, device="cuda:0"))
    C2 = plan.forward(Bd, torch.empty(M, N, device="cuda:0"))  # the piece counters re-arm themselves
    torch.cuda.synchronize()
    assert plan.tc_info()["last_path"] == 0
    assert torch.equal(C, C2), "repeated launches must be bit-identical"
    lens = np.diff(rp)
    split = np.where(lens > 256)[0]  # rows longer than one 256-item task are cut into pieces
    assert len(split) > 1000
    rng = np.random.default_rng(0)
    rows = np.unique(np.concatenate([split, np.clip(split - 1, 0, M - 1), np.clip(split + 1, 0, M - 1),
                                     rng.integers(0, M, 4096), np.where(lens == 0)[0][:64], [0, M - 1]]))
    srp, sci, sva = _sub_csr(rp, ci, va, rows)
    ref = oracle.spmm_f64(srp, sci, sva, (len(rows), K), B)
    got = C[torch.from_numpy(rows).to("cuda:0")].cpu().numpy()
    assert oracle.rel_err(got, ref) <= TOL
    # empty rows are exact zeros everywhere
    empty = torch.from_numpy(np.where(lens == 0)[0]).to("cuda:0")
    assert not bool(C[empty].any())
    # backward of the same operand on the sampled rows' transposed problem: grad_B column sums
    G = torch.ones(M, N, device="cuda:0")
    dB = plan.backward_B(G, torch.empty(K, N, device="cuda:0"))
    torch.cuda.synchronize()
    colsum = np.bincount(ci, weights=va.astype(np.float64), minlength=K)
    assert oracle.rel_err(dB[:, 0].cpu().numpy(), colsum) <= TOL
    plan.close()
