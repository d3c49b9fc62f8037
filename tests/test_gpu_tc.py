"""Parity of the block-dense tensor-core path (csrc/tc_gemm.cu: tcgen05.mma over densified 128-row tiles,
split-precision hi + lo operands) against the fp64 oracle, through the C ABI's extended forward
(b200_plan_forward_ex / b200_plan_backward_B_ex).  Runs on the B200 box only (`-m gpu`).

Tolerance (BASELINE.json north_star): rel-err = max|x - ref64| / max|ref64| <= 1e-5."""
import numpy as np
import pytest

from matrix_multiplication_b200 import capi
from matrix_multiplication_b200._capi import (OP_B_TRANSPOSED, OP_C_TRANSPOSED, OP_FORCE_TC, OP_NO_TC, OP_RELU)
from oracle import oracle
from conftest import random_csr

pytestmark = pytest.mark.gpu
TOL = 1e-5
KINDS = [0, 1]  # 0: tf32 hi + lo, 1: fp16 hi + lo with row / column scaling


@pytest.fixture(autouse=True)
def _reset_options():
    yield
    for k, v in (("tc_kind", capi.TC_KIND_DEFAULT), ("tc_nt", 0), ("tc_chunk", 8), ("tc_terms", 7), ("tc_mode", 1),
                 ("tc_sddmm_short", 1), ("tc_pair", 2)):
        capi.set_option(k, v)


def _dev(*arrays):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0") for a in arrays]


def _forward(rp, ci, va, shape, B, flags=OP_FORCE_TC, bias=None, bt=False, ct=False):
    import torch
    M, K = shape
    N = B.shape[1]
    crow, col, val = _dev(rp, ci, va)
    plan = capi.Plan(crow, col, val, shape)
    Bd, = _dev(B.T if bt else B)
    Cd = torch.full((N, M) if ct else (M, N), float("nan"), device="cuda:0")
    bd = None if bias is None else _dev(bias)[0]
    f = flags | (OP_B_TRANSPOSED if bt else 0) | (OP_C_TRANSPOSED if ct else 0)
    plan.forward_ex(Bd, Cd, bias=bd, flags=f)
    torch.cuda.synchronize()
    assert plan.tc_timeout() == 0, "a pipeline wait of the tensor-core kernel timed out"
    info = plan.tc_info()
    plan.close()
    C = Cd.cpu().numpy()
    return (C.T if ct else C), info


TC_CASES = [
    # M, K, N, density, heavy rows, empty rows
    (1, 1, 8, 1.0, (), ()),
    (7, 5, 9, 0.5, (), (2,)),
    (128, 32, 32, 0.3, (), ()),
    (129, 33, 65, 0.3, (0,), (128,)),
    (300, 200, 100, 0.3, (5,), (0, 299)),
    (1000, 1000, 777, 0.1, (), ()),
    (513, 777, 257, 0.05, (100,), (511, 512)),
    (64, 4100, 300, 0.2, (), ()),
    (2000, 70, 64, 0.5, (), tuple(range(0, 2000, 2))),
]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("case", TC_CASES, ids=[f"{c[0]}x{c[1]}x{c[2]}" for c in TC_CASES])
def test_tc_forward_edge_cases(case, kind):
    M, K, N, dens, heavy, empty = case
    rng = np.random.default_rng(M * 31 + N)
    rp, ci, va, _ = random_csr(rng, M, K, dens, heavy_rows=heavy, empty_rows=empty)
    B = rng.standard_normal((K, N)).astype(np.float32)
    capi.set_option("tc_kind", kind)
    n0 = capi.kernel_launch_count()
    C, info = _forward(rp, ci, va, (M, K), B)
    assert capi.kernel_launch_count() > n0 and info["last_path"] == 1 and info["tiles"] > 0
    assert oracle.rel_err(C, oracle.spmm_f64(rp, ci, va, (M, K), B)) <= TOL
    for r in empty:
        assert not C[r].any(), "empty rows must produce exact zeros"


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("nt", [64, 128, 192, 256])
def test_tc_tile_widths(nt, kind):
    rng = np.random.default_rng(nt)
    M, K, N = 400, 500, 600
    rp, ci, va, _ = random_csr(rng, M, K, 0.2)
    B = rng.standard_normal((K, N)).astype(np.float32)
    capi.set_option("tc_kind", kind)
    capi.set_option("tc_nt", nt)
    C, _ = _forward(rp, ci, va, (M, K), B)
    assert oracle.rel_err(C, oracle.spmm_f64(rp, ci, va, (M, K), B)) <= TOL


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("bt,ct", [(False, False), (True, False), (False, True), (True, True)])
def test_tc_transposed_operands_bias_relu(bt, ct, kind):
    """The FC-layer form: activations given as [tokens, features] (B^T), output written as [tokens, out] (C^T),
    bias per output feature and ReLU fused (replaces benchmarks/cusparse_fc_layer.py:41-46)."""
    rng = np.random.default_rng(5)
    M, K, N = 300, 260, 200
    rp, ci, va, _ = random_csr(rng, M, K, 0.15)
    B = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(M).astype(np.float32)
    capi.set_option("tc_kind", kind)
    C, _ = _forward(rp, ci, va, (M, K), B, flags=OP_FORCE_TC | OP_RELU, bias=bias, bt=bt, ct=ct)
    ref = np.maximum(oracle.spmm_f64(rp, ci, va, (M, K), B) + bias[:, None].astype(np.float64), 0.0)
    assert oracle.rel_err(C, ref) <= TOL
    # the same call off the tiles must agree: gather kernels with the epilogue fused into their stores (path 0), or
    # the plain CSR-order kernel when an operand is transposed (path 2)
    C2, info = _forward(rp, ci, va, (M, K), B, flags=OP_NO_TC | OP_RELU, bias=bias, bt=bt, ct=ct)
    assert info["last_path"] == (2 if (bt or ct) else 0)
    assert oracle.rel_err(C2, ref) <= TOL


@pytest.mark.parametrize("kind", KINDS)
def test_tc_long_contraction_of_positive_values(kind):
    """The tensor core adds into TMEM with truncation: without the chunked fp32 accumulation in registers a
    K = 16384 contraction of positive values comes out 1.3e-4 low (measured).  With it the error stays ~1e-6."""
    rng = np.random.default_rng(9)
    M, K, N = 128, 16384, 64
    rp, ci, va, _ = random_csr(rng, M, K, 0.5, signed=False)
    B = rng.random((K, N)).astype(np.float32)
    capi.set_option("tc_kind", kind)
    C, _ = _forward(rp, ci, va, (M, K), B)
    assert oracle.rel_err(C, oracle.spmm_f64(rp, ci, va, (M, K), B)) <= TOL


@pytest.mark.parametrize("kind", KINDS)
def test_tc_wide_dynamic_range(kind):
    """Rows of A and columns of B spanning 2^+-40, plus a huge-times-tiny pairing inside one contraction: the fp16
    kind scales per row / per column, the tf32 kind keeps fp32's exponent range."""
    rng = np.random.default_rng(10)
    M, K, N = 256, 300, 128
    rp, ci, va, dense = random_csr(rng, M, K, 0.2)
    B = rng.standard_normal((K, N)).astype(np.float32)
    rs = np.exp2(np.linspace(-40, 40, M)).astype(np.float32)
    cs = np.exp2(np.linspace(30, -30, N)).astype(np.float32)
    rows = np.repeat(np.arange(M), np.diff(rp))
    va2 = (va * rs[rows]).astype(np.float32)
    B2 = (B * cs[None, :]).astype(np.float32)
    capi.set_option("tc_kind", kind)
    C, _ = _forward(rp, ci, va2, (M, K), B2)
    ref = oracle.spmm_f64(rp, ci, va2, (M, K), B2)
    # element-wise against the magnitude each element is summed from (row scale x column scale)
    scale = np.abs(ref / (rs[:, None].astype(np.float64) * cs[None, :].astype(np.float64))).max()
    err = np.abs(C.astype(np.float64) - ref) / (rs[:, None].astype(np.float64) * cs[None, :].astype(np.float64))
    assert err.max() / scale <= TOL
    va3 = va.copy()
    va3[ci == 17] *= 1e6
    B3 = B.copy()
    B3[17, :] *= 1e-6
    C, _ = _forward(rp, ci, va3, (M, K), B3)
    assert oracle.rel_err(C, oracle.spmm_f64(rp, ci, va3, (M, K), B3)) <= TOL


@pytest.mark.parametrize("kind", KINDS)
def test_tc_nonfinite_dense_operand_falls_back_on_the_device(kind):
    """0 * Inf = NaN in a dense contraction; the sparse product skips the term.  The packer flags Inf/NaN in B and the
    same call then runs in CSR order: identical NaN/Inf pattern to the gather kernels, finite entries within tolerance."""
    rng = np.random.default_rng(12)
    M, K, N = 500, 400, 96
    rp, ci, va, _ = random_csr(rng, M, K, 0.1)
    B = rng.standard_normal((K, N)).astype(np.float32)
    B[5, 7] = np.inf
    B[9, 3] = np.nan
    capi.set_option("tc_kind", kind)
    C, info = _forward(rp, ci, va, (M, K), B)
    assert info["last_path"] == 1
    Cg, _ = _forward(rp, ci, va, (M, K), B, flags=OP_NO_TC)
    assert np.array_equal(np.isnan(C), np.isnan(Cg)) and np.array_equal(np.isinf(C), np.isinf(Cg))
    fin = np.isfinite(Cg)
    assert fin.sum() > 0.9 * fin.size  # only the rows touching k = 5 / k = 9 in columns 7 / 3
    assert np.abs(C[fin] - Cg[fin]).max() <= 1e-5 * np.abs(Cg[fin]).max()


@pytest.mark.parametrize("kind", KINDS)
def test_tc_block_sparse_tiles_are_skipped(kind):
    rng = np.random.default_rng(13)
    M, K, N = 1000, 1000, 128
    rp, ci, va, dense = random_csr(rng, M, K, 0.1)
    dense[128:384] = 0
    dense[:, 64:256] = 0
    rp, ci, va = oracle.dense_to_csr(dense)
    B = rng.standard_normal((K, N)).astype(np.float32)
    capi.set_option("tc_kind", kind)
    C, info = _forward(rp, ci, va, (M, K), B)
    bk = 64 if kind == 1 else 32
    occupied = {(r // 128, c // bk) for r, c in zip(np.repeat(np.arange(M), np.diff(rp)), ci)}
    assert info["tiles"] == len(occupied)
    assert oracle.rel_err(C, oracle.spmm_f64(rp, ci, va, (M, K), B)) <= TOL
    assert not C[128:384].any()


@pytest.mark.parametrize("kind", KINDS)
def test_tc_plan_backward_and_value_refresh(kind):
    import torch
    rng = np.random.default_rng(14)
    M, K, N = 700, 500, 160
    rp, ci, va, _ = random_csr(rng, M, K, 0.12)
    crow, col, val = _dev(rp, ci, va)
    B, G = _dev(rng.standard_normal((K, N)).astype(np.float32), rng.standard_normal((M, N)).astype(np.float32))
    capi.set_option("tc_kind", kind)
    capi.set_option("tc_mode", 2)
    plan = capi.Plan(crow, col, val, (M, K))
    C = plan.forward(B, torch.empty(M, N, device="cuda:0"))          # the plain entry points take the same path
    dB = plan.backward_B(G, torch.empty(K, N, device="cuda:0"))
    torch.cuda.synchronize()
    assert plan.tc_info(0)["tiles"] > 0 and plan.tc_info(1)["tiles"] > 0 and plan.tc_info()["last_path"] == 1
    assert oracle.rel_err(C.cpu().numpy(), oracle.spmm_f64(rp, ci, va, (M, K), B.cpu().numpy())) <= TOL
    assert oracle.rel_err(dB.cpu().numpy(), oracle.spmm_t_f64(rp, ci, va, (M, K), G.cpu().numpy())) <= TOL
    # transposed gradient in, transposed result out
    dBt = plan.backward_B_ex(G.t().contiguous(), torch.empty(N, K, device="cuda:0"),
                             flags=OP_B_TRANSPOSED | OP_C_TRANSPOSED)
    torch.cuda.synchronize()
    assert oracle.rel_err(dBt.t().cpu().numpy(), oracle.spmm_t_f64(rp, ci, va, (M, K), G.cpu().numpy())) <= TOL
    # new values on the same pattern: tiles of both sides are re-packed
    va2 = (va * 3 + 1).astype(np.float32)
    val.copy_(torch.from_numpy(va2))
    plan.values_updated()
    C = plan.forward(B, torch.empty(M, N, device="cuda:0"))
    dB = plan.backward_B(G, torch.empty(K, N, device="cuda:0"))
    torch.cuda.synchronize()
    assert oracle.rel_err(C.cpu().numpy(), oracle.spmm_f64(rp, ci, va2, (M, K), B.cpu().numpy())) <= TOL
    assert oracle.rel_err(dB.cpu().numpy(), oracle.spmm_t_f64(rp, ci, va2, (M, K), G.cpu().numpy())) <= TOL
    plan.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("case", [(128, 64, 64, 0.3), (300, 260, 200, 0.15), (1000, 777, 500, 0.1), (129, 2100, 70, 0.2),
                                  (640, 96, 4100, 0.25)])
def test_tc_sddmm_on_the_pattern(case, kind):
    """grad of the CSR values on the tensor cores (D = G B^T, contraction over the dense columns, sampled on A's
    pattern by the epilogue), operands as [M, N] / [K, N] and transposed as [N, M] / [N, K]."""
    import torch
    M, K, N, dens = case
    rng = np.random.default_rng(M + N)
    rp, ci, va, _ = random_csr(rng, M, K, dens, heavy_rows=(0,), empty_rows=(1,))
    G = rng.standard_normal((M, N)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    ref = oracle.sddmm_f64(rp, ci, (M, K), G, B)
    crow, col, val, Gd, Bd = _dev(rp, ci, va, G, B)
    capi.set_option("tc_kind", kind)
    plan = capi.Plan(crow, col, val, (M, K))
    out = plan.backward_A_ex(Gd, Bd, torch.full((len(va),), float("nan"), device="cuda:0"), flags=OP_FORCE_TC)
    torch.cuda.synchronize()
    assert plan.tc_timeout() == 0 and plan.tc_info()["last_path"] == 1
    assert oracle.rel_err(out.cpu().numpy(), ref) <= TOL
    out_t = plan.backward_A_ex(Gd.t().contiguous(), Bd.t().contiguous(),
                               torch.full((len(va),), float("nan"), device="cuda:0"), flags=OP_FORCE_TC | OP_B_TRANSPOSED)
    torch.cuda.synchronize()
    assert oracle.rel_err(out_t.cpu().numpy(), ref) <= TOL
    # Inf / NaN in an operand: the launch runs in CSR order and propagates them like the gather SDDMM
    G2 = Gd.clone()
    G2[3, 5] = float("inf")
    o1 = plan.backward_A_ex(G2, Bd, torch.empty(len(va), device="cuda:0"), flags=OP_FORCE_TC)
    o2 = plan.backward_A_ex(G2, Bd, torch.empty(len(va), device="cuda:0"), flags=OP_NO_TC)
    torch.cuda.synchronize()
    assert torch.equal(torch.isfinite(o1), torch.isfinite(o2))
    fin = torch.isfinite(o2)
    assert float((o1[fin] - o2[fin]).abs().max()) <= 1e-5 * float(o2[fin].abs().max())
    plan.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("mode", [2, 3, 0])
@pytest.mark.parametrize("case", [(1, 33, 40, 1.0), (300, 500, 40, 0.2), (777, 3001, 100, 0.05), (5000, 31, 9, 0.7),
                                  (1030, 1030, 256, 0.5)])
def test_tc_sddmm_short_contraction_kernels(case, mode, kind):
    """Short contractions (one TMEM accumulation): the transposed-product kernel with 128-row tiles (mode 2) and
    256-row tiles (mode 3), and the general kernel (mode 0), on ragged shapes — same entries, fp64 oracle."""
    import torch
    M, K, N, dens = case
    rng = np.random.default_rng(M * 3 + N)
    rp, ci, va, _ = random_csr(rng, M, K, dens, heavy_rows=(0,), empty_rows=(1,) if M > 2 else ())
    G = rng.standard_normal((M, N)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    ref = oracle.sddmm_f64(rp, ci, (M, K), G, B)
    crow, col, val, Gd, Bd = _dev(rp, ci, va, G, B)
    capi.set_option("tc_kind", kind)
    capi.set_option("tc_sddmm_short", mode)
    plan = capi.Plan(crow, col, val, (M, K))
    for tr in (False, True):
        g, b = (Gd.t().contiguous(), Bd.t().contiguous()) if tr else (Gd, Bd)
        out = plan.backward_A_ex(g, b, torch.full((len(va),), float("nan"), device="cuda:0"),
                                 flags=OP_FORCE_TC | (OP_B_TRANSPOSED if tr else 0))
        torch.cuda.synchronize()
        assert plan.tc_timeout() == 0 and plan.tc_info()["last_path"] == 1
        short = (N + (63 if kind == 1 else 31)) // (64 if kind == 1 else 32) <= 8
        want = {2: (3,), 3: (4,), 0: (1, 2)}[mode] if short else (1, 2)  # (2: the CTA-pair form of the general kernel)
        assert capi.get_option("tc_last_kernel") in want
        assert oracle.rel_err(out.cpu().numpy(), ref) <= TOL
    plan.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("case", [(129, 33, 65, 0.3, (0,), (128,)), (300, 200, 100, 0.3, (5,), (0, 299)),
                                  (513, 777, 257, 0.05, (100,), (511, 512)), (1000, 1000, 777, 0.1, (), ())])
def test_tc_cta_pair_kernel(case, kind):
    """The cta_group::2 form of the tile kernel (two CTAs of a cluster share every B tile; the default for 256-wide
    tiles on long contractions) forced on small ragged operands: forward, transposed operands, bias + ReLU."""
    M, K, N, dens, heavy, empty = case
    rng = np.random.default_rng(M + K)
    rp, ci, va, _ = random_csr(rng, M, K, dens, heavy_rows=heavy, empty_rows=empty)
    B = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(M).astype(np.float32)
    capi.set_option("tc_kind", kind)
    capi.set_option("tc_pair", 1)
    ref = oracle.spmm_f64(rp, ci, va, (M, K), B)
    C, info = _forward(rp, ci, va, (M, K), B)
    assert info["last_path"] == 1 and capi.get_option("tc_last_kernel") == 2
    assert oracle.rel_err(C, ref) <= TOL
    Ct, _ = _forward(rp, ci, va, (M, K), B, bias=bias, bt=True, ct=True, flags=OP_FORCE_TC | OP_RELU)
    assert capi.get_option("tc_last_kernel") == 2
    assert oracle.rel_err(Ct, np.maximum(ref + bias[:, None], 0.0)) <= TOL


def test_tc_auto_rule_picks_paths():
    """Auto mode: large, dense-enough operands take the tensor-core tiles; sparse, small (launch-bound) or very narrow
    ones the gather kernels (crossover rule: a cost model of both paths)."""
    import torch
    rng = np.random.default_rng(15)
    for S, dens, N, want in ((2048, 0.10, 512, 1), (2048, 0.01, 512, 0), (512, 0.10, 128, 0), (2048, 0.10, 4, 0)):
        rp, ci, va, _ = random_csr(rng, S, S, dens)
        crow, col, val = _dev(rp, ci, va)
        B, = _dev(rng.standard_normal((S, N)).astype(np.float32))
        plan = capi.Plan(crow, col, val, (S, S))
        C = plan.forward(B, torch.empty(S, N, device="cuda:0"))
        torch.cuda.synchronize()
        assert plan.tc_info()["last_path"] == want, (S, dens, N)
        assert plan.would_use_tc(N) == bool(want)
        assert oracle.rel_err(C.cpu().numpy(), oracle.spmm_f64(rp, ci, va, (S, S), B.cpu().numpy())) <= TOL
        plan.close()


def test_tc_deterministic_and_leading_dimensions():
    import torch
    rng = np.random.default_rng(16)
    M, K, N = 384, 320, 200
    rp, ci, va, _ = random_csr(rng, M, K, 0.2)
    crow, col, val = _dev(rp, ci, va)
    Bbig, = _dev(rng.standard_normal((K, 232)).astype(np.float32))
    plan = capi.Plan(crow, col, val, (M, K))
    outs = []
    for _ in range(2):
        Cbig = torch.full((M, 260), -7.0, device="cuda:0")
        plan.forward_ex(Bbig[:, 9:9 + N], Cbig[:, 17:17 + N], flags=OP_FORCE_TC)  # unaligned windows: scalar paths
        torch.cuda.synchronize()
        assert (Cbig[:, :17] == -7.0).all() and (Cbig[:, 17 + N:] == -7.0).all(), "wrote outside the C window"
        outs.append(Cbig[:, 17:17 + N].cpu().numpy())
    np.testing.assert_array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert oracle.rel_err(outs[0], oracle.spmm_f64(rp, ci, va, (M, K), Bbig[:, 9:9 + N].cpu().numpy())) <= TOL
    plan.close()
