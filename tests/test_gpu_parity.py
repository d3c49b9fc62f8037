"""Parity of the CUDA path against the CPU oracle, through the C ABI (ctypes, host and device
entry points).  Runs on the B200 box only (`-m gpu`).

Tolerances (BASELINE.json north_star): SpMM / SDDMM rel-err = max|x - ref64| / max|ref64| <= 1e-5
against the fp64 oracle; CSR construction and transposition bit-exact."""
import numpy as np
import pytest

from matrix_multiplication_b200 import capi, synthetic
from oracle import oracle
from conftest import load_golden, random_csr

pytestmark = pytest.mark.gpu
TOL = 1e-5


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(autouse=True)
def _reset_options():
    yield
    capi.set_option("stage_mode", 1)
    capi.set_option("items_per_task", 0)


def test_smoke_kernel_reports_b200():
    assert capi.smoke() >= 100  # 148 SMs on B200


# ---- forward ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [f"fwd_{i}" for i in range(5)])
@pytest.mark.parametrize("stage", [1, 0])
def test_forward_golden(name, stage):
    g = load_golden(name)
    shape = tuple(int(x) for x in g["shape"])
    capi.set_option("stage_mode", stage)
    n0 = capi.kernel_launch_count()
    C = capi.spmm_host(g["row_ptr"], g["col"], g["val"], shape, g["B"])
    assert capi.kernel_launch_count() > n0
    ref = oracle.spmm_f64(g["row_ptr"], g["col"], g["val"], shape, g["B"])
    assert oracle.rel_err(C, ref) <= TOL
    assert np.allclose(C, g["C"], rtol=1e-5, atol=1e-8)  # the reference tests' own criterion


CASES = [
    # M, K, N, density, heavy rows, empty rows
    (1, 1, 1, 1.0, (), ()),
    (7, 5, 3, 0.5, (), (2,)),
    (64, 64, 4, 0.1, (), ()),
    (100, 3000, 8, 0.01, (0, 50, 99), (1, 2, 3, 98)),
    (257, 513, 17, 0.05, (128,), (0, 256)),
    (300, 20000, 32, 0.002, (5, 6, 7), ()),
    (1000, 1000, 64, 0.1, (), ()),
    (513, 777, 65, 0.03, (100,), (511, 512)),
    (2000, 500, 128, 0.02, (), tuple(range(0, 2000, 2))),
    (300, 300, 130, 0.3, (), ()),
    (50, 40000, 256, 0.0005, (10, 11), ()),
    (400, 400, 300, 0.05, (), ()),
    (128, 128, 2, 0.2, (), ()),
    (128, 128, 6, 0.2, (), ()),
    (33, 65, 1000, 0.1, (3,), ()),
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}x{c[1]}x{c[2]}" for c in CASES])
@pytest.mark.parametrize("stage", [1, 0])
def test_forward_edge_cases(case, stage):
    M, K, N, dens, heavy, empty = case
    rng = np.random.default_rng(M * 31 + N)
    rp, ci, va, _ = random_csr(rng, M, K, dens, heavy_rows=heavy, empty_rows=empty)
    B = rng.standard_normal((K, N)).astype(np.float32)
    capi.set_option("stage_mode", stage)
    C = capi.spmm_host(rp, ci, va, (M, K), B)
    ref = oracle.spmm_f64(rp, ci, va, (M, K), B)
    assert oracle.rel_err(C, ref) <= TOL
    for r in empty:
        assert not C[r].any(), "empty rows must produce exact zeros"


@pytest.mark.parametrize("W", [32, 64, 512, 2048])
def test_forward_task_sizes(W):
    rng = np.random.default_rng(W)
    M, K, N = 500, 9000, 48
    rp, ci, va, _ = random_csr(rng, M, K, 0.004, heavy_rows=(0, 250, 251, 499), empty_rows=(1, 2, 498))
    B = rng.standard_normal((K, N)).astype(np.float32)
    capi.set_option("items_per_task", W)
    C = capi.spmm_host(rp, ci, va, (M, K), B)
    assert oracle.rel_err(C, oracle.spmm_f64(rp, ci, va, (M, K), B)) <= TOL


def test_forward_all_zero_matrix_and_empty_shapes():
    rp = np.zeros(11, dtype=np.int32)
    B = np.ones((6, 5), dtype=np.float32)
    C = capi.spmm_host(rp, np.zeros(0, np.int32), np.zeros(0, np.float32), (10, 6), B)
    assert C.shape == (10, 5) and not C.any()
    C = capi.spmm_host(np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), (0, 6), B)
    assert C.shape == (0, 5)


def test_forward_special_values_propagate():
    rp = np.array([0, 2, 3, 3], dtype=np.int32)
    ci = np.array([0, 1, 1], dtype=np.int32)
    va = np.array([1.0, np.inf, np.nan], dtype=np.float32)
    B = np.array([[1, 2, 3, 4], [1, 1, 1, 1]], dtype=np.float32)
    C = capi.spmm_host(rp, ci, va, (3, 2), B)
    assert np.isinf(C[0]).all() and np.isnan(C[1]).all() and not C[2].any()


def test_forward_is_deterministic():
    rng = np.random.default_rng(11)
    rp, ci, va, _ = random_csr(rng, 800, 5000, 0.01, heavy_rows=(3, 400))
    B = rng.standard_normal((5000, 64)).astype(np.float32)
    a = capi.spmm_host(rp, ci, va, (800, 5000), B)
    b = capi.spmm_host(rp, ci, va, (800, 5000), B)
    np.testing.assert_array_equal(bits(a), bits(b))


def test_forward_matches_csr_order_fp32_when_rows_are_not_split():
    """Light rows are accumulated in CSR order with fmaf: bit-identical to the naive kernel's
    arithmetic (oracle_spmm_f32 fused, reference src/naive_sparse_mm.cu:84-90)."""
    rng = np.random.default_rng(12)
    rp, ci, va, _ = random_csr(rng, 300, 400, 0.05)
    B = rng.standard_normal((400, 128)).astype(np.float32)  # N >= 128: one lane group per row
    capi.set_option("items_per_task", 256)  # no row is longer than a task: nothing is split
    C = capi.spmm_host(rp, ci, va, (300, 400), B)
    np.testing.assert_array_equal(bits(C), bits(oracle.spmm_f32(rp, ci, va, (300, 400), B, fused=True)))


# ---- narrow N over shared-memory slabs of B (spmm_slab.cu) -----------------------------------------------------
@pytest.mark.parametrize("case", [(4096, 8192, 16, 0.05), (3000, 15000, 16, 0.05), (5000, 5000, 8, 0.05),
                                  (200000, 4096, 4, 0.015), (16384, 1500, 16, 0.08)])
def test_forward_narrow_n_slab_kernel(case):
    """Plans route N in {4, 8, 16} with a B that fits a few shared-memory slabs to spmm_slab_kernel (path 3): K cut
    into slabs, rows owned by persistent CTAs (more rows than one accumulation block holds in the 200000-row case),
    running sums in shared memory.  Against the fp64 oracle, with bias + ReLU fused, and bit-identical when repeated."""
    import torch
    M, K, N, dens = case
    rng = np.random.default_rng(M + N)
    nnz_target = int(M * K * dens)
    rows = np.sort(rng.integers(0, M, nnz_target))
    cols = rng.integers(0, K, nnz_target)
    key = np.unique(rows.astype(np.int64) * K + cols)
    rows, cols = (key // K), (key % K).astype(np.int32)
    rows[rows == 7] = 8  # row 7 empty (keys stay sorted: duplicates of row 8 are fine for this purpose? no: re-sort)
    key = np.unique(rows.astype(np.int64) * K + cols)
    rows, cols = (key // K), (key % K).astype(np.int32)
    rp = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=rp[1:])
    rp = rp.astype(np.int32)
    va = rng.standard_normal(len(cols)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(M).astype(np.float32)
    dev = torch.device("cuda:0")
    crow, col, val, Bd, bd = (torch.from_numpy(x).to(dev) for x in (rp, cols, va, B, bias))
    plan = capi.Plan(crow, col, val, (M, K))
    C = plan.forward_ex(Bd, torch.full((M, N), float("nan"), device=dev), flags=capi.OP_NO_TC)
    torch.cuda.synchronize()
    assert plan.tc_info()["last_path"] == 3, "expected the shared-memory slab kernel"
    ref = oracle.spmm_f64(rp, cols, va, (M, K), B)
    assert oracle.rel_err(C.cpu().numpy(), ref) <= TOL
    assert not bool(C[7].any()), "empty rows must produce exact zeros"
    C2 = plan.forward_ex(Bd, torch.empty(M, N, device=dev), flags=capi.OP_NO_TC)
    assert torch.equal(C, C2)
    Cb = plan.forward_ex(Bd, torch.empty(M, N, device=dev), bias=bd, flags=capi.OP_NO_TC | capi.OP_RELU)
    torch.cuda.synchronize()
    assert oracle.rel_err(Cb.cpu().numpy(), np.maximum(ref + bias[:, None].astype(np.float64), 0.0)) <= TOL
    # values changed in place: the plan's sliced copy of A is rewritten after b200_plan_values_updated
    val.mul_(-2.0)
    plan.values_updated()
    Cn = plan.forward_ex(Bd, torch.empty(M, N, device=dev), flags=capi.OP_NO_TC)
    torch.cuda.synchronize()
    assert plan.tc_info()["last_path"] == 3
    assert oracle.rel_err(Cn.cpu().numpy(), -2.0 * ref) <= TOL
    val.mul_(-0.5)
    plan.values_updated()
    try:
        capi.set_option("slab_mode", 2)  # the first slab kernel (CSR walked in place)
        C1 = plan.forward_ex(Bd, torch.empty(M, N, device=dev), flags=capi.OP_NO_TC)
        torch.cuda.synchronize()
        assert plan.tc_info()["last_path"] == 3
        assert oracle.rel_err(C1.cpu().numpy(), ref) <= TOL
        capi.set_option("slab_mode", 0)
        Cg = plan.forward_ex(Bd, torch.empty(M, N, device=dev), flags=capi.OP_NO_TC)
        torch.cuda.synchronize()
        assert plan.tc_info()["last_path"] == 0
        assert oracle.rel_err(Cg.cpu().numpy(), ref) <= TOL
    finally:
        capi.set_option("slab_mode", 1)
    plan.close()


# ---- dense -> CSR (bit-exact) --------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["csr_special", "csr_ragged", "csr_sparse", "csr_dense", "csr_allzero",
                                  "fwd_0", "fwd_3", "bwd_0"])
def test_dense_to_csr_golden(name):
    g = load_golden(name)
    rp, ci, va = capi.dense_to_csr_host(g["A_dense"])
    np.testing.assert_array_equal(rp, g["row_ptr"])
    np.testing.assert_array_equal(ci, g["col"])
    np.testing.assert_array_equal(bits(va), bits(g["val"]))


@pytest.mark.parametrize("shape", [(1, 1), (3, 129), (1000, 7), (513, 1023), (64, 4096), (5000, 33)])
def test_dense_to_csr_random(shape):
    rng = np.random.default_rng(shape[0])
    d = (rng.standard_normal(shape) * (rng.random(shape) < 0.3)).astype(np.float32)
    d.flat[:: 17] = -0.0
    d.flat[5 % d.size] = np.nan
    d.flat[3 % d.size] = np.float32(1e-42)  # denormal: kept
    rp, ci, va = capi.dense_to_csr_host(d)
    orp, oci, ova = oracle.dense_to_csr(d)
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(ci, oci)
    np.testing.assert_array_equal(bits(va), bits(ova))


@pytest.mark.parametrize("shape,dens", [((1, 64), 0.5), ((3, 129), 0.3), ((40, 512), 0.1), ((37, 513), 0.1),
                                        ((300, 2000), 0.02), ((513, 1023), 0.3), ((64, 4096), 0.9), ((5000, 65), 0.2),
                                        ((2048, 3000), 0.001), ((100, 70000), 0.05)])
def test_dense_to_csr_one_pass(shape, dens):
    """b200_dense_to_csr_onepass (count + decoupled look-back + fill in one kernel): bit-exact against the oracle on
    ragged widths, rows spanning several chunks / tiles, empty and full rows, special values."""
    import torch
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    d = (rng.standard_normal(shape) * (rng.random(shape) < dens)).astype(np.float32)
    d.flat[:: 17] = -0.0
    d.flat[5 % d.size] = np.nan
    d.flat[3 % d.size] = np.float32(1e-42)  # denormal: kept
    if shape[0] > 2:
        d[1] = 0      # an empty row
        d[2] = 1.5    # a full row
    orp, oci, ova = oracle.dense_to_csr(d)
    dd = torch.from_numpy(d).to("cuda:0")
    crow, col, val, nnz = capi.dense_to_csr_onepass_device(dd)
    assert nnz == len(oci)
    np.testing.assert_array_equal(crow.cpu().numpy(), orp)
    np.testing.assert_array_equal(col[:nnz].cpu().numpy(), oci)
    np.testing.assert_array_equal(bits(val[:nnz].cpu().numpy()), bits(ova))
    # a strided window (leading dimension > K)
    if shape[1] % 4 == 0 and shape[1] >= 64:
        big = torch.zeros(shape[0], shape[1] + 8, device="cuda:0")
        big[:, :shape[1]] = dd
        crow2, col2, val2, nnz2 = capi.dense_to_csr_onepass_device(big[:, :shape[1]])
        assert nnz2 == nnz and torch.equal(crow2, crow) and torch.equal(col2[:nnz], col[:nnz])
    # too little room: reported, row_ptr complete, the second phase finishes the job
    if nnz > 8:
        with pytest.raises(capi.B200Error) as e:
            capi.dense_to_csr_onepass_device(dd, capacity=nnz // 2)
        assert e.value.status == capi.ERR_INVALID and e.value.nnz == nnz
        np.testing.assert_array_equal(e.value.crow.cpu().numpy(), orp)


# ---- transpose (bit-exact) and backward ----------------------------------------------------------------
@pytest.mark.parametrize("case", [(1, 1, 1.0), (50, 70, 0.2), (700, 300, 0.05), (300, 5000, 0.01), (2000, 3, 0.5)])
def test_transpose_bit_exact(case):
    M, K, dens = case
    rng = np.random.default_rng(M + K)
    rp, ci, va, _ = random_csr(rng, M, K, dens)
    t = capi.transpose_host(rp, ci, va, (M, K))
    o = oracle.csr_transpose(rp, ci, va, (M, K))
    np.testing.assert_array_equal(t[0], o[0])
    np.testing.assert_array_equal(t[1], o[1])
    np.testing.assert_array_equal(bits(t[2]), bits(o[2]))


def test_backward_golden():
    g = load_golden("bwd_0")
    shape = tuple(int(x) for x in g["shape"])
    gb = capi.backward_B_host(g["row_ptr"], g["col"], g["val"], shape, g["G"])
    assert oracle.rel_err(gb, oracle.spmm_t_f64(g["row_ptr"], g["col"], g["val"], shape, g["G"])) <= TOL
    assert np.allclose(gb, g["grad_B"], rtol=1e-5, atol=1e-7)
    ga = capi.sddmm_host(g["row_ptr"], g["col"], shape, g["G"], g["B"])
    assert oracle.rel_err(ga, oracle.sddmm_f64(g["row_ptr"], g["col"], shape, g["G"], g["B"])) <= TOL


@pytest.mark.parametrize("case", [(64, 64, 4, 0.1), (257, 513, 17, 0.05), (1000, 800, 64, 0.05),
                                  (300, 20000, 128, 0.002), (400, 300, 300, 0.05), (50, 60, 1, 0.3),
                                  (90, 70, 6, 0.3), (128, 4000, 16, 0.02)])
def test_backward_random(case):
    M, K, N, dens = case
    rng = np.random.default_rng(N)
    rp, ci, va, _ = random_csr(rng, M, K, dens, heavy_rows=(0,), empty_rows=(1,))
    G = rng.standard_normal((M, N)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    gb = capi.backward_B_host(rp, ci, va, (M, K), G)
    assert oracle.rel_err(gb, oracle.spmm_t_f64(rp, ci, va, (M, K), G)) <= TOL
    ga = capi.sddmm_host(rp, ci, (M, K), G, B)
    assert oracle.rel_err(ga, oracle.sddmm_f64(rp, ci, (M, K), G, B)) <= TOL


# ---- device entry points: plans, strides, workspace ------------------------------------------------------
def test_plan_device_path_and_value_refresh():
    import torch
    rng = np.random.default_rng(21)
    M, K, N = 600, 900, 96
    rp, ci, va, _ = random_csr(rng, M, K, 0.03, heavy_rows=(7,))
    dev = torch.device("cuda:0")
    crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, ci, va))
    B = torch.from_numpy(rng.standard_normal((K, N)).astype(np.float32)).to(dev)
    G = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).to(dev)
    plan = capi.Plan(crow, col, val, (M, K))
    assert plan.stats()[0] == int(np.diff(rp).max())
    C = plan.forward(B, torch.empty(M, N, device=dev))
    dB = plan.backward_B(G, torch.empty(K, N, device=dev))
    dA = plan.backward_A(G, B, torch.empty(len(va), device=dev))
    torch.cuda.synchronize()
    assert oracle.rel_err(C.cpu().numpy(), oracle.spmm_f64(rp, ci, va, (M, K), B.cpu().numpy())) <= TOL
    assert oracle.rel_err(dB.cpu().numpy(), oracle.spmm_t_f64(rp, ci, va, (M, K), G.cpu().numpy())) <= TOL
    assert oracle.rel_err(dA.cpu().numpy(),
                          oracle.sddmm_f64(rp, ci, (M, K), G.cpu().numpy(), B.cpu().numpy())) <= TOL
    # new values on the same pattern: the cached transpose is re-gathered
    va2 = (va * 3 + 1).astype(np.float32)
    val.copy_(torch.from_numpy(va2))
    plan.values_updated()
    dB = plan.backward_B(G, torch.empty(K, N, device=dev))
    torch.cuda.synchronize()
    assert oracle.rel_err(dB.cpu().numpy(), oracle.spmm_t_f64(rp, ci, va2, (M, K), G.cpu().numpy())) <= TOL
    plan.close()


def test_leading_dimensions_and_non_default_stream():
    import torch
    rng = np.random.default_rng(22)
    M, K, N = 200, 300, 24
    rp, ci, va, _ = random_csr(rng, M, K, 0.05)
    dev = torch.device("cuda:0")
    crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, ci, va))
    Bbig = torch.from_numpy(rng.standard_normal((K, 40)).astype(np.float32)).to(dev)
    Cbig = torch.full((M, 56), -7.0, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        capi.spmm_device(crow, col, val, (M, K), Bbig[:, 8:8 + N], Cbig[:, 16:16 + N], ldb=40, ldc=56,
                         stream=s.cuda_stream)
    s.synchronize()
    ref = oracle.spmm_f64(rp, ci, va, (M, K), Bbig[:, 8:8 + N].cpu().numpy())
    assert oracle.rel_err(Cbig[:, 16:16 + N].cpu().numpy(), ref) <= TOL
    assert (Cbig[:, :16] == -7.0).all() and (Cbig[:, 16 + N:] == -7.0).all(), "wrote outside the C window"


def test_workspace_too_small_is_reported():
    import torch
    rng = np.random.default_rng(23)
    rp, ci, va, _ = random_csr(rng, 100, 100, 0.1)
    dev = torch.device("cuda:0")
    crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, ci, va))
    B = torch.zeros(100, 8, device=dev)
    with pytest.raises(capi.B200Error) as e:
        capi.spmm_device(crow, col, val, (100, 100), B, torch.empty(100, 8, device=dev),
                         workspace=torch.empty(16, dtype=torch.uint8, device=dev))
    assert e.value.status == capi.ERR_INVALID


# ---- BASELINE-size properties (sizes the fp64 oracle would take too long on) -----------------------------
def test_full_size_rmat_properties():
    """Config 4 at reduced scale 17 (131k rows, 2M nnz, N=128) + size-independent properties:
    rows sums (A * ones), linearity, and agreement with the oracle on a sample of rows."""
    import torch
    rp, ci, va, shape = synthetic.rmat_csr(scale=17, edge_factor=16, seed=0)
    M, K = shape
    N = 128
    dev = torch.device("cuda:0")
    crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, ci, va))
    plan = capi.Plan(crow, col, val, shape)
    ones = torch.ones(K, N, device=dev)
    C1 = plan.forward(ones, torch.empty(M, N, device=dev))
    rowsum = np.add.reduceat(np.concatenate([va.astype(np.float64), [0.0]]), np.minimum(rp[:-1], len(va)))
    rowsum[np.diff(rp) == 0] = 0.0
    torch.cuda.synchronize()
    assert oracle.rel_err(C1[:, 0].cpu().numpy(), rowsum) <= TOL
    assert torch.equal(C1[:, :1].expand(-1, N), C1)
    g = torch.Generator(device="cpu").manual_seed(0)
    B1 = torch.rand(K, N, generator=g).to(dev)
    B2 = torch.rand(K, N, generator=g).to(dev)
    Ca = plan.forward(B1, torch.empty(M, N, device=dev))
    Cb = plan.forward(B2, torch.empty(M, N, device=dev))
    Cs = plan.forward(B1 + B2, torch.empty(M, N, device=dev))
    torch.cuda.synchronize()
    assert float((Cs - (Ca + Cb)).abs().max() / Cs.abs().max()) <= 1e-5
    # a sample of rows (the heaviest, some light, some empty) against the fp64 oracle
    lens = np.diff(rp)
    rows = np.unique(np.concatenate([np.argsort(lens)[-8:], np.arange(0, M, M // 64), np.where(lens == 0)[0][:8]]))
    sub_rp = np.concatenate([[0], np.cumsum(lens[rows])]).astype(np.int32)
    idx = np.concatenate([np.arange(rp[r], rp[r + 1]) for r in rows])
    ref = oracle.spmm_f64(sub_rp, ci[idx], va[idx], (len(rows), K), B1.cpu().numpy())
    assert oracle.rel_err(Ca[torch.from_numpy(rows).to(dev)].cpu().numpy(), ref) <= TOL
    plan.close()


def test_peer_store_epilogue_on_one_gpu():
    """The multi-GPU epilogue with two stand-in 'peers' (extra buffers on the same GPU): every row —
    light, heavy (fix-up path) and empty (zero fill) — must land identically in all three copies."""
    import torch
    rng = np.random.default_rng(31)
    M, K, N = 700, 3000, 128
    rp, ci, va, _ = random_csr(rng, M, K, 0.01, heavy_rows=(5, 300), empty_rows=(0, 6, 699))
    dev = torch.device("cuda:0")
    crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, ci, va))
    B = torch.from_numpy(rng.standard_normal((K, N)).astype(np.float32)).to(dev)
    plan = capi.Plan(crow, col, val, (M, K))
    C0 = torch.full((M, N), 7.0, device=dev)
    P1 = torch.full((M, N), 8.0, device=dev)
    P2 = torch.full((M, N), 9.0, device=dev)
    plan.forward_peers(B, C0, [P1.data_ptr(), P2.data_ptr()])
    torch.cuda.synchronize()
    ref = oracle.spmm_f64(rp, ci, va, (M, K), B.cpu().numpy())
    assert oracle.rel_err(C0.cpu().numpy(), ref) <= TOL
    assert torch.equal(C0, P1) and torch.equal(C0, P2)
    plan.close()


# ---- against outputs of the REAL reference (recorded on a B200 by tests/golden/make_ref_golden.py) ----------
@pytest.mark.parametrize("i", range(5))
def test_engine_matches_recorded_reference_forward(i):
    g = load_golden(f"ref_fwd_{i}")
    shape = tuple(int(x) for x in g["shape"])
    rp, ci, va = capi.dense_to_csr_host(g["A_dense"])
    np.testing.assert_array_equal(rp, g["row_ptr"])     # bit-exact with the CSR the reference built
    np.testing.assert_array_equal(ci, g["col"])
    np.testing.assert_array_equal(bits(va), bits(g["val"]))
    C = capi.spmm_host(rp, ci, va, shape, g["B"])
    assert oracle.rel_err(C, g["C"].astype(np.float64)) <= TOL
    assert np.allclose(C, g["C"], rtol=1e-5, atol=1e-6)


def test_engine_matches_recorded_reference_backward():
    g = load_golden("ref_bwd_0")
    shape = tuple(int(x) for x in g["shape"])
    gb = capi.backward_B_host(g["row_ptr"], g["col"], g["val"], shape, g["G"])
    assert oracle.rel_err(gb, g["grad_B"].astype(np.float64)) <= TOL


@pytest.mark.parametrize("N,panels", [(128, 0), (128, 2), (96, 0), (40, 0), (256, 8)])
def test_plan_forward_host_pipeline(N, panels):
    """b200_plan_forward_host: column-panel pipeline with host buffers (pinned and pageable)."""
    import torch
    rng = np.random.default_rng(N)
    M, K = 900, 1100
    rp, ci, va, _ = random_csr(rng, M, K, 0.02, heavy_rows=(4,), empty_rows=(0, 899))
    dev = torch.device("cuda:0")
    crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, ci, va))
    plan = capi.Plan(crow, col, val, (M, K))
    B = rng.standard_normal((K, N)).astype(np.float32)
    ref = oracle.spmm_f64(rp, ci, va, (M, K), B)
    Bp = torch.from_numpy(B).pin_memory()
    Cp = torch.empty(M, N).pin_memory()
    plan.forward_host(Bp, Cp, panels)
    assert oracle.rel_err(Cp.numpy(), ref) <= TOL
    Cn = np.empty((M, N), dtype=np.float32)          # pageable numpy buffers work too
    plan.forward_host(B, Cn, panels)
    np.testing.assert_array_equal(bits(Cn), bits(Cp.numpy()))
    plan.close()


@pytest.mark.parametrize("M,K", [(3000, 250000), (250000, 3000)])
def test_plan_very_rectangular(M, K):
    """Row shards of a big matrix are very rectangular (few rows, all the columns): the task size — and with it
    the workspace — differs between the forward (M + nnz items) and the transposed product (K + nnz items)."""
    import torch
    rng = np.random.default_rng(M)
    nnz = 400000
    rows = np.sort(rng.integers(0, M, nnz))
    cols = rng.integers(0, K, nnz)
    key = np.unique(rows.astype(np.int64) * K + cols)
    rows, cols = (key // K).astype(np.int64), (key % K).astype(np.int32)
    rp = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=rp[1:])
    rp = rp.astype(np.int32)
    va = rng.standard_normal(len(cols)).astype(np.float32)
    N = 64
    dev = torch.device("cuda:0")
    crow, col, val = (torch.from_numpy(x).to(dev) for x in (rp, cols, va))
    B = torch.from_numpy(rng.standard_normal((K, N)).astype(np.float32)).to(dev)
    G = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).to(dev)
    plan = capi.Plan(crow, col, val, (M, K))
    C = plan.forward(B, torch.empty(M, N, device=dev))
    dB = plan.backward_B(G, torch.empty(K, N, device=dev))
    dA = plan.backward_A(G, B, torch.empty(len(va), device=dev))
    torch.cuda.synchronize()
    assert oracle.rel_err(C.cpu().numpy(), oracle.spmm_f64(rp, cols, va, (M, K), B.cpu().numpy())) <= TOL
    assert oracle.rel_err(dB.cpu().numpy(), oracle.spmm_t_f64(rp, cols, va, (M, K), G.cpu().numpy())) <= TOL
    assert oracle.rel_err(dA.cpu().numpy(), oracle.sddmm_f64(rp, cols, (M, K), G.cpu().numpy(), B.cpu().numpy())) <= TOL
    plan.close()
