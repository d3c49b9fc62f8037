"""CPU-only tests of the host side: the C-ABI library loads and exports what the header declares,
the merge-path cut and the row sharding behave, and the forward kernel's decomposition (simulated in
numpy on the library's own boundary function) reproduces the oracle.  No compute calls on a GPU."""
import numpy as np
import pytest

from matrix_multiplication_b200 import capi
from oracle import oracle
from conftest import random_csr
from taskwalk_sim import simulate


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    declared = capi.declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/spmm_b200.h but not exported"
        assert name in capi.SIGNATURES, f"{name} has no ctypes signature in _capi.py"
    assert L.b200_abi_version() == 1
    assert b"sm_100a" in L.b200_build_info()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rp = np.array([0, 1], dtype=np.int32)
    with pytest.raises(capi.B200Error) as e:
        capi.spmm_host(rp, np.zeros(1, np.int32), np.ones(1, np.float32), (1, 1), np.ones((1, 4), np.float32))
    assert e.value.status in (capi.ERR_NO_DEVICE, capi.ERR_CUDA)


def test_argument_validation():
    L = capi.lib()
    assert L.b200_spmm_csr_forward(None, None, None, -1, 1, 0, None, 1, 1, None, 1, None, 0, None) == capi.ERR_INVALID
    assert b"negative" in L.b200_last_error()
    assert L.b200_spmm_csr_forward(None, None, None, 4, 4, 0, None, 2, 8, None, 8, None, 0, None) == capi.ERR_INVALID
    assert L.b200_set_option(b"no_such_option", 1) == capi.ERR_INVALID
    capi.set_option("items_per_task", 64)
    assert capi.get_option("items_per_task") == 64
    capi.set_option("items_per_task", 0)
    # kernel-shape options of the tile path: defaults (auto) and clamping
    assert capi.get_option("tc_pair") == 2 and capi.get_option("tc_sddmm_short") == 1
    for name, values in (("tc_pair", (0, 1, 2)), ("tc_sddmm_short", (0, 2, 3, 1))):
        for v in values:
            capi.set_option(name, v)
            assert capi.get_option(name) == v
    capi.set_option("tc_pair", 7)
    assert capi.get_option("tc_pair") == 2
    assert L.b200_spmm_workspace_bytes(1000, 5000, 128) >= 256 + 2 * ((6000 + 127) // 128) * 128 * 4


def test_merge_path_boundary_properties():
    rng = np.random.default_rng(0)
    for M, K, dens, heavy_rows in [(50, 40, 0.1, ()), (200, 300, 0.02, (7, 150)), (30, 2000, 0.001, (3,)),
                                   (1, 500, 1.0, (0,)), (64, 8, 0.0, ())]:
        rp, ci, va, _ = random_csr(rng, M, K, dens, heavy_rows=heavy_rows)
        nnz = int(rp[-1])
        for W in (1, 7, 32, 256):
            n_tasks = max(1, -(-(M + nnz) // W))
            prev = (0, 0)
            assert capi.merge_path_boundary(rp, 0, W) == (0, 0)
            for t in range(1, n_tasks + 1):
                i, j = capi.merge_path_boundary(rp, t * W, W)
                assert prev[0] <= i <= M and prev[1] <= j <= nnz
                if i < M:
                    assert rp[i] <= j <= rp[i + 1]
                    if j > rp[i]:                       # split rows are heavy ones only
                        assert rp[i + 1] - rp[i] > W
                        assert i + j == t * W           # and split exactly on the diagonal
                assert (j - prev[1]) <= 2 * W           # task size bounded by W + heavy
                prev = (i, j)
            assert prev == (M, nnz)


@pytest.mark.parametrize("W", [4, 32, 128])
def test_taskwalk_simulation_matches_oracle(W):
    rng = np.random.default_rng(W)
    M, K, N = 120, 700, 5
    rp, ci, va, dense = random_csr(rng, M, K, 0.01, heavy_rows=(0, 17, 18, 119), empty_rows=(1, 2, 3, 60, 118))
    B = rng.standard_normal((K, N)).astype(np.float32)
    C, written, n_tasks, n_pieces = simulate(rp, ci, va, (M, K), B, W)
    assert np.all(written == 1), "every row of C must be written exactly once"
    ref = oracle.spmm_f64(rp, ci, va, (M, K), B)
    assert oracle.rel_err(C, ref) <= 1e-12
    assert n_pieces > 0


def test_taskwalk_heavy_row_ending_on_a_diagonal():
    # row lengths chosen so that a heavy row's row-end item lands exactly on a task boundary
    W = 8
    lens = [3, 20, 0, 0, 13, 1]           # items: 3+1, 20+1, 1, 1, 13+1, 1+1
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nnz = int(rp[-1])
    rng = np.random.default_rng(1)
    ci = rng.integers(0, 9, nnz).astype(np.int32)
    va = rng.standard_normal(nnz).astype(np.float32)
    B = rng.standard_normal((9, 3)).astype(np.float32)
    for W in range(1, 30):
        C, written, _, _ = simulate(rp, ci, va, (len(lens), 9), B, W)
        assert np.all(written == 1), W
        assert oracle.rel_err(C, oracle.spmm_f64(rp, ci, va, (len(lens), 9), B)) <= 1e-12, W


def test_shard_rows_balances_nnz():
    rng = np.random.default_rng(5)
    lens = np.minimum((rng.pareto(1.2, 20000) * 3).astype(np.int64), 5000)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    for parts in (1, 2, 4, 8):
        b = capi.shard_rows(rp, parts, 1.0, 0.0)
        assert b[0] == 0 and b[-1] == len(lens) and np.all(np.diff(b) >= 0)
        per = np.diff(rp[b].astype(np.int64))
        assert per.sum() == rp[-1]
        assert per.max() <= rp[-1] / parts + lens.max()
    b = capi.shard_rows(rp, 4, 0.0, 1.0)      # rows only
    assert np.all(np.abs(np.diff(b) - len(lens) / 4) <= 1)


def test_new_entry_points_validate_arguments_without_a_device():
    L = capi.lib()
    assert L.b200_plan_forward_host(None, None, 8, None, 0) == capi.ERR_INVALID
    assert b"null plan" in L.b200_last_error()
    assert L.b200_plan_forward_peers(None, None, 8, 8, None, 8, None, 0, None, None) == capi.ERR_INVALID
    assert L.b200_csr_block_diag_cols(None, None, 10, 3, 4, None) == capi.ERR_INVALID      # rows % M != 0
    assert L.b200_csr_block_diag_cols(None, None, 2**30, 1, 4, None) == capi.ERR_INVALID   # L * K overflows int32
    assert L.b200_csr_block_diag_cols(None, None, 0, 3, 4, None) == capi.OK


def test_task_size_rule_is_reflected_in_the_workspace_size():
    L = capi.lib()
    big = L.b200_spmm_workspace_bytes(1 << 20, 16 << 20, 128)      # 256-item tasks
    small = L.b200_spmm_workspace_bytes(1000, 5000, 128)           # 32-item tasks: more tasks per item
    assert big / ((1 << 20) + (16 << 20)) < small / 6000
    capi.set_option("items_per_task", 512)
    assert L.b200_spmm_workspace_bytes(1 << 20, 16 << 20, 128) < big
    capi.set_option("items_per_task", 0)


def test_python_surface_has_the_reference_names():
    """The 16 pybind entry points of the reference's custom_mm (src/custom_mm.cpp:393-416) and the public names
    of its matmuls.py — importable without a GPU (calling them needs one)."""
    import custom_mm
    import matmuls
    for name in ["init_cublas", "destroy_cublas", "init_cusparse", "destroy_cusparse", "cublas_mmul", "cublas_bmm",
                 "cusparse_mmul", "dummy_kernel", "naive_spmm", "tiledspmm_inspect_csr", "tiledspmm_inspect_coo",
                 "tiledspmm_mm", "tiledspmm_clean", "cusparse_inspect", "cusparse_mmul_opt", "cusparse_clean"]:
        assert callable(getattr(custom_mm, name)), name
    for name in ["custom_matmul", "cublasMM", "cublasTransaMM", "cublasTransbMM", "cublasTransabMM",
                 "get_sparse_tensor_properties", "sparse_matmul", "cusparseMM", "naive_matmul", "naiveSpMM"]:
        assert hasattr(matmuls, name), name
    import inspect
    assert list(inspect.signature(matmuls.sparse_matmul).parameters) == ["a", "b", "mm_op"]
    assert list(inspect.signature(matmuls.custom_matmul).parameters) == ["a", "b", "mm_op", "bmm_op", "transa", "transb"]
    assert matmuls.sparse_matmul.__defaults__[0] is custom_mm.cusparse_mmul      # the mm_op plug-in seam
    assert matmuls.naive_matmul.__defaults__[0] is custom_mm.naive_spmm


@pytest.mark.parametrize("W,NW", [(8, 4), (32, 4), (32, 1), (64, 8), (256, 4)])
def test_flat_kernel_decomposition_matches_oracle(W, NW):
    """The flat kernel's bookkeeping (chunk masks, task-edge masks, empty-row fill, head / tail pieces, fix list)
    modelled in numpy on the library's own boundary function: every row written exactly once, product exact."""
    from taskwalk_sim import simulate_flat
    rng = np.random.default_rng(W * 10 + NW)
    M, K, N = 150, 900, 3
    rp, ci, va, _ = random_csr(rng, M, K, 0.012, heavy_rows=(0, 17, 18, 77, 149), empty_rows=(1, 2, 3, 60, 148))
    B = rng.standard_normal((K, N)).astype(np.float32)
    C, written, n_fix = simulate_flat(rp, ci, va, (M, K), B, W, NW)
    assert np.all(written == 1), np.where(written != 1)
    assert oracle.rel_err(C, oracle.spmm_f64(rp, ci, va, (M, K), B)) <= 1e-12
    if W < 900:
        assert n_fix >= 5      # the five full rows are split


def test_flat_kernel_decomposition_edge_shapes():
    from taskwalk_sim import simulate_flat
    rng = np.random.default_rng(3)
    for lens in ([0, 0, 0, 5, 0], [40], [1] * 30, [0] * 9 + [100] + [0] * 9, [7, 8, 9, 16, 24, 0, 1], [3, 20, 0, 0, 13, 1]):
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        nnz = int(rp[-1])
        ci = rng.integers(0, 11, nnz).astype(np.int32)
        va = rng.standard_normal(nnz).astype(np.float32)
        B = rng.standard_normal((11, 2)).astype(np.float32)
        ref = oracle.spmm_f64(rp, ci, va, (len(lens), 11), B)
        for W in (1, 2, 5, 8, 9, 16, 33):
            for NW in (1, 4):
                C, written, _ = simulate_flat(rp, ci, va, (len(lens), 11), B, W, NW)
                assert np.all(written == 1), (lens, W, NW, written)
                assert oracle.rel_err(C, ref) <= 1e-12, (lens, W, NW)
