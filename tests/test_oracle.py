"""The oracle (oracle/) against the fixtures generated from the reference's own Python layer
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle
from conftest import load_golden

FWD = [f"fwd_{i}" for i in range(5)]
CSR = ["csr_special", "csr_ragged", "csr_sparse", "csr_dense", "csr_allzero", "bwd_0"] + FWD[:4]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", CSR)
def test_dense_to_csr_bit_exact(name):
    g = load_golden(name)
    for impl in (oracle.dense_to_csr, oracle.dense_to_csr_numpy):
        rp, ci, va = impl(g["A_dense"])
        assert rp.dtype == np.int32 and ci.dtype == np.int32 and va.dtype == np.float32
        np.testing.assert_array_equal(rp, g["row_ptr"])
        np.testing.assert_array_equal(ci, g["col"])
        np.testing.assert_array_equal(bits(va), bits(g["val"]))  # NaN payloads, denormals, signs


def test_special_values_kept_and_dropped():
    g = load_golden("csr_special")
    rp, ci, va = oracle.dense_to_csr(g["A_dense"])
    assert list(rp) == [0, 2, 6, 6, 8]          # -0.0 dropped, NaN / inf / denormals kept, empty row
    assert np.isnan(va[1]) and va[4] != 0 and va[5] != 0


@pytest.mark.parametrize("name", FWD)
def test_spmm_forward(name):
    g = load_golden(name)
    shape = tuple(g["shape"])
    c64 = oracle.spmm_f64(g["row_ptr"], g["col"], g["val"], shape, g["B"])
    assert oracle.rel_err(g["C"], c64) <= 1e-6   # fixture is fp64 rounded to fp32
    for fused in (True, False):
        c32 = oracle.spmm_f32(g["row_ptr"], g["col"], g["val"], shape, g["B"], fused=fused)
        assert oracle.rel_err(c32, c64) <= 1e-5  # the BASELINE tolerance
        assert np.allclose(c32, g["C"], rtol=1e-5, atol=1e-8)  # the reference tests' own check


@pytest.mark.parametrize("i", range(3))
def test_naive_dense_inputs(i):
    g = load_golden(f"naive_{i}")
    A, B, C = g["A_dense"], g["B"], g["C"]
    if A.ndim == 2:
        A, B, C = A[None], B[None], C[None]
    for a, b, c in zip(A, B, C):
        rp, ci, va = oracle.dense_to_csr(a)
        c64 = oracle.spmm_f64(rp, ci, va, a.shape, b)
        assert oracle.rel_err(c, c64) <= 1e-6


def test_backward_fixture():
    g = load_golden("bwd_0")
    shape = tuple(g["shape"])
    t_rp, t_ci, t_va = oracle.csr_transpose(g["row_ptr"], g["col"], g["val"], shape)
    np.testing.assert_array_equal(t_rp, g["t_row_ptr"])
    np.testing.assert_array_equal(t_ci, g["t_col"])
    np.testing.assert_array_equal(bits(t_va), bits(g["t_val"]))
    gb = oracle.spmm_t_f64(g["row_ptr"], g["col"], g["val"], shape, g["G"])
    assert oracle.rel_err(g["grad_B"], gb) <= 1e-6
    gb2 = oracle.spmm_f64(t_rp, t_ci, t_va, (shape[1], shape[0]), g["G"])
    assert oracle.rel_err(gb2, gb) <= 1e-12
    ga = oracle.sddmm_f64(g["row_ptr"], g["col"], shape, g["G"], g["B"])
    assert oracle.rel_err(g["grad_A_val"], ga) <= 1e-6


def test_oracle_vs_scipy_and_torch():
    import scipy.sparse as sp
    import torch
    rng = np.random.default_rng(3)
    M, K, N = 200, 150, 33
    dense = (rng.standard_normal((M, K)) * (rng.random((M, K)) < 0.1)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    rp, ci, va = oracle.dense_to_csr(dense)
    c64 = oracle.spmm_f64(rp, ci, va, (M, K), B)
    ref = sp.csr_matrix((va.astype(np.float64), ci, rp), shape=(M, K)) @ B.astype(np.float64)
    assert np.abs(c64 - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    t = torch.sparse.mm(torch.from_numpy(dense).double().to_sparse_csr(), torch.from_numpy(B).double())
    assert np.abs(c64 - t.numpy()).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    tc = torch.from_numpy(dense).to_sparse_csr()
    np.testing.assert_array_equal(rp, tc.crow_indices().numpy().astype(np.int32))
    np.testing.assert_array_equal(ci, tc.col_indices().numpy().astype(np.int32))


# ---- fixtures recorded from the REAL reference on a B200 (tests/golden/make_ref_golden.py: oracle/_ref =
# the reference's own custom_mm + matmuls, nothing stubbed) -------------------------------------------------
REF_FWD = [f"ref_fwd_{i}" for i in range(5)]


@pytest.mark.parametrize("name", REF_FWD + ["ref_bwd_0"])
def test_oracle_csr_matches_what_the_reference_built(name):
    g = load_golden(name)
    rp, ci, va = oracle.dense_to_csr(g["A_dense"])
    np.testing.assert_array_equal(rp, g["row_ptr"])      # get_sparse_tensor_properties (matmuls.py:178-187)
    np.testing.assert_array_equal(ci, g["col"])
    np.testing.assert_array_equal(bits(va), bits(g["val"]))


@pytest.mark.parametrize("name", REF_FWD)
def test_oracle_matches_reference_cusparse_output(name):
    g = load_golden(name)
    shape = tuple(int(x) for x in g["shape"])
    c64 = oracle.spmm_f64(g["row_ptr"], g["col"], g["val"], shape, g["B"])
    assert oracle.rel_err(g["C"], c64) <= 1e-5          # cusparseSpMM CSR_ALG2 as the reference calls it
    assert np.allclose(g["C"], c64, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("i", range(3))
def test_oracle_matches_reference_naive_kernel(i):
    g = load_golden(f"ref_naive_{i}")
    A, B, C = g["A_dense"], g["B"], g["C"]
    if A.ndim == 2:
        A, B, C = A[None], B[None], C[None]
    for a, b, c in zip(A, B, C):
        rp, ci, va = oracle.dense_to_csr(a)
        assert oracle.rel_err(c, oracle.spmm_f64(rp, ci, va, a.shape, b)) <= 1e-5
        # the open kernel accumulates in CSR order with fmaf: the fp32 restatement reproduces it bit for bit
        np.testing.assert_array_equal(bits(c), bits(oracle.spmm_f32(rp, ci, va, a.shape, b, fused=True)))


def test_oracle_matches_reference_backward():
    g = load_golden("ref_bwd_0")
    shape = tuple(int(x) for x in g["shape"])
    gb = oracle.spmm_t_f64(g["row_ptr"], g["col"], g["val"], shape, g["G"])
    assert oracle.rel_err(g["grad_B"], gb) <= 1e-5      # matmuls.py:253 through cuSPARSE
    assert oracle.rel_err(g["C"], oracle.spmm_f64(g["row_ptr"], g["col"], g["val"], shape, g["B"])) <= 1e-5
