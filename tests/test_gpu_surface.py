"""The reference-shaped surface on the GPU: `custom_mm` entry points and the `matmuls` autograd
wrappers, written the way the reference's own test scripts use them
(tests/cusparse_kernel_test.py, tests/naive_kernel_test.py, tests/cublas_kernel_test.py,
tests/tiledsppm_kernel_test.py) but seeded and asserting."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    import custom_mm
    import matmuls
    custom_mm.init_cublas()
    custom_mm.init_cusparse()
    yield torch, custom_mm, matmuls
    custom_mm.destroy_cusparse()
    custom_mm.destroy_cublas()


def generate_sparse_dense(torch, shape, n_vals, seed):
    """tests/cusparse_kernel_test.py:13-30: n_vals random coordinates != (0,0), values U[0,1)."""
    rnd = random.Random(seed)
    a = torch.zeros(shape)
    seen = {(0, 0)}
    for _ in range(n_vals):
        x, y = 0, 0
        while (x, y) in seen:
            x, y = rnd.randrange(shape[0]), rnd.randrange(shape[1])
        seen.add((x, y))
        a[x, y] = rnd.random()
    return a


REF_SHAPES = [((10, 10), (10, 10)), ((10, 20), (20, 10)), ((10, 10), (10, 5)), ((20, 10), (10, 5)),
              ((512, 1024), (1024, 256))]  # tests/cusparse_kernel_test.py:32-38


@pytest.mark.parametrize("i", range(len(REF_SHAPES)))
@pytest.mark.parametrize("op", ["cusparseMM", "naiveSpMM"])
def test_reference_forward_shapes(mods, i, op):
    torch, custom_mm, matmuls = mods
    a_shape, b_shape = REF_SHAPES[i]
    torch.manual_seed(i)
    a = generate_sparse_dense(torch, a_shape, a_shape[0] * a_shape[1] // 10, i).cuda()
    b = torch.rand(b_shape).cuda()
    n0 = custom_mm.kernel_launch_count()
    ours = getattr(matmuls, op).apply(a.to_sparse_csr(), b)
    assert custom_mm.kernel_launch_count() > n0, "the CUDA engine did not run"
    expected = (a.double() @ b.double()).float()
    assert torch.allclose(expected, ours, rtol=1e-5, atol=1e-7)
    ours_dense_in = getattr(matmuls, op).apply(a, b)  # dense `a`: converted on the device, plan-less gather kernels
    # (a CSR tensor is inspected and may take the block-dense tensor-core tiles: same tolerance, other rounding)
    assert torch.allclose(expected, ours_dense_in, rtol=1e-5, atol=1e-7)


def test_dense_to_csr_matches_torch_bit_for_bit(mods):
    torch, custom_mm, _ = mods
    torch.manual_seed(0)
    d = torch.randn(777, 1234, device="cuda") * (torch.rand(777, 1234, device="cuda") < 0.2)
    d[5, 7] = float("nan")
    d[9, 9] = -0.0
    crow, col, val = custom_mm.dense_to_csr(d)
    t = d.to_sparse_csr()
    assert crow.dtype == torch.int32 and col.dtype == torch.int32
    assert torch.equal(crow.long(), t.crow_indices())
    assert torch.equal(col.long(), t.col_indices())
    assert torch.equal(val.view(torch.int32), t.values().view(torch.int32))


@pytest.mark.parametrize("shape", [(64, 48, 16), (300, 500, 65), (1024, 1024, 64)])
def test_cusparseMM_backward(mods, shape):
    torch, _, matmuls = mods
    M, K, N = shape
    torch.manual_seed(M)
    dense = torch.randn(M, K, device="cuda") * (torch.rand(M, K, device="cuda") < 0.1)
    a = dense.to_sparse_csr().requires_grad_(True)
    b = torch.randn(K, N, device="cuda", requires_grad=True)
    g = torch.randn(M, N, device="cuda")
    out = matmuls.cusparseMM.apply(a, b)
    out.backward(g)
    ref_gb = (dense.double().t() @ g.double())
    assert float((b.grad.double() - ref_gb).abs().max() / ref_gb.abs().max()) <= 1e-5
    assert a.grad.layout == torch.sparse_csr
    full = g.double() @ b.detach().double().t()
    rows = torch.repeat_interleave(torch.arange(M, device="cuda"), a.detach().crow_indices().diff())
    ref_ga = full[rows, a.detach().col_indices()]
    assert float((a.grad.values().double() - ref_ga).abs().max() / ref_ga.abs().max()) <= 1e-5
    # dense `a`: the full dense gradient the reference intended (matmuls.py:250)
    a2 = dense.clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True)
    matmuls.cusparseMM.apply(a2, b2).backward(g)
    assert float((a2.grad.double() - full).abs().max() / full.abs().max()) <= 1e-5
    assert float((b2.grad.double() - ref_gb).abs().max() / ref_gb.abs().max()) <= 1e-5


def test_transposed_csr_is_cached_per_tensor(mods):
    torch, custom_mm, matmuls = mods
    dense = torch.randn(200, 100, device="cuda") * (torch.rand(200, 100, device="cuda") < 0.1)
    a = dense.to_sparse_csr()
    m = matmuls.as_csr(a)
    assert matmuls.as_csr(a) is m
    t1 = m.transposed()
    assert m.transposed() is t1


@pytest.mark.parametrize("shapes", [((4, 2), (2, 3)), ((2, 4, 2), (2, 2, 3)), ((3, 2, 8, 16), (3, 2, 16, 4)),
                                    ((6, 5), (3, 5, 4)), ((3, 6, 5), (5, 4))])
def test_batched_dispatch(mods, shapes):
    """tests/naive_kernel_test.py:62-64 shapes (dense inputs) plus the 2-D x 3-D / 3-D x 2-D cases the
    reference gets wrong."""
    torch, _, matmuls = mods
    torch.manual_seed(1)
    a = torch.rand(shapes[0], device="cuda")
    b = torch.rand(shapes[1], device="cuda")
    expected = (a.double() @ b.double()).float()
    for fn in (matmuls.naive_matmul, matmuls.sparse_matmul):
        out = fn(a, b)
        assert out.shape == expected.shape
        assert torch.allclose(expected, out, rtol=1e-5, atol=1e-6)
    out = matmuls.naiveSpMM.apply(a, b)
    assert torch.allclose(expected, out, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("cls,ta,tb", [("cublasMM", False, False), ("cublasTransaMM", True, False),
                                       ("cublasTransbMM", False, True), ("cublasTransabMM", True, True)])
@pytest.mark.parametrize("batch", [(), (3,), (2, 3)])
def test_cublas_family_forward_backward(mods, cls, ta, tb, batch):
    torch, _, matmuls = mods
    torch.manual_seed(2)
    M, K, N = 24, 40, 16
    a = torch.randn(*batch, *((K, M) if ta else (M, K)), device="cuda", requires_grad=True)
    b = torch.randn(*batch, *((N, K) if tb else (K, N)), device="cuda", requires_grad=True)
    out = getattr(matmuls, cls).apply(a, b)
    ad, bd = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    ref = torch.matmul(ad.transpose(-1, -2) if ta else ad, bd.transpose(-1, -2) if tb else bd)
    g = torch.randn_like(ref)
    out.backward(g)
    ref.backward(g)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4)
    assert torch.allclose(a.grad, ad.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(b.grad, bd.grad, rtol=1e-4, atol=1e-4)


def test_inspect_execute_clean_registries(mods):
    torch, custom_mm, _ = mods
    torch.manual_seed(3)
    M, Kin, Nout = 96, 160, 24  # A is M x Kin, dense operand Kin x Nout
    dense = torch.randn(M, Kin, device="cuda") * (torch.rand(M, Kin, device="cuda") < 0.2)
    csr = dense.to_sparse_csr()
    crow, col, val = csr.crow_indices().int(), csr.col_indices().int(), csr.values()
    Brm = torch.randn(Kin, Nout, device="cuda")
    expected = (dense.double() @ Brm.double()).float()
    # reference argument order: (displ, colindex, value, nnz, M, N, K, layer) with N = dense columns
    custom_mm.cusparse_inspect(crow, col, val, val.numel(), M, Nout, Kin, "fc1")
    # column-major B / C, as the reference compiles them (src/baseline_mm.cu:297,300)
    Bcm = Brm.t().contiguous()
    Ccm = torch.empty(Nout, M, device="cuda")
    custom_mm.cusparse_mmul_opt(Bcm, Ccm, "fc1")
    assert torch.allclose(Ccm.t(), expected, rtol=1e-5, atol=1e-5)
    Crm = torch.empty(M, Nout, device="cuda")
    custom_mm.cusparse_mmul_opt(Brm, Crm, "fc1", True)
    assert torch.allclose(Crm, expected, rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        custom_mm.cusparse_mmul_opt(Brm, Crm, "no-such-layer", True)
    custom_mm.cusparse_clean()
    with pytest.raises(RuntimeError):
        custom_mm.cusparse_mmul_opt(Brm, Crm, "fc1", True)

    # TiledSpMM surface (src/custom_mm.cpp:283-290: C = M x K, A = M x N, B = N x K; CPU inputs)
    cpu = csr.cpu()
    custom_mm.tiledspmm_inspect_csr(M, Kin, Nout, cpu.crow_indices(), cpu.col_indices(), cpu.values(), "t1")
    Ccm.zero_()
    custom_mm.tiledspmm_mm(Bcm, Ccm, "t1")
    assert torch.allclose(Ccm.t(), expected, rtol=1e-5, atol=1e-5)
    coo = dense.cpu().to_sparse_coo().coalesce()
    perm = torch.randperm(coo.values().numel())  # COO in arbitrary order
    custom_mm.tiledspmm_inspect_coo(M, Kin, Nout, coo.values().numel(), coo.indices()[0][perm].int(),
                                    coo.indices()[1][perm].int(), coo.values()[perm], "t2")
    Crm.zero_()
    custom_mm.tiledspmm_mm(Brm, Crm, "t2", True)
    assert torch.allclose(Crm, expected, rtol=1e-5, atol=1e-5)
    custom_mm.tiledspmm_clean()


def test_errors_are_raised_not_swallowed(mods):
    torch, custom_mm, matmuls = mods
    a = torch.rand(4, 4, device="cuda").to_sparse_csr()
    crow, col, val = a.crow_indices().int(), a.col_indices().int(), a.values()
    b = torch.rand(4, 4, device="cuda")
    with pytest.raises(RuntimeError):  # wrong dtype (reference: data_ptr<float>() throws)
        custom_mm.cusparse_mmul(val.double(), col, crow, 16, 4, 4, b, torch.empty(4, 4, device="cuda"))
    with pytest.raises(RuntimeError):  # strided operand: rejected, not reinterpreted
        custom_mm.cusparse_mmul(val, col, crow, 16, 4, 4, torch.rand(4, 8, device="cuda").t()[:4, :4],
                                torch.empty(4, 4, device="cuda"))
    with pytest.raises(RuntimeError):
        matmuls.sparse_matmul(torch.rand(4, 4), torch.rand(4, 4))  # CPU tensors: no CPU fallback
    with pytest.raises(ValueError):
        custom_mm.cublas_bmm(b, b, b, 7, False, False)
    custom_mm.dummy_kernel()


def test_fc_layer_like_the_reference_benchmark(mods):
    """benchmarks/cusparse_fc_layer.py:41 — `cusparseMM.apply(inp, self.weight.t())` — with the sparse
    operand first (SURVEY.md 'decisions'): y^T = W_csr x^T, MNIST-shaped MLP 784-512-10, batch 256."""
    torch, _, matmuls = mods
    from matrix_multiplication_b200.layers import cusparseLinear
    torch.manual_seed(4)
    x = torch.rand(256, 784, device="cuda")
    fc1 = cusparseLinear(784, 512, sparsity=0.9, seed=1).cuda()
    fc2 = cusparseLinear(512, 10, sparsity=0.9, seed=2).cuda()
    y = fc2(torch.relu(fc1(x)))
    y.sum().backward()
    w1, w2 = fc1.dense_weight().double(), fc2.dense_weight().double()
    xd = x.double()
    h = torch.relu(xd @ w1.t() + fc1.bias.double())
    ref = h @ w2.t() + fc2.bias.double()
    assert float((y.double() - ref).abs().max() / ref.abs().max()) <= 1e-5
    # grad of the CSR values = (dL/dy)^T x sampled on the pattern
    gy = torch.ones_like(ref)
    gh = (gy @ w2) * (h > 0)
    gw1 = gh.t() @ xd
    got = fc1.dense_weight_grad().double()
    mask = fc1.dense_weight() != 0
    assert float((got - gw1 * mask).abs().max() / gw1.abs().max()) <= 1e-5
    assert float((fc1.bias.grad.double() - gh.sum(0)).abs().max() / gh.sum(0).abs().max()) <= 1e-5


@pytest.mark.parametrize("tc_mode", [0, 2])
def test_fc_layer_fused_relu_both_paths(mods, tc_mode):
    """cusparseLinear(activation='relu') on the gather kernels (tc_mode 0: explicit transposes, bias + ReLU fused into
    the kernels' store) and on the block-dense tensor-core tiles (tc_mode 2: activations read and written as
    [tokens, features], no transposed copies): forward, dx, dvalues, dbias against fp64."""
    torch, custom_mm, _ = mods
    from matrix_multiplication_b200.layers import cusparseLinear
    custom_mm.set_option("tc_mode", tc_mode)
    try:
        torch.manual_seed(6)
        x = torch.randn(200, 300, device="cuda", requires_grad=True)
        fc = cusparseLinear(300, 260, sparsity=0.85, seed=3, activation="relu").cuda()
        with torch.no_grad():
            fc.bias.copy_(torch.randn(260, device="cuda") * 0.1)
        y = fc(x)
        assert fc.csr().plan().last_path() == (1 if tc_mode == 2 else 0)
        w = fc.dense_weight().double()
        pre = x.detach().double() @ w.t() + fc.bias.detach().double()
        ref = pre.clamp_min(0)
        assert float((y.double() - ref).abs().max() / ref.abs().max()) <= 1e-5
        gy = torch.randn_like(y)
        y.backward(gy)
        g = gy.double() * (pre > 0)
        assert float((x.grad.double() - g @ w).abs().max() / (g @ w).abs().max()) <= 1e-5
        gw = g.t() @ x.detach().double()
        mask = fc.dense_weight() != 0
        assert float((fc.dense_weight_grad().double() - gw * mask).abs().max() / gw.abs().max()) <= 1e-5
        assert float((fc.bias.grad.double() - g.sum(0)).abs().max() / g.sum(0).abs().max()) <= 1e-5
        # in-place update of the values (an optimizer step): the next forward sees them, on both plan sides
        with torch.no_grad():
            fc.values.mul_(0.5)
        y2 = fc(x.detach())
        ref2 = (x.detach().double() @ (w * 0.5).t() + fc.bias.detach().double()).clamp_min(0)
        assert float((y2.double() - ref2).abs().max() / ref2.abs().max()) <= 1e-5
    finally:
        custom_mm.set_option("tc_mode", 1)


def test_fc_layer_loads_the_reference_layers_dense_weight(mods):
    torch, _, _ = mods
    from matrix_multiplication_b200.layers import cusparseLinear
    fc = cusparseLinear(64, 32, sparsity=0.8, seed=1).cuda()
    w = fc.dense_weight() * 3.0
    fc2 = cusparseLinear(64, 32, sparsity=0.8, seed=1).cuda()
    fc2.load_state_dict({"weight": w, "bias": torch.ones(32)})   # keys of benchmarks/cusparse_fc_layer.py:17-21
    assert torch.equal(fc2.dense_weight(), w) and bool((fc2.bias == 1).all())
    with pytest.raises(RuntimeError):
        fc2.load_state_dict({"weight": torch.ones(32, 64), "bias": torch.ones(32)})  # a different pattern


def test_cusparsemm_on_a_csr_tensor_uses_the_inspected_plan(mods):
    """cusparseMM.apply on a CSR tensor: the operand is inspected once (plan cached on the tensor) and a dense-enough
    one runs on the tensor-core tiles; forward and both gradients against fp64."""
    torch, custom_mm, matmuls = mods
    torch.manual_seed(7)
    a = (torch.rand(1536, 1024, device="cuda") * (torch.rand(1536, 1024, device="cuda") < 0.2)).to_sparse_csr()
    b = torch.rand(1024, 1024, device="cuda", requires_grad=True)
    a.requires_grad_(True)
    out = matmuls.cusparseMM.apply(a, b)
    assert matmuls.as_csr(a).plan().last_path() == 1  # large and dense enough: the cost model picks the tiles
    ad = a.detach().to_dense().double()
    ref = ad @ b.detach().double()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) <= 1e-5
    g = torch.rand_like(out)
    out.backward(g)
    gb = ad.t() @ g.double()
    assert float((b.grad.double() - gb).abs().max() / gb.abs().max()) <= 1e-5
    ga = (g.double() @ b.detach().double().t()) * (ad != 0)
    assert float((a.grad.to_dense().double() - ga).abs().max() / ga.abs().max()) <= 1e-5


def test_batched_product_is_one_launch(mods):
    """N-D x N-D: the reference loops over the batch in Python (matmuls.py:221-225; its own test drives 4096
    sequential calls, tests/naive_kernel_test.py:67).  Here the batch is one block-diagonal SpMM."""
    torch, custom_mm, matmuls = mods
    torch.manual_seed(5)
    a = torch.rand(16, 4, 64, 64, device="cuda") * (torch.rand(16, 4, 64, 64, device="cuda") < 0.2)
    b = torch.rand(16, 4, 64, 32, device="cuda")
    matmuls.naiveSpMM.apply(a, b)  # warm-up (scratch buffers)
    n0 = custom_mm.kernel_launch_count()
    out = matmuls.naiveSpMM.apply(a, b)
    launches = custom_mm.kernel_launch_count() - n0
    assert launches <= 8, f"{launches} kernel launches for one batched product"
    expected = (a.double() @ b.double()).float()
    assert out.shape == expected.shape
    assert torch.allclose(expected, out, rtol=1e-5, atol=1e-6)
    # gradients through the batched wrapper
    a2 = a.clone().requires_grad_(True)
    b2 = b.clone().requires_grad_(True)
    matmuls.cusparseMM.apply(a2, b2).sum().backward()
    ones = torch.ones_like(expected).double()
    assert torch.allclose(b2.grad.double(), a.double().transpose(-1, -2) @ ones, rtol=1e-4, atol=1e-4)
    assert torch.allclose(a2.grad.double(), ones @ b.double().transpose(-1, -2), rtol=1e-4, atol=1e-4)


def test_dense_to_csr_adaptive_one_pass(mods):
    """custom_mm.dense_to_csr on large matrices: the one-pass kernel for sparse inputs, its capacity overflow (denser
    than one kept element in eight) finished by the exact fill, the remembered density sending the next dense
    conversion straight to the two-phase pair — every route bit-identical to torch.to_sparse_csr."""
    torch, custom_mm, _ = mods
    torch.manual_seed(11)

    def check(d, compact=False):
        crow, col, val = custom_mm.dense_to_csr(d, compact)
        t = d.contiguous().to_sparse_csr()
        assert torch.equal(crow.long(), t.crow_indices())
        assert torch.equal(col.long(), t.col_indices())
        assert torch.equal(val.view(torch.int32), t.values().view(torch.int32))

    custom_mm.set_onepass_min_elems(1 << 20)                        # (also forgets the remembered density)
    base = torch.randn(1500, 2052, device="cuda")
    keep = torch.rand(1500, 2052, device="cuda")
    for dens in (0.001, 0.02, 0.1):                                  # one pass: compacted chunks
        check((base * (keep < dens))[:, :2048])                      # ld = 2052: strided rows, 16-byte aligned
    sparse = base * (keep < 0.02)
    sparse[3] = 0
    sparse[4] = 2.5                                                  # a full row inside a sparse matrix: re-read chunks
    sparse[5, 7] = float("nan")
    sparse[6, 9] = -0.0
    check(sparse, compact=True)
    check(base * (keep < 0.5))                                       # overflow: one pass, then the exact fill
    check(base * (keep < 0.5))                                       # remembered: two-phase directly
    check(base * (keep < 0.01))                                      # (still two-phase: one sparse result re-arms)
    check(base * (keep < 0.01))                                      # one pass again
    custom_mm.set_onepass_min_elems(-1)
    check(base * (keep < 0.01))                                      # never
    custom_mm.set_onepass_min_elems(1 << 20)


@pytest.mark.parametrize("K,ld", [(130, 132), (129, 132), (131, 256), (4, 8), (5, 8), (1000, 1024), (4096, 4096)])
def test_dense_to_csr_strided_and_ragged_widths(mods, K, ld):
    """The 128-bit fill path (ld % 4 == 0) with K not a multiple of 4 / of 128, on a strided view."""
    torch, custom_mm, _ = mods
    torch.manual_seed(K)
    big = torch.randn(300, ld, device="cuda") * (torch.rand(300, ld, device="cuda") < 0.3)
    big[7] = 0                      # an empty row
    big[9] = 1.5                    # a full row
    d = big[:, :K]
    crow, col, val = custom_mm.dense_to_csr(d)
    t = d.contiguous().to_sparse_csr()
    assert torch.equal(crow.long(), t.crow_indices())
    assert torch.equal(col.long(), t.col_indices())
    assert torch.equal(val.view(torch.int32), t.values().view(torch.int32))
