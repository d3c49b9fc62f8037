"""Small run of every kernel family for compute-sanitizer (memcheck): `compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py`."""
import sys; sys.path.insert(0, '.')
import numpy as np
from matrix_multiplication_b200 import capi
sys.path.insert(0, 'tests')
from conftest import random_csr
from oracle import oracle

rng = np.random.default_rng(0)
for (M, K, N, dens, heavy, empty) in [(300, 900, 128, 0.02, (3, 250), (0, 1, 299)), (257, 513, 40, 0.05, (128,), (256,)),
                                      (100, 3000, 8, 0.01, (0, 99), (1,)), (64, 64, 300, 0.2, (), ()), (50, 70, 1, 0.3, (), ()),
                                      (40, 5000, 256, 0.2, (0, 1, 2), ())]:
    rp, ci, va, dense = random_csr(rng, M, K, dens, heavy_rows=heavy, empty_rows=empty)
    B = rng.standard_normal((K, N)).astype(np.float32)
    G = rng.standard_normal((M, N)).astype(np.float32)
    for stage in (1, 0):
        capi.set_option("stage_mode", stage)
        C = capi.spmm_host(rp, ci, va, (M, K), B)
        assert oracle.rel_err(C, oracle.spmm_f64(rp, ci, va, (M, K), B)) <= 1e-5
    capi.set_option("stage_mode", 1)
    assert oracle.rel_err(capi.backward_B_host(rp, ci, va, (M, K), G), oracle.spmm_t_f64(rp, ci, va, (M, K), G)) <= 1e-5
    assert oracle.rel_err(capi.sddmm_host(rp, ci, (M, K), G, B), oracle.sddmm_f64(rp, ci, (M, K), G, B)) <= 1e-5
    r2, c2, v2 = capi.dense_to_csr_host(dense)
    assert np.array_equal(r2, rp) and np.array_equal(c2, ci)
print("sanitizer smoke ok")
