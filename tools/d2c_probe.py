"""dense -> CSR timing: one-pass vs two-pass (count+fill) vs torch, 16384^2 at several sparsities."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matrix_multiplication_b200 import capi
import custom_mm
dev = torch.device("cuda")
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
M = K = 16384
for sp in (0.5, 0.9, 0.99, 0.999):
    g = torch.Generator(device=dev).manual_seed(0)
    d = torch.rand(M, K, device=dev, generator=g) * (torch.rand(M, K, device=dev, generator=g) >= sp)
    crow = torch.empty(M + 1, dtype=torch.int32, device=dev)
    nnz = int((d != 0).sum())
    col = torch.empty(nnz, dtype=torch.int32, device=dev); val = torch.empty(nnz, device=dev)
    sb = capi.lib().b200_dense_to_csr_onepass_scratch_bytes(M, K)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    import ctypes as C
    def onepass():
        capi.check(capi.lib().b200_dense_to_csr_onepass(d.data_ptr(), M, K, K, crow.data_ptr(), col.data_ptr(), val.data_ptr(), nnz, scratch.data_ptr(), sb, None, None))
    def twopass():
        capi.check(capi.lib().b200_dense_to_csr_count(d.data_ptr(), M, K, K, crow.data_ptr(), None, None))
        capi.check(capi.lib().b200_dense_to_csr_fill(d.data_ptr(), M, K, K, crow.data_ptr(), col.data_ptr(), val.data_ptr(), None))
    for wv in (16, 20, 24):
        for small in (0, 1):
            capi.set_option("d2c_warps", wv); capi.set_option("d2c_lag", small)
            print(f"  count warps {wv} pair ring {256 if small else 512}: {timeit(onepass):.4f} ms", flush=True)
    for wv in (16, 20, 24):
        capi.set_option("d2c_warps", wv); capi.set_option("d2c_lag", 256)
        print(f"  count warps {wv} count side alone (no fill, no gating): {timeit(onepass):.4f} ms", flush=True)
    capi.set_option("d2c_warps", 0); capi.set_option("d2c_lag", 0)
    twopass(); torch.cuda.synchronize()
    c2, v2, r2 = col.clone(), val.clone(), crow.clone()
    col.zero_(); val.zero_(); crow.zero_()
    onepass(); torch.cuda.synchronize()
    print("  one-pass == two-pass:", bool(torch.equal(c2, col) and torch.equal(v2.view(torch.int32), val.view(torch.int32)) and torch.equal(r2, crow)), flush=True)
    t1, t2 = timeit(onepass), timeit(twopass)
    custom_mm.set_onepass_min_elems(1 << 20)
    t3 = timeit(lambda: custom_mm.dense_to_csr(d))
    by = 4.0 * M * K + 4 * (M + 1) + 8.0 * nnz
    print(f"sparsity {sp}: nnz {nnz}  one-pass {t1:.4f} ms ({by/t1/1e6:.0f} GB/s, {by/t1/1e6/6549.1:.3f} of roof)  two-pass {t2:.4f} ms ({by/t2/1e6/6549.1:.3f})  custom_mm.dense_to_csr (adaptive; alloc + sync) {t3:.4f} ms", flush=True)
