"""dense -> CSR of a 16384 x 16384 matrix at 90 % sparsity, a few times (driver for ncu; prints CUDA-event timings)."""
import sys; sys.path.insert(0, '.')
import torch, custom_mm
torch.manual_seed(0)
d = torch.rand(16384, 16384, device='cuda') * (torch.rand(16384, 16384, device='cuda') >= 0.9)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(2):
    custom_mm.dense_to_csr(d)
ts = []
for _ in range(5):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); crow, col, val = custom_mm.dense_to_csr(d); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
nnz = val.numel()
by = 4.0 * d.numel() + 4.0 * (d.shape[0] + 1) + 8.0 * nnz  # the dense matrix counted once
t = sorted(ts)[len(ts) // 2]
print(f"dense_to_csr 16384^2 @90%: nnz {nnz}, median {t:.4f} ms, {by / t / 1e6:.0f} GB/s on algorithmic bytes")
t = d.to_sparse_csr()
assert torch.equal(crow.long(), t.crow_indices()) and torch.equal(col.long(), t.col_indices()) and torch.equal(val, t.values())
print('bit-identical to torch.to_sparse_csr')
