'''
Row sharding of a CSR matrix across the GPUs of one box (the reference has no multi-GPU code; this is
the one parallelism BASELINE.json's north_star adds: rows of A — and of C — are independent, so A is cut
into contiguous row ranges of equal *cost* and no reduction ever crosses ranks).

  shard_bounds   cost(rows) = nnz * (8 + 4N * miss) + rows * 4N bytes — the non-zero stream, the B rows
                 it gathers from HBM, and the C rows written (pure nnz balance gives row counts that
                 differ by ~100x on the R-MAT matrix)
  shard_csr      the local CSR of one shard (row_ptr rebased to 0)
  RowGather      every rank ends up with all row blocks of C: one symmetric [M,N] buffer per rank,
                 the local SpMM writes its block in place, the other blocks arrive over NCCL
                 (all_gather: one grouped collective; gather_to: only one rank collects)
  backward       grad_A is local to a shard (SDDMM on the shard's pattern with the shard's rows of grad_C);
                 grad_B = sum_g A_g^T grad_C_g: every rank forms its [K, N] partial product over its cached
                 transposed shard and the partials are reduce-scattered (rank g ends with rows
                 [g K / G, (g + 1) K / G) of grad_B), matmuls.py:250,253 being the formulas.  Sharding A^T by rows
                 and all-gathering grad_C instead moves the same (G - 1) / G * 4 K N bytes per rank for square
                 operands and needs a second, differently balanced sharding of A: not built.
'''
from __future__ import annotations

import numpy as np

from . import _capi as capi


def shard_bounds(row_ptr, parts: int, N: int, miss: float = 0.29) -> np.ndarray:
    '''Row boundaries (parts + 1 entries) of cost-balanced contiguous shards.

    ``miss`` sets the relative price of a gathered B row against a written C row; 0.29 reproduces the
    measured per-shard times on the R-MAT config (benchmarks/shard_balance.py: one row costs about 3.3 non-zeros).'''
    if parts <= 1:
        return np.array([0, len(row_ptr) - 1], dtype=np.int64)
    return capi.shard_rows(row_ptr, parts, 8.0 + 4.0 * N * miss, 4.0 * N)


def rebalance_bounds(row_ptr, bounds, times, N: int, miss: float = 0.29, fixed_ms: float = 0.02) -> np.ndarray:
    '''New row boundaries from MEASURED per-shard times (inspector step of a multi-GPU plan: every rank times its own
    shard during warm-up, the times are all-gathered, the rows are re-cut).  A shard's time is a fixed part (launch,
    first-wave staging, last-wave drain: `fixed_ms`, 0.02 ms measured on the gather kernels, profiles/r2_scaling.md)
    plus a part that follows its rows; the model cost of a row, nnz * (8 + 4N * miss) + 4N, is scaled inside every
    current shard by (measured time - fixed part) / (model cost of the shard), and the boundaries move to the equal
    quantiles of that corrected cumulative cost.  (Scaling by the whole time, as the first version did, under-corrects:
    moving rows only moves the variable part.)  The caller probes a few rounds and keeps the best boundaries it saw
    (bench.py): max/mean of the shard times ~1.01-1.02 on the R-MAT matrix (static model: ~1.05).'''
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    bounds = np.asarray(bounds, dtype=np.int64)
    parts = len(bounds) - 1
    M = len(row_ptr) - 1
    cost = np.diff(row_ptr).astype(np.float64) * (8.0 + 4.0 * N * miss) + 4.0 * N
    tmin = min(float(t) for t in times) if len(times) else 0.0
    fixed = min(float(fixed_ms), 0.5 * tmin) if tmin > 0 else 0.0  # never more than half of the fastest shard
    for g in range(parts):
        r0, r1 = int(bounds[g]), int(bounds[g + 1])
        c = cost[r0:r1].sum()
        if c > 0 and times[g] > 0:
            cost[r0:r1] *= (float(times[g]) - fixed) / c
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    targets = cum[-1] * np.arange(1, parts) / parts
    new = np.concatenate([[0], np.searchsorted(cum, targets, side='left'), [M]]).astype(np.int64)
    return np.maximum.accumulate(np.clip(new, 0, M))


def shard_csr(row_ptr, col, val, bounds, g: int):
    r0, r1 = int(bounds[g]), int(bounds[g + 1])
    p0, p1 = int(row_ptr[r0]), int(row_ptr[r1])
    return (np.asarray(row_ptr[r0:r1 + 1]) - p0).astype(np.int32), col[p0:p1], val[p0:p1], (r0, r1)


class RowGather:
    '''All-gather of unequal row blocks of a row-major [M, N] matrix.

    ``buffer`` is the full matrix on this rank; rank g owns rows [bounds[g], bounds[g+1]) and writes them
    itself; :meth:`all_gather` fills in the rest (one broadcast per owner, issued back to back so NCCL
    pipelines them over NVLink).  Works with any torch.distributed backend (gloo on CPU in the tests).'''

    def __init__(self, bounds, N: int, device, dist, dtype=None):
        import torch
        self.bounds = [int(b) for b in bounds]
        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.buffer = torch.empty(self.bounds[-1], N, device=device, dtype=dtype or torch.float32)
        self.blocks = [self.buffer[self.bounds[g]:self.bounds[g + 1]] for g in range(self.world)]

    @property
    def local(self):
        return self.blocks[self.rank]

    def all_gather(self):
        '''Every rank receives every other rank's row block.  NCCL: ONE grouped collective (all_gather with unequal
        block sizes is issued by torch as a single ncclGroup of broadcasts); other backends: one broadcast per owner.'''
        if self.dist.get_backend() == 'nccl' and all(b.numel() > 0 for b in self.blocks):
            self.dist.all_gather(list(self.blocks), self.local)
            return self.buffer
        works = []
        for g in range(self.world):
            if self.blocks[g].numel() == 0:
                continue
            works.append(self.dist.broadcast(self.blocks[g], src=g, async_op=True))
        for wk in works:
            wk.wait()
        return self.buffer

    def gather_to(self, dst: int = 0):
        '''Only rank ``dst`` collects all row blocks (north_star: "C gathered via NCCL"): one batch of point-to-point
        transfers; dst's NVLink ingress is the bound ((world - 1) / world of C).'''
        ops = []
        if self.rank == dst:
            for g in range(self.world):
                if g != dst and self.blocks[g].numel() > 0:
                    ops.append(self.dist.P2POp(self.dist.irecv, self.blocks[g], g))
        elif self.local.numel() > 0:
            ops.append(self.dist.P2POp(self.dist.isend, self.local, dst))
        if ops:
            for wk in self.dist.batch_isend_irecv(ops):
                wk.wait()
        return self.buffer


def reduce_scatter_rows(partial, out, dist):
    '''out[K / world, N] = rows [rank K / world, (rank + 1) K / world) of the sum over ranks of partial[K, N]
    (the grad_B exchange of the row-sharded backward).  K must divide by the world size.'''
    world, rank = dist.get_world_size(), dist.get_rank()
    K = partial.shape[0]
    assert K % world == 0 and out.shape[0] == K // world
    if dist.get_backend() == 'nccl':
        dist.reduce_scatter_tensor(out, partial)
    else:  # gloo (CPU tests): all-reduce, keep the own slice
        tmp = partial.clone()
        dist.all_reduce(tmp)
        out.copy_(tmp[rank * (K // world):(rank + 1) * (K // world)])
    return out


class SymmetricRowGather:
    '''The all-gather of C fused into the SpMM epilogue: C lives in a symmetric-memory buffer (same
    allocation on every rank, every rank's copy mapped into every process over NVLink), and the forward
    kernel stores each finished row into all copies at once — P2P stores to the mapped peer pointers, or a
    single ``multimem.st`` to the NVLS multicast address when the fabric supports it.  After the kernel a
    device-side barrier on the signal pads makes the rows visible everywhere; no NCCL collective follows.

    torch.distributed._symmetric_memory is used for the plumbing only (allocation, handle exchange, barrier).'''

    def __init__(self, bounds, N: int, device, dist, use_multicast: bool = True):
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        self.bounds = [int(b) for b in bounds]
        self.dist = dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        M = self.bounds[-1]
        self.buffer = symm_mem.empty(M, N, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buffer, dist.group.WORLD)
        r0 = self.bounds[self.rank]
        off = r0 * N * 4
        self.local = self.buffer[r0:self.bounds[self.rank + 1]]
        ptrs = list(self.handle.buffer_ptrs)
        self.peer_ptrs = [int(ptrs[q]) + off for q in range(self.world) if q != self.rank]
        mc = int(self.handle.multicast_ptr or 0) if use_multicast else 0  # 0: no NVLS multicast on this fabric
        self.mc_ptr = (mc + off) if mc else 0

    def forward(self, plan, B, stream=0):
        '''Rank-local SpMM whose epilogue writes this rank's rows of C into every rank's buffer.'''
        if self.mc_ptr:
            plan.forward_peers(B, self.local, (), self.mc_ptr, stream=stream)
        else:
            plan.forward_peers(B, self.local, self.peer_ptrs, 0, stream=stream)

    def barrier(self):
        self.handle.barrier(channel=0)
