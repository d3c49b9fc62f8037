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
'''
from __future__ import annotations

import numpy as np

from . import _capi as capi


def shard_bounds(row_ptr, parts: int, N: int, miss: float = 0.5) -> np.ndarray:
    '''Row boundaries (parts + 1 entries) of cost-balanced contiguous shards.'''
    if parts <= 1:
        return np.array([0, len(row_ptr) - 1], dtype=np.int64)
    return capi.shard_rows(row_ptr, parts, 8.0 + 4.0 * N * miss, 4.0 * N)


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
        works = []
        for g in range(self.world):
            if self.blocks[g].numel() == 0:
                continue
            works.append(self.dist.broadcast(self.blocks[g], src=g, async_op=True))
        for wk in works:
            wk.wait()
        return self.buffer
