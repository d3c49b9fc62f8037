"""Seeded synthetic inputs for the BASELINE.json configs.  Host-side numpy only.

  bernoulli_csr   element-wise Bernoulli(1 - sparsity) mask, values U[0,1) or N(0,1)
                  (the reference's generators: tests/cusparse_kernel_test.py:13-30,46-50 use
                  random coordinates with U[0,1) values; benchmarks/random_tensor_benchmark.py:29-52
                  sparsifies a random tensor)
  rmat_csr        R-MAT power-law graph (a,b,c,d) = (.57,.19,.19,.05), duplicates merged, rows sorted
"""
from __future__ import annotations

import numpy as np


def bernoulli_csr(M: int, K: int, sparsity: float, seed: int = 0, signed: bool = False):
    """CSR (row_ptr int32, col int32, val fp32) of an M x K matrix with P(non-zero) = 1 - sparsity.

    Generated row block by row block so that 16k x 16k fits comfortably in host memory."""
    rng = np.random.default_rng(seed)
    density = 1.0 - sparsity
    counts = np.zeros(M, dtype=np.int64)
    cols, vals = [], []
    step = max(1, min(M, (1 << 24) // max(K, 1)))
    for r0 in range(0, M, step):
        r1 = min(M, r0 + step)
        mask = rng.random((r1 - r0, K), dtype=np.float32) < density
        counts[r0:r1] = mask.sum(axis=1)
        _, c = np.nonzero(mask)
        cols.append(c.astype(np.int32))
        n = c.size
        v = rng.standard_normal(n, dtype=np.float32) if signed else rng.random(n, dtype=np.float32)
        v[v == 0] = 0.5  # keep the pattern canonical (no explicit zeros)
        vals.append(v)
    row_ptr = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    assert row_ptr[-1] < 2**31
    col = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    val = np.concatenate(vals) if vals else np.zeros(0, np.float32)
    return row_ptr.astype(np.int32), col, val


def rmat_csr(scale: int = 20, edge_factor: int = 16, abc=(0.57, 0.19, 0.19), seed: int = 0,
             signed: bool = False):
    """R-MAT matrix 2^scale x 2^scale with edge_factor * 2^scale sampled edges (duplicates merged).

    scale 20 / edge factor 16 / seed 0 is BASELINE config 4 (about 16.09 M unique non-zeros, 48 % empty
    rows, longest row about 40 k non-zeros).  No vertex permutation: row 0 is the heaviest."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = edge_factor * n
    a, b, c = abc
    row = np.zeros(m, dtype=np.int64)
    col = np.zeros(m, dtype=np.int64)
    for _ in range(scale):
        r = rng.random(m, dtype=np.float32)
        # quadrant: a -> (0,0), b -> (0,1), c -> (1,0), d -> (1,1)
        rb = (r >= a + b).astype(np.int64)
        cb = (((r >= a) & (r < a + b)) | (r >= a + b + c)).astype(np.int64)
        row = (row << 1) | rb
        col = (col << 1) | cb
    key = np.unique((row << scale) | col)  # sorted by (row, col), duplicates dropped
    row = (key >> scale).astype(np.int64)
    col = (key & (n - 1)).astype(np.int32)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(row, minlength=n), out=row_ptr[1:])
    nnz = key.size
    val = rng.standard_normal(nnz, dtype=np.float32) if signed else rng.random(nnz, dtype=np.float32)
    val[val == 0] = 0.5
    return row_ptr.astype(np.int32), col, val, (n, n)


def dense_operand(K: int, N: int, seed: int = 1, signed: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.standard_normal((K, N), dtype=np.float32) if signed else rng.random((K, N), dtype=np.float32)


def spmm_algorithmic_bytes(M: int, K: int, nnz: int, N: int) -> int:
    """row_ptr + (col,val) + B read once + C written once."""
    return 4 * (M + 1) + 8 * nnz + 4 * K * N + 4 * M * N


def sddmm_algorithmic_bytes(M: int, K: int, nnz: int, N: int) -> int:
    return 4 * (M + 1) + 4 * nnz + 4 * M * N + 4 * K * N + 4 * nnz
