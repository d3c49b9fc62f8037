"""Pure-numpy model of the forward kernel's work decomposition (spmm_fwd.cu): merge-path tasks,
row-start snapping for light rows, exact splits with 'pieces' for heavy rows, fix-up pass.
Uses the library's own host boundary function, so the arithmetic of the cut is the shipped one.
Validates the ALGORITHM on the CPU; the CUDA code itself is validated by the -m gpu tests."""
import numpy as np

from matrix_multiplication_b200 import capi


def simulate(row_ptr, col, val, shape, B, W, heavy=None):
    M, K = shape
    heavy = W if heavy is None else heavy
    nnz = int(row_ptr[-1])
    N = B.shape[1]
    n_tasks = max(1, -(-(M + nnz) // W))
    bnd = [capi.merge_path_boundary(row_ptr, t * W, heavy) for t in range(n_tasks + 1)]
    C = np.full((M, N), np.nan, dtype=np.float64)        # NaN = never written
    written = np.zeros(M, dtype=np.int64)
    pieces = {}
    B64 = B.astype(np.float64)
    for t in range(n_tasks):
        (i0, j0), (i1, j1) = bnd[t], bnd[t + 1]
        assert (i0, j0) <= (i1, j1) and j0 <= j1
        row = i0
        head = i0 < M and j0 > row_ptr[i0]
        row_end = row_ptr[row + 1] if row < M else 2**31 - 1
        acc = np.zeros(N)

        def flush(r):
            nonlocal head, acc
            if head:
                pieces[2 * t] = acc.copy()
            else:
                C[r] = acc
                written[r] += 1
            head = False
            acc = np.zeros(N)

        for p in range(j0, j1):
            while p == row_end:
                flush(row)
                row += 1
                row_end = row_ptr[row + 1]
            acc = acc + float(val[p]) * B64[col[p]]
        while row < i1:
            flush(row)
            row += 1
        if i1 < M and j1 > j0 and j1 > row_ptr[i1]:
            pieces[2 * t + (0 if head else 1)] = acc.copy()
    # fix-up
    for b in range(1, n_tasks):
        i, j = bnd[b]
        if i >= M or j <= row_ptr[i]:
            continue
        if bnd[b + 1][0] == i:
            continue
        t_first = (int(row_ptr[i]) + i) // W
        s = pieces[2 * t_first + 1].copy()
        for t in range(t_first + 1, b + 1):
            s = s + pieces[2 * t]
        C[i] = s
        written[i] += 1
    return C, written, n_tasks, len(pieces)
