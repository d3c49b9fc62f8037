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


def simulate_flat(row_ptr, col, val, shape, B, W, NW=4):
    """Numpy model of spmm_flat_kernel + partition/fix list + spmm_fixup_kernel (spmm_fwd.cu): CTAs of NW warp
    tasks, aligned chunks of 8 non-zeros, per-chunk row-end masks built from row_ptr, edge masks at task
    boundaries, CTA-level zero fill of empty rows, head / tail pieces and the fix list.  Returns
    (C, times_written_per_row, n_fix_entries)."""
    M, K = shape
    heavy = W
    nnz = int(row_ptr[-1])
    N = B.shape[1]
    n_tasks = max(1, -(-(M + nnz) // W))
    bnd = [capi.merge_path_boundary(row_ptr, t * W, heavy) for t in range(n_tasks + 1)]
    # partition_kernel's fix list: boundaries inside a row that ends in the task starting there
    fix = [t for t in range(1, n_tasks)
           if bnd[t][0] < M and bnd[t][1] > row_ptr[bnd[t][0]] and bnd[t + 1][0] != bnd[t][0]]
    C = np.full((M, N), np.nan)
    written = np.zeros(M, dtype=np.int64)
    pieces = {}
    B64 = B.astype(np.float64)
    for t0 in range(0, n_tasks, NW):
        tb = [bnd[min(t0 + i, n_tasks)] for i in range(NW + 1)]
        a0 = tb[0][1] & ~7
        je = tb[NW][1]
        mask = {}       # chunk index -> 8-bit row-end mask
        rowend = {}     # staged offset -> row id
        R0, R1 = tb[0][0], tb[NW][0]
        for r in range(R0, R1):                      # rows that END inside this CTA
            rs, re = int(row_ptr[r]), int(row_ptr[r + 1])
            if re > rs:
                pos = re - 1 - a0
                assert 0 <= pos < je - a0, "row end outside the staged window"
                mask[pos >> 3] = mask.get(pos >> 3, 0) | (1 << (pos & 7))
                rowend[pos] = r
            else:
                C[r] = 0.0
                written[r] += 1
        for w in range(NW):
            t = t0 + w
            if t >= n_tasks:
                break
            (i0, j0), (i1, j1) = tb[w], tb[w + 1]
            if j1 <= j0:
                continue
            head = j0 > row_ptr[i0]
            acc = np.zeros(N)
            for c0 in range(j0 & ~7, j1, 8):
                off = c0 - a0
                m = mask.get(off >> 3, 0)
                lo, hi = max(j0 - c0, 0), min(j1 - c0, 8)
                vm = ((1 << hi) - 1) & ~((1 << lo) - 1)
                m &= vm
                for k in range(8):
                    if (vm >> k) & 1:
                        p = c0 + k
                        acc = acc + float(val[p]) * B64[col[p]]
                    if (m >> k) & 1:
                        if head:
                            pieces[2 * t] = acc.copy()
                            head = False
                        else:
                            r = rowend[off + k]
                            C[r] = acc
                            written[r] += 1
                        acc = np.zeros(N)
            last = j1 - 1 - a0
            if not ((mask.get(last >> 3, 0) >> (last & 7)) & 1):
                pieces[2 * t + (0 if head else 1)] = acc.copy()
    for b in fix:                                     # spmm_fixup_kernel
        i = bnd[b][0]
        t_first = (int(row_ptr[i]) + i) // W
        s = pieces[2 * t_first + 1].copy()
        for k in range(1, b - t_first + 1):
            s = s + pieces[2 * (t_first + k)]
        C[i] = s
        written[i] += 1
    return C, written, len(fix)
