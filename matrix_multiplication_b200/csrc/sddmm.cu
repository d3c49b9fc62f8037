// sddmm.cu — grad_A on A's sparsity pattern:
//     out[p] = sum_j G[row(p), j] * B[col(p), j]        for every non-zero p of A.
//
// This is the backward of C = A*B with respect to the CSR values (reference
// matmuls.py:250 computes a dense grad_C * m2^T instead, on a mis-strided view;
// the engine evaluates the intended formula only where A has non-zeros).
//
// Same merge-path task cut as the forward (heavy threshold 0: every row may be
// split, no partial sums cross tasks because each non-zero is independent).
// A virtual warp of LPR lanes spans the dense dimension; it keeps the current
// G row slab in registers, gathers B rows with 128-bit loads and reduces the
// per-lane partial dot products with a transposing warp-shuffle reduction
// (U values across LPR lanes in ~U + log2 steps instead of U*log2).
#include "common.cuh"

namespace b200 {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SD_WARPS = 8;

template <int LPR, int VEC>
__global__ void __launch_bounds__(SD_WARPS * 32)
sddmm_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int M, int nnz,
             const float* __restrict__ G, long long ldg, const float* __restrict__ B,
             long long ldb, int N, float* __restrict__ out, int W, int n_tasks) {
  constexpr int GROUPS = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR;
  long long t = ((long long)blockIdx.x * SD_WARPS + warp) * GROUPS + grp;
  if (t > n_tasks) t = n_tasks;
  long long t1 = t + 1;
  if (t1 > n_tasks) t1 = n_tasks;
  const Boundary b0 = merge_path_boundary(row_ptr, M, nnz, t * (long long)W, 0);
  const Boundary b1 = merge_path_boundary(row_ptr, M, nnz, t1 * (long long)W, 0);
  const int j0 = b0.nz, j1 = b1.nz;
  int row = b0.row;
  int row_end = (row < M) ? __ldg(row_ptr + row + 1) : 0x7fffffff;
  const int slab_w = LPR * VEC;
  const int slabs = (N + slab_w - 1) / slab_w;

  int n_iter = (j1 - j0 + LPR - 1) / LPR;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) n_iter = max(n_iter, __shfl_xor_sync(FULL, n_iter, o));

  for (int it = 0; it < n_iter; ++it) {
    const int chunk = j0 + it * LPR;
    const int idx = chunk + gl;
    const int creg = (idx < j1) ? __ldg(col + idx) : 0;
#pragma unroll 4
    for (int k = 0; k < LPR; ++k) {
      const int pp = chunk + k;
      const int c = __shfl_sync(FULL, creg, k, LPR);
      const bool valid = pp < j1;
      float part = 0.f;
      if (valid) {
        while (pp >= row_end) {
          ++row;
          row_end = __ldg(row_ptr + row + 1);
        }
        const float* g = G + (long long)row * ldg + gl * VEC;
        const float* b = B + (long long)c * ldb + gl * VEC;
        for (int s = 0; s < slabs; ++s) {
          if (s * slab_w + gl * VEC < N) {
            Vec<VEC> gv, bv;
            gv.load(g + s * slab_w);
            bv.load(b + s * slab_w);
#pragma unroll
            for (int v = 0; v < VEC; ++v) part = fmaf(gv.v[v], bv.v[v], part);
          }
        }
      }
#pragma unroll
      for (int o = LPR / 2; o >= 1; o >>= 1) part += __shfl_xor_sync(FULL, part, o, LPR);
      if (valid && gl == 0) out[pp] = part;
    }
  }
}

template <int LPR, int VEC>
int launch_one(const int* row_ptr, const int* col, int M, int nnz, const float* G, long long ldg,
               const float* B, long long ldb, int N, float* out, cudaStream_t s) {
  constexpr int GROUPS = 32 / LPR;
  const int W = 256;
  const long long items = (long long)M + nnz;
  const int n_tasks = (int)((items + W - 1) / W);
  const int per_cta = SD_WARPS * GROUPS;
  dim3 grid((unsigned)((n_tasks + per_cta - 1) / per_cta));
  sddmm_kernel<LPR, VEC><<<grid, SD_WARPS * 32, 0, s>>>(row_ptr, col, M, nnz, G, ldg, B, ldb, N,
                                                       out, W, n_tasks);
  B200_LAUNCH_OK("sddmm_kernel");
  return B200_OK;
}

inline bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace

int launch_sddmm(const int* row_ptr, const int* col, int M, int nnz, const float* G, long long ldg,
                 const float* B, long long ldb, int N, float* out, cudaStream_t s) {
  if (nnz == 0 || M == 0) return B200_OK;
  if (N == 0) {
    B200_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)nnz * sizeof(float), s));
    return B200_OK;
  }
  const bool a4 = N % 4 == 0 && is_aligned(G, 16) && is_aligned(B, 16) && ldg % 4 == 0 && ldb % 4 == 0;
  const bool a2 = N % 2 == 0 && is_aligned(G, 8) && is_aligned(B, 8) && ldg % 2 == 0 && ldb % 2 == 0;
  if (a4) {
    if (N <= 16) return launch_one<4, 4>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
    if (N <= 32) return launch_one<8, 4>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
    if (N <= 64) return launch_one<16, 4>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
    return launch_one<32, 4>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
  }
  if (a2) return launch_one<32, 2>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
  return launch_one<32, 1>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
}

}  // namespace b200
