// sddmm.cu — grad_A on A's sparsity pattern:
//     out[p] = sum_j G[row(p), j] * B[col(p), j]        for every non-zero p of A.
//
// This is the backward of C = A*B with respect to the CSR values (reference
// matmuls.py:250 computes a dense grad_C * m2^T instead, on a mis-strided view;
// the engine evaluates the intended formula only where A has non-zeros).
//
// Same shape as the flat forward kernel (spmm_fwd.cu): merge-path tasks (heavy threshold 0: any
// row may be cut, non-zeros are independent so nothing crosses tasks), one warp per task, the
// CTA's `col` segment staged by a TMA bulk copy, and the row id of every staged non-zero expanded
// into shared memory once per CTA.  A warp walks aligned chunks of 8 non-zeros: 16-byte broadcast
// reads of (col, row), 8 independent 128-bit gathers of B rows — plus 8 of G rows only when the
// chunk is not entirely inside the row whose G slab is already in registers — VEC FMAs per lane
// and non-zero, and one transposing shuffle reduction per chunk (9 SHFL for 8 dot products
// instead of 40).  Results leave as one 32-byte store per chunk.
#include "common.cuh"

namespace b200 {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SD_NW = 4;    // warps (= tasks) per CTA
constexpr int SD_W = 256;   // items per task (smaller for small problems, see sd_items)

__host__ __device__ inline int sd_capacity(int W) { return ((SD_NW * W + 24) + 7) & ~7; }

__global__ void __launch_bounds__(256)
sddmm_partition_kernel(const int* __restrict__ row_ptr, int M, int nnz, int W, int n_tasks,
                       Boundary* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= n_tasks) out[t] = merge_path_boundary(row_ptr, M, nnz, (long long)t * W, 0);
}

template <int VEC, bool TMA, bool MULTI>
__global__ void __launch_bounds__(SD_NW * 32, MULTI ? 3 : 8)
sddmm_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int M, int nnz,
             const float* __restrict__ G, unsigned ldg, const float* __restrict__ B, unsigned ldb, int N,
             float* __restrict__ out, const Boundary* __restrict__ bnd, int W, int n_tasks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [mbar 16B][Boundary x (NW+1), 48B][s_col CAP][s_row CAP]
  const int cap = sd_capacity(W);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  Boundary* bnd_s = reinterpret_cast<Boundary*>(smem_raw + 16);
  constexpr int BND_BYTES = ((SD_NW + 1) * (int)sizeof(Boundary) + 15) & ~15;
  int* s_col = reinterpret_cast<int*>(smem_raw + 16 + BND_BYTES);
  int* s_row = s_col + cap;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int task0 = blockIdx.x * SD_NW;
  if (tid <= SD_NW) bnd_s[tid] = bnd[min(task0 + tid, n_tasks)];
  if (TMA && tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int a0 = bnd_s[0].nz & ~7, je = bnd_s[SD_NW].nz;
  const int n_bulk = TMA ? ((je - a0) & ~3) : 0;
  if (TMA && tid == 0 && n_bulk > 0) {
    mbar_expect_tx(bar, (uint32_t)n_bulk * 4u);
    tma_bulk_g2s(s_col, col + a0, (uint32_t)n_bulk * 4u, bar);
  }
  for (int q = a0 + n_bulk + tid; q < je; q += blockDim.x) s_col[q - a0] = __ldg(col + q);

  // row id of every staged non-zero: rows [R0, R1] intersect the CTA's range [j0c, je)
  {
    const int j0c = bnd_s[0].nz;
    const int R0 = bnd_s[0].row, R1 = min(bnd_s[SD_NW].row, M - 1);
    for (int r = R0 + warp; r <= R1; r += SD_NW) {
      const int rs = max(__ldg(row_ptr + r), j0c), re = min(__ldg(row_ptr + r + 1), je);
      for (int q = rs + lane; q < re; q += 32) s_row[q - a0] = r;
    }
  }
  __syncthreads();
  if (TMA && n_bulk > 0) mbar_wait(bar, 0);

  const Boundary b0 = bnd_s[warp], b1 = bnd_s[warp + 1];
  const int j0 = b0.nz, j1 = b1.nz;
  if (j1 <= j0) return;

  const int slab_w = 32 * VEC;
  const int n_slabs = MULTI ? (N + slab_w - 1) / slab_w : 1;  // MULTI: N wider than one slab of 32 * VEC columns
  int cur_row = -1;   // row whose G slab is cached in gcur (reused across chunks only when n_slabs == 1)
  Vec<VEC> gcur;
#pragma unroll
  for (int v = 0; v < VEC; ++v) gcur.v[v] = 0.f;

  for (int c0 = j0 & ~7; c0 < j1; c0 += 8) {
    const int off = c0 - a0;
    const int4 ca = *reinterpret_cast<const int4*>(s_col + off);
    const int4 cb = *reinterpret_cast<const int4*>(s_col + off + 4);
    const int c[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
    const int r0 = s_row[off], r7 = s_row[off + 7];  // rows of the first / last position (garbage outside vm)
    const int lo = max(j0 - c0, 0), hi = min(j1 - c0, 8);
    const unsigned vm = ((1u << hi) - 1u) & ~((1u << lo) - 1u);  // positions of this task
    float d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = 0.f;

    const bool one_row = vm == 0xffu && r0 == r7;  // rows are sorted: the whole chunk is in one row
    if (MULTI && one_row) {
      // wide N: walk the slabs two deep — the gathers of slab s+1 are in flight while slab s is consumed
      auto load_slab = [&](int s, Vec<VEC>(&b)[8], Vec<VEC>& g) {
        const int col0 = s * slab_w + lane * VEC;
        const bool active = col0 < N;
        const int cc = active ? col0 : 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k].load(B + (unsigned long long)(unsigned)c[k] * ldb + cc);
        g.load(G + (unsigned long long)(unsigned)r0 * ldg + cc);
        if (!active) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) g.v[v] = 0.f;
        }
      };
      auto fma_slab = [&](const Vec<VEC>(&b)[8], const Vec<VEC>& g) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int v = 0; v < VEC; ++v) d[k] = fmaf(g.v[v], b[k].v[v], d[k]);
      };
      Vec<VEC> b0[8], b1[8], g0, g1;
      load_slab(0, b0, g0);
      for (int s = 0; s < n_slabs; s += 2) {
        if (s + 1 < n_slabs) load_slab(s + 1, b1, g1);
        fma_slab(b0, g0);
        if (s + 2 < n_slabs) load_slab(s + 2, b0, g0);
        if (s + 1 < n_slabs) fma_slab(b1, g1);
      }
      cur_row = -1;
    } else
    for (int s = 0; s < n_slabs; ++s) {
      const int col0 = s * slab_w + lane * VEC;
      const bool active = col0 < N;
      const int cc = active ? col0 : 0;  // inactive lanes read column 0 and contribute 0
      Vec<VEC> b[8];
      if (one_row) {  // one G slab serves all 8 non-zeros (kept in registers across chunks when N fits one slab)
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k].load(B + (unsigned long long)(unsigned)c[k] * ldb + cc);
        if (n_slabs > 1 || r0 != cur_row) {
          gcur.load(G + (unsigned long long)(unsigned)r0 * ldg + cc);
          if (!active) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) gcur.v[v] = 0.f;
          }
          cur_row = r0;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int v = 0; v < VEC; ++v) d[k] = fmaf(gcur.v[v], b[k].v[v], d[k]);
      } else {  // mixed chunk (row ends inside it, or a task edge): one G gather per non-zero, in two halves
        // of 4 so that only 4 B slabs + 4 G slabs are live at a time (register budget = occupancy here)
        Vec<VEC> g4[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          Vec<VEC> b4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int kk = 4 * h + k;
            if ((vm >> kk) & 1u) {
              b4[k].load(B + (unsigned long long)(unsigned)c[kk] * ldb + cc);
              g4[k].load(G + (unsigned long long)(unsigned)s_row[off + kk] * ldg + cc);
            } else {
#pragma unroll
              for (int v = 0; v < VEC; ++v) { b4[k].v[v] = 0.f; g4[k].v[v] = 0.f; }
            }
          }
          if (!active) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int v = 0; v < VEC; ++v) g4[k].v[v] = 0.f;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int v = 0; v < VEC; ++v) d[4 * h + k] = fmaf(g4[k].v[v], b4[k].v[v], d[4 * h + k]);
        }
        if (n_slabs == 1 && ((vm >> 7) & 1u)) {
          cur_row = r7;
          gcur = g4[3];
        } else {
          cur_row = -1;
        }
      }
    }
    // transposing reduction: 8 per-lane partials x 32 lanes -> lane l (l % 4 == 0) holds dot (l >> 2)
    float e[4], f[2];
    {
      const bool up = lane & 16;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float send = up ? d[k] : d[k + 4];
        const float keep = up ? d[k + 4] : d[k];
        e[k] = keep + __shfl_xor_sync(FULL, send, 16);
      }
    }
    {
      const bool up = lane & 8;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float send = up ? e[k] : e[k + 2];
        const float keep = up ? e[k + 2] : e[k];
        f[k] = keep + __shfl_xor_sync(FULL, send, 8);
      }
    }
    float h;
    {
      const bool up = lane & 4;
      const float send = up ? f[0] : f[1];
      const float keep = up ? f[1] : f[0];
      h = keep + __shfl_xor_sync(FULL, send, 4);
    }
    h += __shfl_xor_sync(FULL, h, 2);
    h += __shfl_xor_sync(FULL, h, 1);
    const int k_mine = lane >> 2;
    if ((lane & 3) == 0 && ((vm >> k_mine) & 1u)) out[c0 + k_mine] = h;
  }
}

int sd_items(long long M, long long nnz);

// ---- small N (< 24): G = 32 / LPR lane groups, each walking its own merge-path task, LPR lanes across the
// dense dimension; boundaries are searched in the kernel (tasks are tiny, no staging).
constexpr int SDS_WARPS = 8;

template <int LPR, int VEC>
__global__ void __launch_bounds__(SDS_WARPS * 32)
sddmm_grouped_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int M, int nnz,
             const float* __restrict__ G, long long ldg, const float* __restrict__ B,
             long long ldb, int N, float* __restrict__ out, int W, int n_tasks) {
  constexpr int GROUPS = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR;
  long long t = ((long long)blockIdx.x * SDS_WARPS + warp) * GROUPS + grp;
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
int launch_grouped(const int* row_ptr, const int* col, int M, int nnz, const float* G, long long ldg,
               const float* B, long long ldb, int N, float* out, cudaStream_t s) {
  constexpr int GROUPS = 32 / LPR;
  const int W = std::max(32, sd_items(M, nnz) / 2);  // a lane group, not a warp, walks a task here
  const long long items = (long long)M + nnz;
  const int n_tasks = (int)((items + W - 1) / W);
  const int per_cta = SDS_WARPS * GROUPS;
  dim3 grid((unsigned)((n_tasks + per_cta - 1) / per_cta));
  sddmm_grouped_kernel<LPR, VEC><<<grid, SDS_WARPS * 32, 0, s>>>(row_ptr, col, M, nnz, G, ldg, B, ldb, N,
                                                       out, W, n_tasks);
  B200_LAUNCH_OK("sddmm_grouped_kernel");
  return B200_OK;
}

inline bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <int VEC, bool TMA, bool MULTI>
int launch_one(const int* row_ptr, const int* col, int M, int nnz, const float* G, long long ldg,
               const float* B, long long ldb, int N, float* out, const Boundary* bnd, int n_tasks,
               cudaStream_t s) {
  const int W = sd_items(M, nnz);
  const int cap = sd_capacity(W);
  constexpr int BND_BYTES = ((SD_NW + 1) * (int)sizeof(Boundary) + 15) & ~15;
  const size_t smem = 16 + BND_BYTES + (size_t)cap * 8;
  dim3 grid((unsigned)((n_tasks + SD_NW - 1) / SD_NW));
  sddmm_kernel<VEC, TMA, MULTI><<<grid, SD_NW * 32, smem, s>>>(row_ptr, col, M, nnz, G, (unsigned)ldg, B,
                                                         (unsigned)ldb, N, out, bnd, W, n_tasks);
  B200_LAUNCH_OK("sddmm_kernel");
  return B200_OK;
}

}  // namespace

namespace {
int sd_items(long long M, long long nnz) {  // same rule as the forward: spread small problems over the SMs
  const long long spread = (M + nnz) / (148 * 96);
  const int W = (int)std::min<long long>(SD_W, std::max<long long>(32, spread));
  return (W + 31) & ~31;
}
}  // namespace
int sddmm_n_tasks(long long M, long long nnz) {
  const int W = sd_items(M, nnz);
  const long long n = (M + nnz + W - 1) / W;
  return (int)(n < 1 ? 1 : n);
}
size_t sddmm_bounds_bytes(long long M, long long nnz) {
  return (((size_t)sddmm_n_tasks(M, nnz) + 2) * sizeof(Boundary) + 255) & ~(size_t)255;
}
int launch_sddmm_partition(const int* row_ptr, int M, int nnz, Boundary* out, cudaStream_t s) {
  const int n_tasks = sddmm_n_tasks(M, nnz);
  sddmm_partition_kernel<<<(n_tasks + 1 + 255) / 256, 256, 0, s>>>(row_ptr, M, nnz, sd_items(M, nnz), n_tasks, out);
  B200_LAUNCH_OK("sddmm_partition_kernel");
  return B200_OK;
}

int launch_sddmm(const int* row_ptr, const int* col, int M, int nnz, const float* G, long long ldg,
                 const float* B, long long ldb, int N, float* out, const Boundary* cached_bnd,
                 cudaStream_t s) {
  if (nnz == 0 || M == 0) return B200_OK;
  if (N == 0) {
    B200_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)nnz * sizeof(float), s));
    return B200_OK;
  }
  const bool a4 = N % 4 == 0 && is_aligned(G, 16) && is_aligned(B, 16) && ldg % 4 == 0 && ldb % 4 == 0;
  const bool a2 = N % 2 == 0 && is_aligned(G, 8) && is_aligned(B, 8) && ldg % 2 == 0 && ldb % 2 == 0;
  if (N < 24) {  // lane groups: 32 / LPR non-zeros per gather instruction
    if (a4) {
      if (N <= 16) return launch_grouped<4, 4>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
      return launch_grouped<8, 4>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
    }
    if (a2) {
      if (N <= 8) return launch_grouped<4, 2>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
      if (N <= 16) return launch_grouped<8, 2>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
      return launch_grouped<16, 2>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
    }
    if (N <= 4) return launch_grouped<4, 1>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
    if (N <= 8) return launch_grouped<8, 1>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
    if (N <= 16) return launch_grouped<16, 1>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
    return launch_grouped<32, 1>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, s);
  }
  const Boundary* bnd = cached_bnd;
  if (bnd == nullptr) {
    void* w = nullptr;
    int rc = get_scratch(sddmm_bounds_bytes(M, nnz), &w, s);
    if (rc != B200_OK) return rc;
    rc = launch_sddmm_partition(row_ptr, M, nnz, reinterpret_cast<Boundary*>(w), s);
    if (rc != B200_OK) return rc;
    bnd = reinterpret_cast<const Boundary*>(w);
  }
  const int n_tasks = sddmm_n_tasks(M, nnz);
  const int vmax = a4 ? 4 : (a2 ? 2 : 1);
  int vec = 1;
  while (vec < vmax && 32 * vec < N) vec <<= 1;
  const bool tma = is_aligned(col, 16);
  const bool multi = N > 32 * vec;
#define B200_SD2(V, T, MU) launch_one<V, T, MU>(row_ptr, col, M, nnz, G, ldg, B, ldb, N, out, bnd, n_tasks, s)
#define B200_SD(V)                                                                       \
  if (vec == V)                                                                          \
    return tma ? (multi ? B200_SD2(V, true, true) : B200_SD2(V, true, false))            \
               : (multi ? B200_SD2(V, false, true) : B200_SD2(V, false, false))
  B200_SD(4);
  B200_SD(2);
  B200_SD(1);
#undef B200_SD2
#undef B200_SD
  return B200_ERR_INVALID;
}

}  // namespace b200
