// spmm_slab.cu — forward CSR SpMM for narrow dense operands (N = 4, 8 or 16 columns) whose B fits a few shared-memory
// slabs: C[M,N] = A_csr[M,K] * B[K,N], fp32.
//
// Why: with N <= 16 a B row is 16-64 bytes; gathering it from global memory costs one L1 tag lookup (and usually an L2
// round trip: B does not fit L1) per NON-ZERO, and the lane-group kernel of spmm_fwd.cu runs at ~15 % of the HBM
// roof on config 5 x 16 (profiles/r2_ncu_fwd_n16.md).  Here the columns of A are cut into K-slabs; a persistent CTA
// owns an nnz-balanced range of rows and walks the slabs in order: the slab's rows of B are brought into shared
// memory once per CTA by a TMA bulk copy (double buffered: the next slab lands while this one is consumed), the
// non-zeros of every owned row that fall into the slab are gathered from shared memory with LDS.128 (no tags, no L2),
// and the rows' running sums stay in shared memory until the last slab.  col/val are streamed from HBM exactly once.
//
// Replaces spmm_kernel for this regime (reference src/naive_sparse_mm.cu:60-92: one warp per row, one 128-byte B
// transaction per non-zero and 32-column slab).  Rows must be sorted by column (CSR convention of the engine).
#include <algorithm>

#include "common.cuh"

namespace b200 {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SLAB_WARPS = 32;
constexpr int SLAB_ROWS_MAX = 1024;  // rows of C a CTA accumulates at a time (32 per warp; 64 KB of sums at N = 16)

// slab_ptr[row * (S + 1) + s] = first non-zero of `row` whose column is >= s * rows_per_slab
__global__ void slab_ptr_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int M, int S,
                                int rows_per_slab, int* __restrict__ slab_ptr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * (S + 1)) return;
  const int row = (int)(i / (S + 1)), s = (int)(i % (S + 1));
  int lo = __ldg(row_ptr + row), hi = __ldg(row_ptr + row + 1);
  const int key = s * rows_per_slab;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(col + mid) < key) lo = mid + 1; else hi = mid;
  }
  slab_ptr[i] = lo;
}

// cta_rows[c] = first row of CTA c: rows cut so that every CTA owns about (nnz + 4 * M) / G units of work
__global__ void slab_rows_kernel(const int* __restrict__ row_ptr, int M, int nnz, int G, int* __restrict__ cta_rows) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > G) return;
  const long long target = ((long long)nnz + 4LL * M) * c / G;
  int lo = 0, hi = M;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long long)__ldg(row_ptr + mid) + 4LL * mid < target) lo = mid + 1; else hi = mid;
  }
  cta_rows[c] = (c == G) ? M : lo;
}

template <int LPR>  // lanes per non-zero: N = 4 * LPR columns
__global__ void __launch_bounds__(SLAB_WARPS * 32, 1)
spmm_slab_kernel(const int* __restrict__ col, const float* __restrict__ val, const int* __restrict__ slab_ptr,
                 const int* __restrict__ cta_rows, const float* __restrict__ B, int K, float* __restrict__ C,
                 long long ldc, int S, int rows_per_slab, const float* __restrict__ bias, int relu) {
  constexpr int N = 4 * LPR;
  constexpr int G = 32 / LPR;  // non-zeros a warp consumes per sub-step
  extern __shared__ __align__(128) unsigned char slab_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(slab_smem);                     // [2]
  float* s_acc = reinterpret_cast<float*>(slab_smem + 128);                    // [SLAB_ROWS_MAX][N]
  const size_t slab_bytes = (size_t)rows_per_slab * N * sizeof(float);
  float* const slab_buf0 = s_acc + (size_t)SLAB_ROWS_MAX * N;
  float* const slab_buf1 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(slab_buf0) + slab_bytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane / LPR, l = lane % LPR;
  const int R0 = __ldg(cta_rows + blockIdx.x), R1 = __ldg(cta_rows + blockIdx.x + 1);
  if (tid == 0) {
    mbar_init(full + 0, 1);
    mbar_init(full + 1, 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t phase0 = 0u, phase1 = 0u;
  auto issue_slab = [&](int s, int b) {  // thread 0: rows [s * rps, min(K, (s+1) * rps)) of B -> buffer b
    const int k0 = s * rows_per_slab;
    const uint32_t bytes = (uint32_t)(min(rows_per_slab, K - k0)) * N * (uint32_t)sizeof(float);
    mbar_expect_tx(full + b, bytes);
    tma_bulk_g2s(b ? slab_buf1 : slab_buf0, B + (size_t)k0 * N, bytes, full + b);
  };

  for (int c0 = R0; c0 < R1; c0 += SLAB_ROWS_MAX) {  // one pass over the slabs per block of owned rows
    const int c1 = min(c0 + SLAB_ROWS_MAX, R1);
    for (int i = tid; i < (c1 - c0) * N; i += blockDim.x) s_acc[i] = 0.f;
    if (tid == 0) issue_slab(0, 0);
    __syncthreads();
    for (int s = 0; s < S; ++s) {
      const int b = s & 1;
      if (tid == 0 && s + 1 < S) issue_slab(s + 1, b ^ 1);  // buffer b^1 was released by the barrier ending slab s-1
      if (b) { mbar_wait(full + 1, phase1); phase1 ^= 1u; } else { mbar_wait(full + 0, phase0); phase0 ^= 1u; }
      const float* slab = (b ? slab_buf1 : slab_buf0) - (size_t)s * rows_per_slab * N;  // indexed by the global column
      // this warp's rows: c0 + warp + 32 i (at most 32 of them); lane i fetches the bounds of row i's segment in this
      // slab once, the loop below gets them by shuffle: no dependent global load between two batches
      const int nvis = (c1 - c0 - warp + SLAB_WARPS - 1) / SLAB_WARPS;
      int myq0 = 0, myq1 = 0;
      if (lane < nvis) {
        const size_t r = (size_t)(c0 + warp + SLAB_WARPS * lane);
        myq0 = __ldg(slab_ptr + r * (S + 1) + s);
        myq1 = __ldg(slab_ptr + r * (S + 1) + s + 1);
      }
      // flat walk over the batches (64 non-zeros) of all segments, the next batch's loads in flight while the current
      // one is consumed
      int vi = -1, base = 0, q1 = 0;
      auto advance = [&](int& vi_, int& base_, int& q1_) -> bool {
        base_ += 64;
        while (base_ >= q1_) {
          if (++vi_ >= nvis) return false;
          base_ = __shfl_sync(FULL, myq0, vi_);
          q1_ = __shfl_sync(FULL, myq1, vi_);
        }
        return true;
      };
      auto load = [&](int base_, int q1_, int (&cc)[2], float (&aa)[2]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int q = base_ + u * 32 + lane;
          cc[u] = (q < q1_) ? __ldcs(col + q) : -1;
          aa[u] = (q < q1_) ? __ldcs(val + q) : 0.f;
        }
      };
      int cc[2], ncc[2];
      float aa[2], naa[2];
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      bool have = advance(vi, base, q1);
      if (have) load(base, q1, cc, aa);
      while (have) {
        int nvi = vi, nbase = base, nq1 = q1;
        const bool nhave = advance(nvi, nbase, nq1);
        if (nhave) load(nbase, nq1, ncc, naa);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (base + u * 32 < q1) {
#pragma unroll
            for (int sub = 0; sub < LPR; ++sub) {  // group g takes non-zero sub * G + g of the batch
              const int cj = __shfl_sync(FULL, cc[u], sub * G + g);
              const float aj = __shfl_sync(FULL, aa[u], sub * G + g);
              if (cj >= 0) {
                const float4 bv = *reinterpret_cast<const float4*>(slab + (size_t)cj * N + 4 * l);
                acc[0] = fmaf(aj, bv.x, acc[0]);
                acc[1] = fmaf(aj, bv.y, acc[1]);
                acc[2] = fmaf(aj, bv.z, acc[2]);
                acc[3] = fmaf(aj, bv.w, acc[3]);
              }
            }
          }
        }
        if (base + 64 >= q1) {
          // last batch of the segment: combine the G groups (fixed tree: deterministic), group 0 adds into the row's sum
#pragma unroll
          for (int o = 16; o >= LPR; o >>= 1) {
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[v] += __shfl_xor_sync(FULL, acc[v], o);
          }
          if (g == 0) {
            float4* dst = reinterpret_cast<float4*>(s_acc + (size_t)(warp + SLAB_WARPS * vi) * N + 4 * l);
            float4 t = *dst;
            t.x += acc[0]; t.y += acc[1]; t.z += acc[2]; t.w += acc[3];
            *dst = t;
          }
          acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        }
        vi = nvi; base = nbase; q1 = nq1; have = nhave;
#pragma unroll
        for (int u = 0; u < 2; ++u) { cc[u] = ncc[u]; aa[u] = naa[u]; }
      }
      __syncthreads();  // slab s consumed by every warp: its buffer may be refilled
    }
    // the finished rows: + bias, ReLU, coalesced 128-bit stores
    for (int i = tid; i < (c1 - c0) * LPR; i += blockDim.x) {
      const int r = c0 + i / LPR, c4 = i % LPR;
      float4 t = *reinterpret_cast<const float4*>(s_acc + (size_t)(r - c0) * N + 4 * c4);
      if (bias != nullptr) {
        const float bb = __ldg(bias + r);
        t.x += bb; t.y += bb; t.z += bb; t.w += bb;
      }
      if (relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
      __stcs(reinterpret_cast<float4*>(C + (long long)r * ldc + 4 * c4), t);
    }
    __syncthreads();
  }
}

size_t slab_smem_bytes(int rows_per_slab, int N) {
  return 128 + (size_t)SLAB_ROWS_MAX * N * sizeof(float) + 2 * (size_t)rows_per_slab * N * sizeof(float);
}

}  // namespace

// rows of B per slab for N columns: two slab buffers + the row sums within 220 KB
int slab_rows_per_slab(int N) {
  const size_t avail = 220 * 1024 - 128 - (size_t)SLAB_ROWS_MAX * N * sizeof(float);
  return (int)((avail / 2) / ((size_t)N * sizeof(float))) & ~7;
}

// Is the slab kernel the right one for this product?  (narrow N, B in a few slabs, enough non-zeros per visit)
bool slab_wanted(int M, int K, long long nnz, int N, const float* B, long long ldb, const float* C, long long ldc) {
  if (!(N == 4 || N == 8 || N == 16) || ldb != N || ldc % 4 != 0) return false;
  if (reinterpret_cast<uintptr_t>(B) % 16 != 0 || reinterpret_cast<uintptr_t>(C) % 16 != 0) return false;
  const int rps = slab_rows_per_slab(N);
  const long long S = ((long long)K + rps - 1) / rps;
  if (S > 16 || nnz < (1LL << 20) || M <= 0) return false;
  // at least ~48 non-zeros per (row, slab) visit: below that the per-visit bookkeeping (segment bounds, group
  // reduction, running-sum update) outweighs the cheaper gathers (measured: 16384^2 @99 % x 16, 12 per visit, is
  // 2x slower here than on the lane-group kernel; @90 %, 126 per visit, 1.3x faster; profiles/r2_ncu_fwd_n16.md)
  return nnz >= 48LL * M * S;
}

void slab_store_free(SlabStore* st) {
  cudaFree(st->slab_ptr); cudaFree(st->cta_rows);
  *st = SlabStore{};
}

int slab_store_build(const int* row_ptr, const int* col, int M, int K, int nnz, int N, SlabStore* st, cudaStream_t s) {
  *st = SlabStore{};
  st->N = N; st->M = M; st->K = K;
  st->rows_per_slab = slab_rows_per_slab(N);
  st->S = (K + st->rows_per_slab - 1) / st->rows_per_slab;
  st->G = device_sm_count();
  const size_t cells = (size_t)M * (st->S + 1);
  B200_CUDA_OK(cudaMalloc(&st->slab_ptr, cells * sizeof(int)));
  B200_CUDA_OK(cudaMalloc(&st->cta_rows, ((size_t)st->G + 1) * sizeof(int)));
  slab_ptr_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(row_ptr, col, M, st->S, st->rows_per_slab, st->slab_ptr);
  B200_LAUNCH_OK("slab_ptr_kernel");
  slab_rows_kernel<<<(st->G + 256) / 256, 256, 0, s>>>(row_ptr, M, nnz, st->G, st->cta_rows);
  B200_LAUNCH_OK("slab_rows_kernel");
  return B200_OK;
}

int launch_spmm_slab(const SlabStore& st, const int* col, const float* val, const float* B, float* C, long long ldc,
                     const float* bias, int relu, cudaStream_t s) {
  const size_t smem = slab_smem_bytes(st.rows_per_slab, st.N);
#define B200_SLAB(L)                                                                                             \
  do {                                                                                                           \
    static bool attr[64] = {false};                                                                              \
    int dev = 0;                                                                                                 \
    B200_CUDA_OK(cudaGetDevice(&dev));                                                                           \
    if (!attr[dev & 63]) {                                                                                       \
      B200_CUDA_OK(cudaFuncSetAttribute(spmm_slab_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr[dev & 63] = true;                                                                                     \
    }                                                                                                            \
    spmm_slab_kernel<L><<<st.G, SLAB_WARPS * 32, smem, s>>>(col, val, st.slab_ptr, st.cta_rows, B, st.K, C, ldc, st.S, \
                                                          st.rows_per_slab, bias, relu);                         \
  } while (0)
  if (st.N == 16) B200_SLAB(4);
  else if (st.N == 8) B200_SLAB(2);
  else B200_SLAB(1);
#undef B200_SLAB
  B200_LAUNCH_OK("spmm_slab_kernel");
  return B200_OK;
}

}  // namespace b200
