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
#include <vector>

#include <cub/device/device_scan.cuh>

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


// ---- v2: sliced stream of (column, value) pairs in the order the kernel consumes it ------------------------------
// The inner loop of the kernel above spends 3.7 warp instructions per non-zero (two shuffles per 8 non-zeros to hand
// col / val to the lane groups, a validity branch per sub-step, the batch walk).  Here the plan keeps a second copy
// of A in exactly the order the kernel consumes it, so that the loop is one 128-bit read per lane and TWO steps:
//
//   slice    HR consecutive rows (HR a power of two <= G = 32 / LPR lane groups).  A warp works on one slice at a
//            time ("visit" = one slice in one slab): lane group g serves row g % HR; the SPLIT = G / HR groups of a row
//            take its non-zeros round-robin and are summed by a shuffle tree at the end of the visit (fixed order:
//            deterministic).
//   block    2 steps of a visit: G x (col0, val0, col1, val1), 16 bytes per lane group.  Short rows are padded with
//            (the zero row behind the slab, 0.0f).  Inside a row's segment the non-zeros are ordered by the bank class
//            of their column (a B row is 128 / nclass bytes of the 128-byte bank window), starting with the class of
//            the row's own group, so that the lane groups of one LDS.128 phase mostly hit different banks.
//   stream   the blocks of all visits in consumption order: [CTA][warp][pass][slab][visit] — every warp reads ONE
//            contiguous array from the first to the last block of the launch.  It comes in through a per-warp ring
//            of small chunks in shared memory filled by TMA bulk copies (lane 0 re-issues a chunk as soon as the
//            warp has left it), so the loads run ahead across visit, slab and pass boundaries.
//   vis_off  exclusive prefix of the blocks of every visit, in the same order; warp_vis[c * 32 + w] = first visit of
//            warp w of CTA c; cta_slices cuts the slices into G_sm ranges of equal work.
// Per block and lane: LDS.128 (stream) + 2 x (IMAD + LDS.128 + 4 FFMA) = 13 instructions for 2 G non-zeros.
__device__ __forceinline__ uint4 lds128_u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128_f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
constexpr int SLAB2_RING = 2;        // chunks per warp ring
__host__ __device__ constexpr int slab2_chunk_bytes(int N) { return (32 / (N / 4)) * 16 > 512 ? (32 / (N / 4)) * 16 : 512; }  // >= one block

__host__ __device__ inline int slab2_nclass(int N) { return 32 / N; }  // N = 16 / 8 / 4 -> 2 / 4 / 8

// steps of every (slice, slab): max over the slice's rows of ceil(len / SPLIT), rounded up to even
__global__ void slab2_steps_kernel(const int* __restrict__ slab_ptr, int M, int S, int HR, int SPLIT, int n_slices,
                                   int* __restrict__ steps) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx > (long long)n_slices * S) return;
  if (idx == (long long)n_slices * S) { steps[idx] = 0; return; }
  const int i = (int)(idx / S), s = (int)(idx % S);
  int mx = 0;
  for (int h = 0; h < HR; ++h) {
    const long long r = (long long)i * HR + h;
    if (r < M) {
      const int len = __ldg(slab_ptr + r * (S + 1) + s + 1) - __ldg(slab_ptr + r * (S + 1) + s);
      mx = max(mx, (len + SPLIT - 1) / SPLIT);
    }
  }
  steps[idx] = (mx + 1) & ~1;
}

// cta_slices[c] = first slice of CTA c: equal shares of (steps + a per-visit constant); step_off = slice-major prefix
__global__ void slab2_cta_kernel(const int* __restrict__ step_off, int n_slices, int S, int G, int* __restrict__ cta_slices) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > G) return;
  const long long total = (long long)__ldg(step_off + (size_t)n_slices * S) + 8LL * n_slices * S;
  const long long target = total * c / G;
  int lo = 0, hi = n_slices;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long long)__ldg(step_off + (size_t)mid * S) + 8LL * mid * S < target) lo = mid + 1; else hi = mid;
  }
  cta_slices[c] = (c == G) ? n_slices : lo;
}

// visits of warp w among n slices dealt round-robin to 32 warps
__host__ __device__ inline int slab2_cnt(int n, int w) { return n > w ? (n - w + SLAB_WARPS - 1) / SLAB_WARPS : 0; }

// consumption rank of every (slice, slab) cell: cell_gv[i * S + s] = its visit index; vis_blocks[gv] = its blocks
__global__ void slab2_rank_kernel(const int* __restrict__ steps, const int* __restrict__ cta_slices,
                                  const int* __restrict__ warp_vis, int n_slices, int S, int G, int PS,
                                  int* __restrict__ cell_gv, int* __restrict__ vis_blocks) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_slices * S) return;
  const int i = (int)(idx / S), s = (int)(idx % S);
  int lo = 0, hi = G;  // CTA c with cta_slices[c] <= i < cta_slices[c + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cta_slices + mid) <= i) lo = mid; else hi = mid;
  }
  const int c = lo, sl0 = __ldg(cta_slices + c), L = __ldg(cta_slices + c + 1) - sl0;
  const int j = i - sl0, p = j / PS, jj = j % PS, w = jj % SLAB_WARPS, v = jj / SLAB_WARPS;
  const int n_this = min(PS, L - p * PS);
  const int rank = p * S * slab2_cnt(PS, w) + s * slab2_cnt(n_this, w) + v;  // (all earlier passes are full)
  const int gv = __ldg(warp_vis + c * SLAB_WARPS + w) + rank;
  cell_gv[idx] = gv;
  vis_blocks[gv] = steps[idx] >> 1;
}

// one warp per (row, slab): the row's non-zeros of the slab, reordered by bank class, dealt to the row's SPLIT
// sub-streams; the rest of the visit's blocks padded.  Also the refresh after the values changed.
__global__ void __launch_bounds__(256)
slab2_fill_kernel(const int* __restrict__ col, const float* __restrict__ val, const int* __restrict__ slab_ptr,
                  const int* __restrict__ cell_gv, const int* __restrict__ vis_off, int M, int S, int rows_per_slab,
                  int HR, int G, int nclass, long long n_rows_padded, uint2* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_rows_padded * S) return;
  const long long r = w / S;
  const int s = (int)(w % S);
  const int i = (int)(r / HR), rl = (int)(r % HR), SPLIT = G / HR;
  int q0 = 0, q1 = 0;
  if (r < M) {
    q0 = __ldg(slab_ptr + r * (S + 1) + s);
    q1 = __ldg(slab_ptr + r * (S + 1) + s + 1);
  }
  const int gv = __ldg(cell_gv + (size_t)i * S + s);
  const int b0 = __ldg(vis_off + gv), n = 2 * (__ldg(vis_off + gv + 1) - b0);  // first block, steps
  auto slot = [&](int p) {  // p-th non-zero of the row in this slab -> position in the stream (uint2 units)
    const int j = p % SPLIT, t = p / SPLIT;
    return ((size_t)(b0 + (t >> 1)) * G + (size_t)(rl + HR * j)) * 2 + (size_t)(t & 1);
  };
  int running = 0;
  for (int k = 0; k < nclass; ++k) {
    const int cls = (rl + k) % nclass;
    for (int base = q0; base < q1; base += 32) {
      const int q = base + lane;
      const int c = (q < q1) ? __ldg(col + q) : -1;
      const bool ok = c >= 0 && (c % nclass) == cls;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) out[slot(running + __popc(m & lt))] = make_uint2((unsigned)c, __float_as_uint(__ldg(val + q)));
      running += __popc(m);
    }
  }
  // padding points at the row BEHIND the slab in its shared-memory buffer: the kernel keeps zeros there (a 0.0f value
  // times a row of B that holds Inf / NaN would not be zero)
  const unsigned pad_col = (unsigned)((s + 1) * rows_per_slab);
  for (int p = (q1 - q0) + lane; p < n * SPLIT; p += 32) out[slot(p)] = make_uint2(pad_col, 0u);
}

template <int LPR>
__global__ void __launch_bounds__(SLAB_WARPS * 32, 1)
spmm_slab2_kernel(const uint4* __restrict__ stream, const int* __restrict__ vis_off, const int* __restrict__ warp_vis,
                  const int* __restrict__ cta_slices, const float* __restrict__ B, int K, int M, float* __restrict__ C,
                  long long ldc, int S, int rows_per_slab, int HR, int pass_rows, const float* __restrict__ bias,
                  int relu) {
  constexpr int N = 4 * LPR;
  constexpr int G = 32 / LPR;
  constexpr int BLK = G * 16;                      // bytes of one block
  constexpr int CHB = slab2_chunk_bytes(N);        // bytes of one ring chunk
  constexpr int CB = CHB / BLK;                    // blocks per chunk (4 / 2 / 1)
  constexpr int RC = SLAB2_RING;
  extern __shared__ __align__(128) unsigned char slab_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(slab_smem);                       // [2] slab barriers
  uint64_t* rbar = reinterpret_cast<uint64_t*>(slab_smem + 64);                  // [32][RC] ring barriers
  unsigned char* ring = slab_smem + 64 + SLAB_WARPS * RC * 8;                    // [32][RC][512 B]
  float* s_acc = reinterpret_cast<float*>(ring + (size_t)SLAB_WARPS * RC * CHB);  // [pass_rows][N]
  const size_t slab_bytes = (size_t)(rows_per_slab + 1) * N * sizeof(float);  // + the zero row of the padding
  float* const slab_buf0 = s_acc + (size_t)pass_rows * N;
  float* const slab_buf1 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(slab_buf0) + slab_bytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane / LPR, l = lane % LPR;
  const int sl0 = __ldg(cta_slices + blockIdx.x), sl1 = __ldg(cta_slices + blockIdx.x + 1);
  if (tid == 0) {
    mbar_init(full + 0, 1);
    mbar_init(full + 1, 1);
  }
  if (lane == 0) {
    for (int k = 0; k < RC; ++k) mbar_init(rbar + warp * RC + k, 1);
    mbar_fence_init();
  }
  if (tid < N) slab_buf0[(size_t)rows_per_slab * N + tid] = 0.f;
  else if (tid < 2 * N) slab_buf1[(size_t)rows_per_slab * N + tid - N] = 0.f;
  __syncthreads();

  // this warp's stream: visits [gv, gv_end), blocks [wb0, wb0 + nblk) — one contiguous piece
  int gv = __ldg(warp_vis + blockIdx.x * SLAB_WARPS + warp);
  const int gv_end = __ldg(warp_vis + blockIdx.x * SLAB_WARPS + warp + 1);
  const int wb0 = __ldg(vis_off + gv), nblk = __ldg(vis_off + gv_end) - wb0;
  const int nchunks = (nblk + CB - 1) / CB;
  const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(stream) + (size_t)wb0 * BLK;
  unsigned char* wring = ring + (size_t)warp * RC * CHB;
  uint64_t* wbar = rbar + warp * RC;
  // (everything below works on 32-bit shared-window addresses: no generic-pointer arithmetic in the hot loop)
  const uint32_t wring_s = smem_u32(wring), wbar_s = smem_u32(wbar);
  uint64_t policy;  // the stream is read exactly once: evict-first in L2
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  auto issue_chunk = [&](int ci) {  // lane 0: chunk ci of this warp's stream -> ring slot ci % RC
    const int rem = nblk - ci * CB;
    const uint32_t bytes = (uint32_t)(rem >= CB ? CB : rem) * BLK;
    const uint32_t sl = (uint32_t)ci & (RC - 1);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wbar_s + sl * 8u), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            wring_s + sl * CHB),
        "l"(wsrc + (size_t)ci * CHB), "r"(bytes), "r"(wbar_s + sl * 8u), "l"(policy)
        : "memory");
  };
  auto wait_chunk = [&](uint32_t sl, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t"
        "}" ::"r"(wbar_s + sl * 8u),
        "r"(parity)
        : "memory");
  };
  if (lane == 0)
    for (int ci = 0; ci < RC && ci < nchunks; ++ci) issue_chunk(ci);
  // position in the stream: chunk ci (ring slot, phase of its barrier), cb blocks of it consumed, byte offset in the ring
  int ci = 0, slot = 0, cb = 0;
  uint32_t rphase = 0u, roff = 0u;
  const uint32_t ring_s = wring_s + 16u * g;
  static_assert((RC & (RC - 1)) == 0, "ring slots are addressed with a mask");

  uint32_t phase0 = 0u, phase1 = 0u;
  auto issue_slab = [&](int s, int b) {  // thread 0: rows [s * rps, min(K, (s+1) * rps)) of B -> buffer b
    const int k0 = s * rows_per_slab;
    const uint32_t bytes = (uint32_t)(min(rows_per_slab, K - k0)) * N * (uint32_t)sizeof(float);
    mbar_expect_tx(full + b, bytes);
    tma_bulk_g2s(b ? slab_buf1 : slab_buf0, B + (size_t)k0 * N, bytes, full + b);
  };
  const int pass_slices = pass_rows / HR;
  const int red_lo = LPR * HR;  // shuffle distances >= this combine the sub-streams of one row

  for (int p0 = sl0; p0 < sl1; p0 += pass_slices) {
    const int p1 = min(p0 + pass_slices, sl1);
    const int c0 = p0 * HR;
    const int nrows = min((p1 - p0) * HR, M - c0);
    for (int i = tid; i < (p1 - p0) * HR * N; i += blockDim.x) s_acc[i] = 0.f;
    if (tid == 0) issue_slab(0, 0);
    __syncthreads();
    const int nvis = slab2_cnt(p1 - p0, warp);  // this warp's slices of the pass: p0 + warp + 32 v
    for (int s = 0; s < S; ++s) {
      const int b = s & 1;
      if (tid == 0 && s + 1 < S) issue_slab(s + 1, b ^ 1);  // buffer b^1 was released by the barrier ending slab s-1
      // block counts of this warp's visits in this slab (lane v: visit v), fetched before waiting for the slab
      int my_nb = 0;
      if (lane < nvis) my_nb = __ldg(vis_off + gv + lane + 1) - __ldg(vis_off + gv + lane);
      gv += nvis;
      if (b) { mbar_wait(full + 1, phase1); phase1 ^= 1u; } else { mbar_wait(full + 0, phase0); phase0 ^= 1u; }
      // shared-window address of B row 0 of this slab's numbering, this lane's 16 bytes (wraps below the buffer: only
      // rows of the slab are ever addressed)
      const uint32_t slab_s = smem_u32(b ? slab_buf1 : slab_buf0) - (uint32_t)s * (uint32_t)rows_per_slab * (N * 4u) + 16u * l;
      for (int v = 0; v < nvis; ++v) {
        const int nb = __shfl_sync(FULL, my_nb, v);
        if (nb == 0) continue;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < nb; ++k) {
          if (cb == 0) wait_chunk((uint32_t)slot, rphase);  // first block of a chunk: has its copy landed?
          const uint4 cur = lds128_u(ring_s + roff);
          const float4 b0 = lds128_f(slab_s + cur.x * (N * 4u));
          const float4 b1 = lds128_f(slab_s + cur.z * (N * 4u));
          const float a0 = __uint_as_float(cur.y), a1 = __uint_as_float(cur.w);
          acc[0] = fmaf(a0, b0.x, acc[0]); acc[1] = fmaf(a0, b0.y, acc[1]);
          acc[2] = fmaf(a0, b0.z, acc[2]); acc[3] = fmaf(a0, b0.w, acc[3]);
          acc[0] = fmaf(a1, b1.x, acc[0]); acc[1] = fmaf(a1, b1.y, acc[1]);
          acc[2] = fmaf(a1, b1.z, acc[2]); acc[3] = fmaf(a1, b1.w, acc[3]);
          roff = (roff + BLK) & (RC * CHB - 1);
          if (++cb == CB) {  // the warp leaves the chunk: refill its slot with the chunk RC further on
            cb = 0;
            __syncwarp();
            if (lane == 0 && ci + RC < nchunks) issue_chunk(ci + RC);
            ++ci;
            if (++slot == RC) { slot = 0; rphase ^= 1u; }
          }
        }
        for (int o = 16; o >= red_lo; o >>= 1) {
#pragma unroll
          for (int x = 0; x < 4; ++x) acc[x] += __shfl_xor_sync(FULL, acc[x], o);
        }
        const int row_local = (warp + SLAB_WARPS * v) * HR + g;
        if (g < HR && row_local < nrows) {
          float4* dst = reinterpret_cast<float4*>(s_acc + (size_t)row_local * N + 4 * l);
          float4 t = *dst;
          t.x += acc[0]; t.y += acc[1]; t.z += acc[2]; t.w += acc[3];
          *dst = t;
        }
      }
      __syncthreads();  // slab s consumed by every warp: its buffer may be refilled
    }
    for (int i = tid; i < nrows * LPR; i += blockDim.x) {
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

static size_t slab2_fixed_smem(int N) { return 64 + SLAB_WARPS * SLAB2_RING * 8 + (size_t)SLAB_WARPS * SLAB2_RING * slab2_chunk_bytes(N); }
size_t slab2_smem_bytes(int rows_per_slab, int N, int pass_rows) {
  return slab2_fixed_smem(N) + (size_t)pass_rows * N * sizeof(float) + 2 * (size_t)(rows_per_slab + 1) * N * sizeof(float);
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
void slab2_geometry(int M, int K, int N, int sms, int* pass_rows, int* rows_per_slab, int* HR);

bool slab_wanted(int M, int K, long long nnz, int N, const float* B, long long ldb, const float* C, long long ldc) {
  if (!(N == 4 || N == 8 || N == 16) || ldb != N || ldc % 4 != 0 || M <= 0) return false;
  if (reinterpret_cast<uintptr_t>(B) % 16 != 0 || reinterpret_cast<uintptr_t>(C) % 16 != 0) return false;
  const bool v2 = options().slab_mode != 2;
  int rps = slab_rows_per_slab(N), pr = 0, hr = 0;
  if (v2) slab2_geometry(M, K, N, device_sm_count(), &pr, &rps, &hr);
  const long long S = ((long long)K + rps - 1) / rps;
  if (S > 16) return false;
  if (options().slab_mode == 3) return true;  // (probing) whenever the shape allows it
  if (nnz < (1LL << 20)) return false;
  // enough non-zeros per (row, slab) visit: below that the per-visit bookkeeping (block counts, group reduction,
  // running-sum update, padding of the slices) outweighs the cheaper gathers.  Measured on 16384^2
  // (profiles/r2_ncu_fwd_n16.md): sliced stream (v2) x16 @99 % (13 per visit) 0.029 vs 0.031 ms on the lane-group
  // kernel, x8 @99 % (20 per visit) 0.021 vs 0.031, @90 % (130-200 per visit) 0.105 / 0.070 vs 0.215 / 0.205;
  // first kernel (CSR walked in place): break-even near 48 per visit.
  return nnz >= (v2 ? 12LL : 48LL) * M * S;
}

void slab_store_free(SlabStore* st) {
  cudaFree(st->slab_ptr); cudaFree(st->cta_rows);
  cudaFree(st->vis_off); cudaFree(st->cell_gv); cudaFree(st->warp_vis); cudaFree(st->cta_slices); cudaFree(st->stream);
  *st = SlabStore{};
}

// v2 geometry: the row sums of a pass take what ~2x the average CTA share needs (at most 1024 rows), the two slab
// buffers share the rest of 220 KB
void slab2_geometry(int M, int K, int N, int sms, int* pass_rows, int* rows_per_slab, int* HR) {
  const int G = 32 / (N / 4);
  const int nclass = slab2_nclass(N);
  const int share = (M + sms - 1) / sms;
  int hr = 1;
  while (hr * 2 <= G && share / (hr * 2) >= 24) hr *= 2;
  hr = std::max(hr, std::min(nclass, G));
  int pr = std::min(SLAB_ROWS_MAX, std::max(64, ((2 * share + 31) / 32) * 32));
  pr = ((pr + hr - 1) / hr) * hr;
  const size_t avail = 220 * 1024 - slab2_fixed_smem(N) - (size_t)pr * N * sizeof(float) - 2 * (size_t)N * sizeof(float);
  int rps = (int)((avail / 2) / ((size_t)N * sizeof(float))) & ~7;
  rps = std::min(rps, ((K + 7) / 8) * 8);
  *pass_rows = pr; *rows_per_slab = rps; *HR = hr;
}

int slab_store_build(const int* row_ptr, const int* col, const float* val, int M, int K, int nnz, int N, SlabStore* st,
                     cudaStream_t s) {
  *st = SlabStore{};
  st->N = N; st->M = M; st->K = K;
  st->G = device_sm_count();
  st->v2 = options().slab_mode != 2;
  if (st->v2) slab2_geometry(M, K, N, st->G, &st->pass_rows, &st->rows_per_slab, &st->HR);
  else st->rows_per_slab = slab_rows_per_slab(N);
  st->S = (K + st->rows_per_slab - 1) / st->rows_per_slab;
  const size_t cells = (size_t)M * (st->S + 1);
  B200_CUDA_OK(cudaMalloc(&st->slab_ptr, cells * sizeof(int)));
  slab_ptr_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(row_ptr, col, M, st->S, st->rows_per_slab, st->slab_ptr);
  B200_LAUNCH_OK("slab_ptr_kernel");
  if (!st->v2) {
    B200_CUDA_OK(cudaMalloc(&st->cta_rows, ((size_t)st->G + 1) * sizeof(int)));
    slab_rows_kernel<<<(st->G + 256) / 256, 256, 0, s>>>(row_ptr, M, nnz, st->G, st->cta_rows);
    B200_LAUNCH_OK("slab_rows_kernel");
    return B200_OK;
  }
  // v2: steps per (slice, slab) -> slice-major prefix -> CTA ranges -> consumption order -> the stream itself
  const int Gl = 32 / (N / 4), SPLIT = Gl / st->HR;
  st->n_slices = (M + st->HR - 1) / st->HR;
  const size_t n_cells = (size_t)st->n_slices * st->S + 1;
  if (n_cells >= 0x7fffffffULL) { set_error("slab store: too many (slice, slab) cells"); return B200_ERR_INVALID; }
  int *steps = nullptr, *step_off = nullptr, *vis_blocks = nullptr;
  void* temp = nullptr;
  auto cleanup = [&]() { cudaFree(steps); cudaFree(step_off); cudaFree(vis_blocks); cudaFree(temp); };
#define SLAB_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      cleanup();                                                                            \
      set_error("slab store build: %s", cudaGetErrorString(e_));                            \
      return B200_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)
  SLAB_TRY(cudaMalloc(&steps, n_cells * sizeof(int)));
  SLAB_TRY(cudaMalloc(&step_off, n_cells * sizeof(int)));
  SLAB_TRY(cudaMalloc(&vis_blocks, n_cells * sizeof(int)));
  SLAB_TRY(cudaMalloc(&st->vis_off, n_cells * sizeof(int)));
  SLAB_TRY(cudaMalloc(&st->cell_gv, n_cells * sizeof(int)));
  SLAB_TRY(cudaMalloc(&st->cta_slices, ((size_t)st->G + 1) * sizeof(int)));
  SLAB_TRY(cudaMalloc(&st->warp_vis, ((size_t)st->G * SLAB_WARPS + 1) * sizeof(int)));
  slab2_steps_kernel<<<(unsigned)((n_cells + 255) / 256), 256, 0, s>>>(st->slab_ptr, M, st->S, st->HR, SPLIT, st->n_slices, steps);
  size_t temp_bytes = 0;
  SLAB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, steps, step_off, (int)n_cells, s));
  SLAB_TRY(cudaMalloc(&temp, std::max<size_t>(temp_bytes, 16)));
  SLAB_TRY(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, steps, step_off, (int)n_cells, s));
  slab2_cta_kernel<<<(st->G + 256) / 256, 256, 0, s>>>(step_off, st->n_slices, st->S, st->G, st->cta_slices);
  // first visit of every warp (host: G * 32 small sums from the CTA ranges)
  std::vector<int> h_cta((size_t)st->G + 1), h_wv((size_t)st->G * SLAB_WARPS + 1);
  SLAB_TRY(cudaMemcpyAsync(h_cta.data(), st->cta_slices, h_cta.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
  SLAB_TRY(cudaStreamSynchronize(s));
  const int PS = st->pass_rows / st->HR;
  long long run = 0;
  for (int c = 0; c < st->G; ++c) {
    const int L = h_cta[c + 1] - h_cta[c];
    for (int w = 0; w < SLAB_WARPS; ++w) {
      h_wv[(size_t)c * SLAB_WARPS + w] = (int)run;
      run += (long long)st->S * ((long long)(L / PS) * slab2_cnt(PS, w) + slab2_cnt(L % PS, w));
    }
  }
  h_wv[(size_t)st->G * SLAB_WARPS] = (int)run;
  SLAB_TRY(cudaMemcpyAsync(st->warp_vis, h_wv.data(), h_wv.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  SLAB_TRY(cudaMemsetAsync(vis_blocks, 0, n_cells * sizeof(int), s));
  slab2_rank_kernel<<<(unsigned)((n_cells + 255) / 256), 256, 0, s>>>(steps, st->cta_slices, st->warp_vis, st->n_slices, st->S,
                                                                    st->G, PS, st->cell_gv, vis_blocks);
  SLAB_TRY(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, vis_blocks, st->vis_off, (int)n_cells, s));
  int total_blocks = 0;
  SLAB_TRY(cudaMemcpyAsync(&total_blocks, st->vis_off + (n_cells - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
  SLAB_TRY(cudaStreamSynchronize(s));  // (also: h_wv has been copied)
  SLAB_TRY(cudaGetLastError());
  cleanup();
#undef SLAB_TRY
  st->total_steps = 2 * total_blocks;
  B200_CUDA_OK(cudaMalloc(&st->stream, std::max<size_t>((size_t)total_blocks * Gl * sizeof(uint4), 16)));
  return slab_store_refresh(col, val, st, s);
}

// (re)write the stream from the CSR arrays: at build time, and after the values changed in place
int slab_store_refresh(const int* col, const float* val, SlabStore* st, cudaStream_t s) {
  if (!st->v2 || st->stream == nullptr) return B200_OK;
  const int Gl = 32 / (st->N / 4);
  const long long rows_padded = (long long)st->n_slices * st->HR;
  const long long warps = rows_padded * st->S;
  slab2_fill_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(col, val, st->slab_ptr, st->cell_gv, st->vis_off, st->M, st->S,
                                                              st->rows_per_slab, st->HR, Gl, slab2_nclass(st->N),
                                                              rows_padded, reinterpret_cast<uint2*>(st->stream));
  B200_LAUNCH_OK("slab2_fill_kernel");
  return B200_OK;
}

int launch_spmm_slab(const SlabStore& st, const int* col, const float* val, const float* B, float* C, long long ldc,
                     const float* bias, int relu, cudaStream_t s) {
  const size_t smem = st.v2 ? slab2_smem_bytes(st.rows_per_slab, st.N, st.pass_rows) : slab_smem_bytes(st.rows_per_slab, st.N);
#define B200_SLAB(L)                                                                                             \
  do {                                                                                                           \
    static bool attr[64] = {false};                                                                              \
    int dev = 0;                                                                                                 \
    B200_CUDA_OK(cudaGetDevice(&dev));                                                                           \
    if (!attr[dev & 63]) {                                                                                       \
      B200_CUDA_OK(cudaFuncSetAttribute(spmm_slab_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      B200_CUDA_OK(cudaFuncSetAttribute(spmm_slab2_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr[dev & 63] = true;                                                                                     \
    }                                                                                                            \
    if (st.v2)                                                                                                   \
      spmm_slab2_kernel<L><<<st.G, SLAB_WARPS * 32, smem, s>>>(st.stream, st.vis_off, st.warp_vis, st.cta_slices, B, st.K, st.M, C, \
                                                             ldc, st.S, st.rows_per_slab, st.HR, st.pass_rows, bias, relu); \
    else                                                                                                         \
      spmm_slab_kernel<L><<<st.G, SLAB_WARPS * 32, smem, s>>>(col, val, st.slab_ptr, st.cta_rows, B, st.K, C, ldc, st.S, \
                                                            st.rows_per_slab, bias, relu);                       \
  } while (0)
  if (st.N == 16) B200_SLAB(4);
  else if (st.N == 8) B200_SLAB(2);
  else B200_SLAB(1);
#undef B200_SLAB
  B200_LAUNCH_OK("spmm_slab_kernel");
  return B200_OK;
}

}  // namespace b200
