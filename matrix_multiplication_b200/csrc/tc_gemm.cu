// tc_gemm.cu — the block-dense tensor-core path of the SpMM engine (sm_100a: tcgen05 + TMEM + TMA bulk copies).
//
// When a row block of A is dense enough to be a real contraction (BASELINE north_star: "block-dense tiles"),
// gathering one B row per non-zero is the wrong machine: every loaded B element feeds exactly one FMA and the
// L1/L2 gather path saturates at ~10 TFLOP/s.  Here the plan keeps the non-empty 128 x 32 tiles of A densified,
// and C = A * B runs as a tiled GEMM on the 5th-generation tensor cores with fp32-grade accuracy:
//
//   * split precision ("3xTF32"): x = hi + lo with hi = tf32(x), lo = tf32(x - hi);  A*B ~= Ahi*Bhi + Ahi*Blo +
//     Alo*Bhi, three tcgen05.mma kind::tf32 per k-step, fp32 accumulation in TMEM.  The dropped term and the
//     rounding of lo are <= 2^-21 relative per product (the parity tests check <= 1e-5 against the fp64 oracle).
//   * operand layout: both operands K-major in 128-byte-swizzled atoms of 8 rows x 32 tf32 (1 KB), the hi atom
//     and the lo atom of the same 8 rows adjacent (2 KB).  A tiles are packed once per plan (tc_store_build);
//     B is packed per call by tc_pack_b_kernel (one read of B, hi + lo written).  A whole tile operand is then
//     ONE contiguous chunk of memory: the producer warp moves it with cp.async.bulk (TMA, SASS UBLKCP)
//     signalling an mbarrier, no tensor map needed.
//   * kernel: persistent CTAs, static tile schedule, warp-specialised — warp 0 TMA producer, warp 1 (one lane)
//     tcgen05.mma issuer + TMEM owner, warps 2-5 epilogue (tcgen05.ld -> bias/ReLU -> global).  smem ring of
//     2-3 stages, TMEM accumulator double-buffered so the epilogue of tile i overlaps the main loop of tile i+1.
//   * empty 128 x 32 tiles of A are not stored and not multiplied (blk_ptr / blk_kt lists per row block).
//
// Reference being replaced: cusparse_mm_wrapper (src/baseline_mm.cu:167-216) for dense-ish operands, and the
// dense comparison path it is measured against (cublas_mm_wrapper, src/baseline_mm.cu:52-102).
#include <algorithm>
#include <cmath>

#include <cuda_fp16.h>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace b200 {

namespace {

constexpr int BM = 128;           // rows of A per tile = UMMA M
constexpr int ATOM = 2048;        // bytes of one 8-row group: hi atom (1 KB) then lo atom (1 KB)
constexpr int A_TILE = (BM / 8) * ATOM;  // 32 KB
constexpr int EPI_WARPS = 8;      // two warps per TMEM lane quarter, each owns half of the tile's columns
constexpr int TC_THREADS = (2 + EPI_WARPS) * 32;
constexpr int MAX_STAGES = 4;
// Element kinds.  Both use 128-byte swizzled K-major rows, so a stage covers 128 / sizeof(element) values of k
// and one tcgen05.mma consumes 32 bytes of every row (K = 8 tf32 or 16 f16).
constexpr int KIND_TF32 = 0;      // x = hi + lo, both tf32 (exponent range of fp32: no scaling needed)
constexpr int KIND_F16 = 1;       // x * 2^s = hi + lo, both fp16, s per row of A / per column of B (twice the MMA rate,
                                  // half the bytes)
__host__ __device__ constexpr int bk_of(int kind) { return kind == KIND_F16 ? 64 : 32; }

// byte offset of element (r, c) inside a 1 KB K-major SWIZZLE_128B atom: r < 8 rows of 128 B, c-th element of
// `esz` bytes in the row.  The 16-byte chunk index is XORed with the row index (what TMA / UMMA call SWIZZLE_128B).
__host__ __device__ __forceinline__ int sw128(int r, int c, int esz) {
  const int b = c * esz;
  return r * 128 + ((((b >> 4) ^ r) & 7) << 4) + (b & 15);
}

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u & 0xffffe000u);
}
// fp16 scaling: values are multiplied by 2^(141 - e) where e = max(biased exponent of the row / column maximum, 15),
// so the largest magnitude lands in [2^14, 2^15) (fp16 overflows at 65504) and results are scaled back by 2^(e - 141).
__device__ __forceinline__ int scale_exp(uint32_t max_abs_bits) { return max((int)(max_abs_bits >> 23) & 0xff, 15); }
__device__ __forceinline__ float scale_up(int e) { return __uint_as_float((uint32_t)(268 - e) << 23); }
__device__ __forceinline__ float scale_down(int e) { return __uint_as_float((uint32_t)(e - 14) << 23); }

template <int KIND>
struct Split;
template <>
struct Split<KIND_TF32> {
  using T = float;
  __device__ static __forceinline__ void split(float x, float, T& hi, T& lo) {
    hi = tf32_rn(x);
    lo = tf32_rn(x - hi);
  }
};
template <>
struct Split<KIND_F16> {
  using T = __half;
  __device__ static __forceinline__ void split(float x, float up, T& hi, T& lo) {
    const float xs = x * up;
    hi = __float2half_rn(xs);
    lo = __float2half_rn(xs - __half2float(hi));
  }
};

// ---- plan-time packing of A --------------------------------------------------------------------------------
__global__ void tc_mark_tiles_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int M, int KT,
                                     int bk, int* __restrict__ flags) {
  // one warp per row: mark the k-tiles this row touches
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int mb = row / BM;
  const int rs = __ldg(row_ptr + row), re = __ldg(row_ptr + row + 1);
  for (int p = rs + lane; p < re; p += 32) flags[(size_t)mb * KT + (__ldg(col + p) / bk)] = 1;
}

__global__ void tc_list_tiles_kernel(const int* __restrict__ flags, const int* __restrict__ scan, int MB, int KT,
                                     int* __restrict__ blk_ptr, int* __restrict__ blk_kt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)MB * KT;
  if (i < n) {
    if (flags[i]) blk_kt[scan[i]] = (int)(i % KT);
    if (i % KT == 0) blk_ptr[i / KT] = scan[i];
  }
  if (i == n - 1) blk_ptr[MB] = scan[i] + flags[i];
}

// CTA-pair lists: per pair of 128-row blocks the union of their stored k-tiles and each block's tile slot
__global__ void tc_pair_flags_kernel(const int* __restrict__ flags, int MB, int MB2, int KT, int* __restrict__ pflags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)MB2 * KT) return;
  const int mp = (int)(i / KT), kt = (int)(i % KT);
  int f = flags[(size_t)(2 * mp) * KT + kt];
  if (2 * mp + 1 < MB) f |= flags[(size_t)(2 * mp + 1) * KT + kt];
  pflags[i] = f;
}
__global__ void tc_pair_list_kernel(const int* __restrict__ flags, const int* __restrict__ scan,
                                    const int* __restrict__ pflags, const int* __restrict__ pscan, int MB, int MB2, int KT,
                                    int zero_slot, int n_pair_tiles, int* __restrict__ pair_ptr, int* __restrict__ pair_kt,
                                    int* __restrict__ pair_slot) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)MB2 * KT;
  if (i >= n) return;
  const int mp = (int)(i / KT), kt = (int)(i % KT);
  if (kt == 0) pair_ptr[mp] = pscan[i];
  if (i == n - 1) pair_ptr[MB2] = pscan[i] + pflags[i];
  if (pflags[i]) {
    const int idx = pscan[i];
    pair_kt[idx] = kt;
    const size_t c0 = (size_t)(2 * mp) * KT + kt, c1 = (size_t)(2 * mp + 1) * KT + kt;
    pair_slot[idx] = flags[c0] ? scan[c0] : zero_slot;
    pair_slot[(size_t)n_pair_tiles + idx] = (2 * mp + 1 < MB && flags[c1]) ? scan[c1] : zero_slot;
  }
}

template <int KIND>
__global__ void tc_pack_a_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col,
                                 const float* __restrict__ val, int M, int KT, const int* __restrict__ scan,
                                 unsigned char* __restrict__ tiles, float* __restrict__ row_scale) {
  using T = typename Split<KIND>::T;
  constexpr int BK = bk_of(KIND);
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int mb = row / BM, r = row % BM;
  const int rs = __ldg(row_ptr + row), re = __ldg(row_ptr + row + 1);
  float up = 1.f;
  if constexpr (KIND == KIND_F16) {
    uint32_t mx = 0;
    for (int p = rs + lane; p < re; p += 32) {
      const uint32_t b = __float_as_uint(__ldg(val + p)) & 0x7fffffffu;
      mx = max(mx, b < 0x7f800000u ? b : 0u);  // Inf/NaN in A stay Inf/NaN after the split whatever the scale
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int e = scale_exp(mx);
    up = scale_up(e);
    if (lane == 0) row_scale[row] = scale_down(e);
  }
  for (int p = rs + lane; p < re; p += 32) {
    const int c = __ldg(col + p);
    const int slot = scan[(size_t)mb * KT + c / BK];
    unsigned char* atom = tiles + (size_t)slot * A_TILE + (size_t)(r >> 3) * ATOM;
    const int off = sw128(r & 7, c % BK, (int)sizeof(T));
    T hi, lo;
    Split<KIND>::split(__ldg(val + p), up, hi, lo);
    *reinterpret_cast<T*>(atom + off) = hi;
    *reinterpret_cast<T*>(atom + 1024 + off) = lo;
  }
}

// ---- per-call packing of the dense operand ----------------------------------------------------------------
// fp16 kind, pass 1: per-column maximum magnitude (as uint bits: monotonic for non-negative floats) and the
// non-finite flag.  TRANS = false: B [K, N] row-major, thread = column, blockIdx.y strides over k.
template <bool TRANS>
__global__ void __launch_bounds__(256)
tc_colmax_kernel(const float* __restrict__ B, long long ld, int K, int N, uint32_t* __restrict__ colmax,
                 int* __restrict__ nonfinite) {
  if constexpr (!TRANS) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    uint32_t mx = 0;
    for (int k = blockIdx.y; k < K; k += 8 * gridDim.y) {  // eight independent loads in flight per thread
      uint32_t v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kk = k + i * (int)gridDim.y;
        v[i] = (kk < K) ? __float_as_uint(__ldg(B + (long long)kk * ld + n)) & 0x7fffffffu : 0u;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) mx = max(mx, v[i]);
    }
    if (mx >= 0x7f800000u) { atomicOr(nonfinite, 1); mx = 0; }
    atomicMax(colmax + n, mx);
  } else {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    uint32_t mx = 0;
    for (int k = lane; k < K; k += 32) mx = max(mx, __float_as_uint(__ldg(B + (long long)n * ld + k)) & 0x7fffffffu);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
      if (mx >= 0x7f800000u) { atomicOr(nonfinite, 1); mx = 0; }
      colmax[n] = mx;
    }
  }
}

// Packed layout: [kt][n group of 8][hi atom | lo atom]; atom row = one n, its BK k values (128 B, swizzled).
// TRANS = false: B is [K, N] row-major (the reference's layout, src/custom_mm.cpp:203-217): a warp reads, for
//   each k row of the tile, the 32-byte sector holding 8 consecutive n and writes full 128-byte atom rows
//   (the 32 lanes of a store cover one row).
// TRANS = true: the operand is given as B^T [N, K] row-major (activations [tokens, features] of an FC layer):
//   atom rows are contiguous in memory, 16/32-byte vectors in and out.
// Rows k >= K and columns n >= N are written as zeros.  Non-finite inputs raise *nonfinite (the dense
// contraction would turn 0 * Inf into NaN where the sparse product skips the term): the caller falls back.
// cm8: the maximum magnitudes (uint bits) of the group's 8 columns (fp16 kind), global or shared memory.
template <int KIND, bool TRANS>
__device__ __forceinline__ bool tc_pack_atom(const float* __restrict__ B, long long ld, int K, int N, int kt, int g,
                                             unsigned char* __restrict__ atom, const uint32_t* cm8, int lane) {
  using T = typename Split<KIND>::T;
  constexpr int BK = bk_of(KIND);
  constexpr int KL = BK / 32;  // k values per lane (non-transposed read)
  bool bad = false;
  const int n0 = g * 8, k0 = kt * BK;
  if constexpr (!TRANS) {
    // lane owns k = k0 + KL*lane .. +KL-1 (adjacent k: one 4-byte store per atom row for fp16)
    float x[KL][8];
#pragma unroll
    for (int i = 0; i < KL; ++i) {
      const int k = k0 + KL * lane + i;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[i][j] = 0.f;
      if (k < K) {
        const float* src = B + (long long)k * ld + n0;
        if (n0 + 8 <= N && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src));
          const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
          x[i][0] = a.x; x[i][1] = a.y; x[i][2] = a.z; x[i][3] = a.w;
          x[i][4] = b.x; x[i][5] = b.y; x[i][6] = b.z; x[i][7] = b.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + j < N) x[i][j] = __ldg(src + j);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float up = 1.f;
      if constexpr (KIND == KIND_F16) up = scale_up(scale_exp(cm8[j]));
      T hi[KL], lo[KL];
#pragma unroll
      for (int i = 0; i < KL; ++i) {
        bad |= !(fabsf(x[i][j]) <= 3.0e38f);
        Split<KIND>::split(x[i][j], up, hi[i], lo[i]);
      }
      const int off = sw128(j, KL * lane, (int)sizeof(T));
      if constexpr (KIND == KIND_F16) {
        *reinterpret_cast<__half2*>(atom + off) = __halves2half2(hi[0], hi[1]);
        *reinterpret_cast<__half2*>(atom + 1024 + off) = __halves2half2(lo[0], lo[1]);
      } else {
        *reinterpret_cast<T*>(atom + off) = hi[0];
        *reinterpret_cast<T*>(atom + 1024 + off) = lo[0];
      }
    }
  } else {
    // lane -> (row rr = lane / 4, output 16-byte chunks q = lane % 4 and q + 4); a chunk holds EC = 16 / sizeof(T)
    // consecutive k of row n
    constexpr int EC = 16 / (int)sizeof(T);
    const int rr = lane >> 2, n = n0 + rr;
    float up = 1.f;
    if constexpr (KIND == KIND_F16) up = scale_up(scale_exp(cm8[rr]));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = (lane & 3) + 4 * h;
      const int k = k0 + EC * q;
      float x[EC];
#pragma unroll
      for (int j = 0; j < EC; ++j) x[j] = 0.f;
      if (n < N) {
        const float* src = B + (long long)n * ld + k;
        if (k + EC <= K && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
          for (int v = 0; v < EC / 4; ++v) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src) + v);
            x[4 * v] = a.x; x[4 * v + 1] = a.y; x[4 * v + 2] = a.z; x[4 * v + 3] = a.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < EC; ++j)
            if (k + j < K) x[j] = __ldg(src + j);
        }
      }
      T hi[EC], lo[EC];
#pragma unroll
      for (int j = 0; j < EC; ++j) {
        bad |= !(fabsf(x[j]) <= 3.0e38f);
        Split<KIND>::split(x[j], up, hi[j], lo[j]);
      }
      const int off = rr * 128 + (((q ^ rr) & 7) << 4);
      *reinterpret_cast<uint4*>(atom + off) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(atom + 1024 + off) = *reinterpret_cast<const uint4*>(lo);
    }
  }
  return bad;
}

// one warp per atom pair, grid-stride; column maxima (fp16 kind) come from tc_colmax_kernel
template <int KIND, bool TRANS>
__global__ void __launch_bounds__(256)
tc_pack_b_kernel(const float* __restrict__ B, long long ld, int K, int N, int KT, int NG,
                 unsigned char* __restrict__ out, const uint32_t* __restrict__ colmax, int* __restrict__ nonfinite) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = KT * NG;
  bool bad = false;
  for (int w = warp; w < total; w += (gridDim.x * blockDim.x) >> 5) {
    const int kt = w / NG, g = w % NG;
    bad |= tc_pack_atom<KIND, TRANS>(B, ld, K, N, kt, g, out + ((size_t)kt * NG + g) * ATOM, colmax + g * 8, lane);
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(nonfinite, 1);
}

// fp16 kind, wide operands (at least one CTA per SM): ONE kernel per call.  A CTA owns one group of 8 columns: it
// first scans the group for the column maxima (kept in shared memory and written out for the epilogue), then packs
// the group's atoms from the now cache-resident data.
template <bool TRANS>
__global__ void __launch_bounds__(256)
tc_prepack_b_kernel(const float* __restrict__ B, long long ld, int K, int N, int KT, int NG,
                    unsigned char* __restrict__ out, uint32_t* __restrict__ colmax, int* __restrict__ nonfinite) {
  __shared__ uint32_t s_max[8];
  const int g = blockIdx.x, n0 = g * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 8) s_max[threadIdx.x] = 0u;
  __syncthreads();
  bool bad = false;
  if constexpr (!TRANS) {
    uint32_t mx[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    const bool vec = n0 + 8 <= N && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(B + n0) & 15) == 0);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const float* src = B + (long long)k * ld + n0;
      float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (vec) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n0 + j < N) x[j] = __ldg(src + j);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) mx[j] = max(mx[j], __float_as_uint(x[j]) & 0x7fffffffu);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = 16; o; o >>= 1) mx[j] = max(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], o));
      if (lane == 0) atomicMax(&s_max[j], mx[j]);
    }
  } else {
    const int n = n0 + warp;  // 8 warps, 8 rows of B^T
    uint32_t mx = 0u;
    if (n < N)
      for (int k = lane; k < K; k += 32) mx = max(mx, __float_as_uint(__ldg(B + (long long)n * ld + k)) & 0x7fffffffu);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_max[warp] = mx;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    uint32_t m = s_max[threadIdx.x];
    if (m >= 0x7f800000u) { bad = true; m = 0u; }
    s_max[threadIdx.x] = m;
    colmax[n0 + threadIdx.x] = m;
  }
  __syncthreads();
  for (int kt = warp; kt < KT; kt += 8)
    bad |= tc_pack_atom<KIND_F16, TRANS>(B, ld, K, N, kt, g, out + ((size_t)kt * NG + g) * ATOM, s_max, lane);
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(nonfinite, 1);
}

// ---- tcgen05 / TMEM / mbarrier PTX ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must not hang the GPU box.  After ~2 s the waiter records where it gave up
// and the kernel drains (every later wait falls through at once); the host reports the failure.
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int* dbg, int code) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return true;
    if ((it & 1023u) == 1023u) {
      if (*reinterpret_cast<volatile int*>(dbg) != 0) return false;  // somebody already gave up
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {
        atomicCAS(dbg, 0, code);
        return false;
      }
    }
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both K-major, one elected thread issues.  SASS: UTCHMMA.
template <int KIND>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                     uint32_t accumulate) {
  if constexpr (KIND == KIND_F16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed (implies fence::before)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive columns of 32-bit TMEM -> 32 registers per thread (thread = lane = row).  SASS: LDTM
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups `sbo` bytes apart (sm_100 version bit)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;                        // leading byte offset: unused for swizzled K-major (canonical 1)
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;   // stride byte offset
  d |= (uint64_t)1 << 46;                        // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A/B tf32 (format 2) or f16 (format 0), both K-major, M = 128, N = nt
__host__ __device__ __forceinline__ uint32_t umma_idesc(int kind, int nt) {
  const uint32_t fmt = kind == KIND_F16 ? 0u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// One k-tile's twelve MMAs (4 k-steps x {hi*lo, lo*hi, hi*hi}) with the descriptor arithmetic taken out of the loop.
// Only the address field of a shared-memory descriptor (bits 0-13, 16-byte units) differs between the operands of a
// stage, so a descriptor is (lo + small constant, hi): one integer add per operand instead of rebuilding 64 bits.
// (The timeline of tools/tc_trace.py showed the issuing lane spending 0.67 us per k-tile in the old loop — shifts, ORs,
// three run-time `terms` branches and a predicate set-up per MMA, ~200 dependent scalar instructions — and a "no MMA,
// no copies" run of config 5 x 64 still taking 180 of 225 us: the issue loop, not HBM or the tensor pipe, set the pace.)
// One lane of a converged warp (elect.sync): the compiler then issues the tcgen05 instructions under one predicate
// instead of wrapping each of them in a loop over the active lanes (ELECT / PLOP3 / BRA.U.ANY per MMA), which is what
// it emits inside `if (lane == 0)`.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
#define B200_UMMA_LO_ASM(CGS, KINDS, PRED_SRC)                                                        \
  "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"                                                      \
  "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"                                               \
  "setp.ne.b32 p, " PRED_SRC ", 0;\n\t"                                                              \
  "tcgen05.mma.cta_group::" CGS ".kind::" KINDS " [%0], da, db, %4, p;\n\t}"
// CG = cta_group (1: one CTA, M = 128; 2: a CTA pair, M = 256); ACC1: accumulate flag known to be set
template <int KIND, bool ACC1, int CG>
__device__ __forceinline__ void umma_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                        uint32_t accumulate) {
#define B200_UMMA_LO_CASE(CGS, KINDS)                                                                               \
  if constexpr (ACC1)                                                                                               \
    asm volatile(B200_UMMA_LO_ASM(CGS, KINDS, "1")::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)      \
                 : "memory");                                                                                       \
  else                                                                                                              \
    asm volatile(B200_UMMA_LO_ASM(CGS, KINDS, "%5")::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc),    \
                 "r"(accumulate)                                                                                    \
                 : "memory");
  if constexpr (CG == 1 && KIND == KIND_F16) { B200_UMMA_LO_CASE("1", "f16") }
  if constexpr (CG == 1 && KIND != KIND_F16) { B200_UMMA_LO_CASE("1", "tf32") }
  if constexpr (CG == 2 && KIND == KIND_F16) { B200_UMMA_LO_CASE("2", "f16") }
  if constexpr (CG == 2 && KIND != KIND_F16) { B200_UMMA_LO_CASE("2", "tf32") }
#undef B200_UMMA_LO_CASE
}
// sa = shared-window address of the stage's M operand (hi atom of its first 8-row group), the N operand A_TILE behind it;
// hi and lo atoms 1 KB apart (64 units), a k-step 32 bytes (2 units) along the swizzled row.  Same order of the three
// products as the general loop (hi*lo, lo*hi, hi*hi), so results are bit-identical to it.
template <int KIND, int CG = 1>
__device__ __forceinline__ void umma_ktile_all_terms(uint32_t d_tmem, uint32_t sa, uint32_t idesc, uint32_t accumulate) {
  const uint64_t d0 = umma_desc(sa, ATOM);
  const uint32_t hi = (uint32_t)(d0 >> 32);
  const uint32_t a = (uint32_t)d0, b = a + (uint32_t)(A_TILE >> 4);
  umma_lo<KIND, false, CG>(d_tmem, a, b + 64u, hi, idesc, accumulate);
  umma_lo<KIND, true, CG>(d_tmem, a + 64u, b, hi, idesc, 1u);
  umma_lo<KIND, true, CG>(d_tmem, a, b, hi, idesc, 1u);
#pragma unroll
  for (uint32_t k = 1; k < 4; ++k) {
    umma_lo<KIND, true, CG>(d_tmem, a + 2u * k, b + 64u + 2u * k, hi, idesc, 1u);
    umma_lo<KIND, true, CG>(d_tmem, a + 64u + 2u * k, b + 2u * k, hi, idesc, 1u);
    umma_lo<KIND, true, CG>(d_tmem, a + 2u * k, b + 2u * k, hi, idesc, 1u);
  }
}

struct TcParams {
  const unsigned char* a_tiles;
  const int* blk_ptr;
  const int* blk_kt;
  const unsigned char* b_pack;
  int NG;          // n groups of 8 in the packed operand (N padded to a multiple of Nt)
  float* C;
  long long ldc;
  int M, N;
  const float* bias;
  const float* row_scale;      // fp16 kind: 2^(e - 141) per row of A / from the column maxima of B; else null
  const uint32_t* colmax;
  unsigned flags;
  int MB, NB, Nt;
  int stages;
  int chunk;       // k-tiles accumulated in TMEM before the epilogue warps fold them into their fp32 registers
  int terms;       // bit 0 hi*hi, bit 1 hi*lo, bit 2 lo*hi (debug / accuracy studies; 7 = all)
  int* ctl;        // [0] the packer's non-finite flag: when set this launch computes C in CSR order instead (0 * Inf
                   //     must stay out of the result); [1] first pipeline wait that timed out (debug); [2] exit ticket
  // the CSR and the raw dense operand, for that CSR-order pass
  const int* row_ptr;
  const int* col;
  const float* val;
  const float* B;
  long long b_sk, b_sn;
  // SDDMM mode (grad of the CSR values): A operand = packed G (a_tiles, NGa groups of 8 rows per k-tile), B operand
  // = packed B, contraction over the KT k-tiles of the dense column index; the epilogue samples the tile at A's
  // non-zeros: per (32-column group, row) a bit mask of the pattern and the position of the group's first non-zero
  int NGa, KT;
  const uint32_t* amax;       // fp16 kind: row maxima of G (uint bits)
  const uint32_t* samp_mask;  // [groups][M]
  const int* samp_off;        // [groups][M]
  int groups;
  const float* G;             // raw G for the CSR-order pass
  long long g_sm, g_sn;
  // CTA-pair kernel (cta_group::2): 128-row blocks 2 mp and 2 mp + 1 share every B tile; per pair the union of their
  // stored k-tiles, and for each of the two blocks the tile slot (the all-zero tile when that block has none there)
  const int* pair_ptr;        // [MB2 + 1]
  const int* pair_kt;         // [n_pair_tiles]
  const int* pair_slot;       // [2][n_pair_tiles]
  int n_pair_tiles, MB2;
  int pair_ptr_ok;            // the pair lists (or the padded SDDMM operand) are in place
  int dense_pattern;          // SDDMM: the pattern keeps >= 30 % of the cells (tile-shape choice of the short kernel)
};

// exit protocol: the last CTA to leave re-arms the non-finite flag and the ticket for the next launch
__device__ __forceinline__ void tc_exit_ticket(int* ctl) {
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ctl + 2, 1) == (int)gridDim.x - 1) {
      ctl[0] = 0;
      ctl[2] = 0;
    }
  }
}

// C = A * B in CSR order with any strides (+bias, ReLU): warp per (row, 32-column slab)
__device__ __forceinline__ void csr_order_spmm(const int* __restrict__ row_ptr, const int* __restrict__ col,
                                               const float* __restrict__ val, int M, int N, const float* __restrict__ B,
                                               long long b_sk, long long b_sn, float* __restrict__ C, long long c_sm,
                                               long long c_sn, const float* __restrict__ bias, bool relu) {
  const int lane = threadIdx.x & 31;
  const int slabs = (N + 31) / 32;
  const long long total = (long long)M * slabs;
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total;
       w += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int row = (int)(w / slabs), n = (int)(w % slabs) * 32 + lane;
    const int rs = __ldg(row_ptr + row), re = __ldg(row_ptr + row + 1);
    if (n < N) {
      float acc = 0.f;
      for (int p = rs; p < re; ++p)
        acc = fmaf(__ldg(val + p), __ldg(B + (long long)__ldg(col + p) * b_sk + (long long)n * b_sn), acc);
      if (bias != nullptr) acc += __ldg(bias + row);
      if (relu) acc = fmaxf(acc, 0.f);
      C[(long long)row * c_sm + (long long)n * c_sn] = acc;
    }
  }
}

// out[p] = sum_n G[row(p), n] * B[col(p), n] in CSR order with any strides: warp per row, lanes over n
__device__ __forceinline__ void csr_order_sddmm(const int* __restrict__ row_ptr, const int* __restrict__ col, int M, int N,
                                                const float* __restrict__ G, long long g_sm, long long g_sn,
                                                const float* __restrict__ B, long long b_sk, long long b_sn,
                                                float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M;
       row += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int rs = __ldg(row_ptr + row), re = __ldg(row_ptr + row + 1);
    for (int p = rs; p < re; ++p) {
      const long long c = __ldg(col + p);
      float acc = 0.f;
      for (int n = lane; n < N; n += 32)
        acc = fmaf(__ldg(G + row * g_sm + (long long)n * g_sn), __ldg(B + c * b_sk + (long long)n * b_sn), acc);
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) out[p] = acc;
    }
  }
}

// Final step of a tile for one epilogue warp: acc[] holds this thread's row (m) x 128-column share of the tile.
// SpMM: scale, + bias, ReLU, store C (or C^T); SDDMM: write the entries that lie on A's pattern.
// A thread owns a ROW of the tile (TMEM lane = row), so storing from acc[] directly makes every store instruction
// touch 32 different rows (32 sectors).  Row-major C and the SDDMM's value array are therefore written through a
// per-warp 32 x 32 transposer in shared memory (`stg`, row stride 33 words: conflict-free both ways): afterwards a
// lane owns a COLUMN and one instruction writes 32 consecutive words of one row.  C^T needs no transposer (a register
// is one n, the 32 lanes are 32 consecutive m).
constexpr int STG_LD = 33;
constexpr int STG_WARP_BYTES = 32 * STG_LD * 4;
constexpr int STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
constexpr int TC_SMEM_MAX = 227 * 1024;

// Debug timeline (tc_terms bit 9): CTA 0's producer lane, MMA lane and first epilogue warp record (code, globaltimer)
// pairs into three regions of a device buffer; read back through b200_set_option("tc_trace_sync") / get_option.
constexpr int TRACE_REGION = 2048;  // words per role
__device__ unsigned long long g_tc_trace[3 * TRACE_REGION];
__device__ __forceinline__ void tc_trace(bool on, int role, int& n, int code) {
  if (on && n + 2 <= TRACE_REGION) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_tc_trace[role * TRACE_REGION + n] = (unsigned long long)code;
    g_tc_trace[role * TRACE_REGION + n + 1] = t;
    n += 2;
  }
}

// What the final step of a tile needs from global memory, fetched in one batch (one round trip, not one per group).
// Lane = row m for rs / bias / msk / bs; lane = column (of the c-th 32-column group) for cs.
struct TilePre {
  float rs, bias;
  float cs[4];
  uint32_t msk[4];  // SDDMM: pattern bits of (row m, group c)
  int bs[4];        //        position of the group's first non-zero in the CSR value array
};

template <int KIND, bool SDDMM>
__device__ __forceinline__ void tc_tile_prefetch(const TcParams& p, int mb, int nb, int q, int half, int ntw, int lane,
                                                 TilePre& t) {
  const int m = mb * BM + q * 32 + lane;
  t.rs = 1.f;
  t.bias = 0.f;
  if constexpr (SDDMM) {
    if constexpr (KIND == KIND_F16) t.rs = (m < p.M) ? scale_down(scale_exp(__ldg(p.amax + m))) : 0.f;
  } else {
    if constexpr (KIND == KIND_F16) t.rs = (m < p.M) ? __ldg(p.row_scale + m) : 0.f;
    if (p.bias != nullptr && m < p.M) t.bias = __ldg(p.bias + m);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int k0 = nb * p.Nt + half * ntw + c * 32;  // first column of the group (a column of A's pattern for the SDDMM)
    t.cs[c] = 1.f;
    t.msk[c] = 0u;
    t.bs[c] = 0;
    if (c * 32 < ntw) {
      if constexpr (SDDMM) {
        const int grp = k0 >> 5;
        if (grp < p.groups) {
          if constexpr (KIND == KIND_F16) t.cs[c] = scale_down(scale_exp(__ldg(p.colmax + k0 + lane)));
          if (m < p.M) {
            t.msk[c] = __ldg(p.samp_mask + (size_t)grp * p.M + m);
            t.bs[c] = __ldg(p.samp_off + (size_t)grp * p.M + m);
          }
        }
      } else {
        if constexpr (KIND == KIND_F16) t.cs[c] = scale_down(scale_exp((k0 + lane < p.N) ? __ldg(p.colmax + k0 + lane) : 0u));
      }
    }
  }
}

// SDDMM, one 32-row x 32-column group of one warp: v[j] = this lane's ROW, column j  ->  the entries on A's pattern.
// Through the warp's transposer: afterwards a lane owns a column and row r's entries are written by one instruction
// (consecutive words).  No branches in the loop: the unrolled iterations' loads and shuffles issue back to back.
template <int UNROLL>
__device__ __forceinline__ void tc_sddmm_group_store(float* __restrict__ out, const float (&v)[32], float rs, float cs,
                                                     uint32_t mask, int base, int lane, float* __restrict__ stg) {
  if (!__any_sync(0xffffffffu, mask != 0u)) return;
  const uint32_t lbit = 1u << lane, lt = lbit - 1u;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) stg[lane * STG_LD + j] = v[j] * rs;
  stg[lane * STG_LD + 32] = __uint_as_float(mask);  // the pad word of the row carries its mask
  __syncwarp();
#pragma unroll UNROLL
  for (int r = 0; r < 32; ++r) {
    const uint32_t mr = __float_as_uint(stg[r * STG_LD + 32]);
    const int br = __shfl_sync(0xffffffffu, base, r);
    const float x = stg[r * STG_LD + lane] * cs;
    if (mr & lbit) out[br + __popc(mr & lt)] = x;
  }
}

// Rows of a filled transposer -> C: lane = column, one 128-byte row segment per store.  Four rows per step, loads
// first, the row pointer carried in a register pair the compiler may not rematerialise (left to itself it emitted a
// branch, a 64-bit multiply and a constant reload per row and serialised the rows: 128 cycles each, measured with
// the timeline of tools/tc_trace.py).  The pad word of a row carries its bias.
template <bool RELU, bool FULL>
__device__ __forceinline__ void tc_store_rows(const float* __restrict__ stg, int lane, float cs, float* dst,
                                              long long ldc, int rows) {
  const long long step = ldc * (long long)sizeof(float);
  unsigned long long a = reinterpret_cast<unsigned long long>(dst);
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 4) {
    float x[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      x[k] = stg[(r0 + k) * STG_LD + lane];
      b[k] = stg[(r0 + k) * STG_LD + 32];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float y = x[k] * cs + b[k];
      if (RELU) y = fmaxf(y, 0.f);
      if (FULL) {
        asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(a), "f"(y) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.s32 q, %2, %3;\n\t@q st.global.cs.f32 [%0], %1;\n\t}"
                     ::"l"(a), "f"(y), "r"(r0 + k), "r"(rows) : "memory");
      }
      a += (unsigned long long)step;
      asm volatile("" : "+l"(a));
    }
  }
}

template <int KIND, bool SDDMM>
__device__ __forceinline__ void tc_tile_store(const TcParams& p, const float (&acc)[128], int mb, int nb, int q, int half,
                                          int ntw, int lane, float* __restrict__ stg, const TilePre& t,
                                          bool tr = false, int* trn = nullptr) {
  const bool relu = (p.flags & 4u) != 0, ct = (p.flags & 2u) != 0;
  const int m0 = mb * BM + q * 32;
  const int m = m0 + lane;
  const float rs = t.rs;
  if constexpr (SDDMM) {
    // sample the tile on A's pattern: bit j of the mask <=> (m, 32 g + j) is a non-zero; its position in the CSR
    // value array is the group's offset plus the number of set bits below j
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int grp = (nb * p.Nt + half * ntw + c * 32) >> 5;
      if (c * 32 < ntw && grp < p.groups)  // warp-uniform
        tc_sddmm_group_store<8>(p.C, reinterpret_cast<const float(&)[32]>(acc[c * 32]), rs, t.cs[c], t.msk[c], t.bs[c],
                                lane, stg);
    }
    return;
  }
  const float bias = t.bias;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int n0 = nb * p.Nt + half * ntw + c * 32;
    if (c * 32 < ntw && n0 < p.N) {
      if (ct) {  // C^T [N, M]: a register is one n, the 32 lanes are 32 consecutive m -> coalesced
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = acc[c * 32 + j];
          if constexpr (KIND == KIND_F16) x = x * rs * __shfl_sync(0xffffffffu, t.cs[c], j);
          x += bias;
          x = relu ? fmaxf(x, 0.f) : x;
          if (m < p.M && n0 + j < p.N) p.C[(long long)(n0 + j) * p.ldc + m] = x;
        }
      } else {   // C [M, N]: through the transposer, one 128-byte row segment per store instruction
        const int n = n0 + lane;
        const float cs = t.cs[c];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * STG_LD + j] = acc[c * 32 + j] * rs;
        stg[lane * STG_LD + 32] = bias;  // the pad word of the row carries its bias
        __syncwarp();
        if (tr) tc_trace(true, 2, *trn, 30 + c);  // transposer filled
        float* dst = p.C + (long long)m0 * p.ldc + n;
        const int rows = (n < p.N) ? min(p.M - m0, 32) : 0;  // rows of this warp's block that exist (for this lane's column)
        const bool full = __all_sync(0xffffffffu, rows == 32);
        if (full) {
          if (relu) tc_store_rows<true, true>(stg, lane, cs, dst, p.ldc, 32);
          else tc_store_rows<false, true>(stg, lane, cs, dst, p.ldc, 32);
        } else {
          if (relu) tc_store_rows<true, false>(stg, lane, cs, dst, p.ldc, rows);
          else tc_store_rows<false, false>(stg, lane, cs, dst, p.ldc, rows);
        }
        if (tr) tc_trace(true, 2, *trn, 40 + c);  // group written
      }
    }
  }
}

// The tensor core adds into its TMEM accumulator with truncation (measured: a K = 16384 contraction of positive
// values comes out 1.3e-4 low), so an accumulator only ever holds `chunk` k-tiles (<= 48 additions, bias <= 1.5e-6
// of the partial sum); the epilogue warps add the partial sums in fp32 registers with round-to-nearest.  The two
// TMEM buffers alternate per chunk: MMAs of chunk i+1 overlap the drain of chunk i.
//
// SDDMM = true: the same pipeline computes D = G * B^T tile by tile (contraction over the dense column index) and the
// epilogue writes only the entries on A's pattern: grad of the CSR values (matmuls.py:250 restricted to the pattern).
template <int KIND, bool SDDMM>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_spmm_kernel(const TcParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  if (*reinterpret_cast<const volatile int*>(p.ctl) != 0) {  // Inf / NaN in a dense operand (uniform over the grid)
    if constexpr (SDDMM) {
      csr_order_sddmm(p.row_ptr, p.col, p.M, p.N, p.G, p.g_sm, p.g_sn, p.B, p.b_sk, p.b_sn, p.C);
    } else {
      const bool ctr = (p.flags & 2u) != 0;
      csr_order_spmm(p.row_ptr, p.col, p.val, p.M, p.N, p.B, p.b_sk, p.b_sn, p.C, ctr ? 1 : p.ldc, ctr ? p.ldc : 1,
                     p.bias, (p.flags & 4u) != 0);
    }
    tc_exit_ticket(p.ctl);
    return;
  }

  const uint32_t raw = smem_u32(tc_smem_raw);
  unsigned char* smem = tc_smem_raw + (((raw + 1023u) & ~1023u) - raw);  // swizzle atoms need 1 KB alignment
  const int stage_bytes = A_TILE + p.Nt * 256;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, EPI_WARPS);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int n_tiles = p.MB * p.NB;
  if (warp == 0) {
    // ===== TMA producer: one elected lane streams (A tile, B tile) pairs into the smem ring =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      const bool tr = (p.terms & 512) && blockIdx.x == 0;
      int trn = 0;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
        const int mb = tile / p.NB, nb = tile % p.NB;
        const int t0 = SDDMM ? 0 : __ldg(p.blk_ptr + mb), t1 = SDDMM ? p.KT : __ldg(p.blk_ptr + mb + 1);
        for (int t = t0; t < t1; ++t) {
          tc_trace(tr, 0, trn, 1);  // producer: starts waiting for the slot
          ok = mbar_wait_bounded(empty + stage, phase ^ 1u, p.ctl + 1, 100 + stage);
          if (!ok) break;
          tc_trace(tr, 0, trn, 2);  // producer: slot free, copies issued next
          const int kt = SDDMM ? t : __ldg(p.blk_kt + t);
          unsigned char* sa = smem + (size_t)stage * stage_bytes;
          const bool cp_a = !(p.terms & 32), cp_b = !(p.terms & 16);  // (timing studies: skip a copy)
          const uint32_t tx = (cp_a ? (uint32_t)A_TILE : 0u) + (cp_b ? (uint32_t)p.Nt * 256u : 0u);
          if (tx) mbar_expect_tx(full + stage, tx); else mbar_arrive(full + stage);
          const unsigned char* a_src = SDDMM ? p.a_tiles + ((size_t)kt * p.NGa + (size_t)mb * (BM / 8)) * ATOM
                                             : p.a_tiles + (size_t)t * A_TILE;
          if (cp_a) tma_bulk_g2s(sa, a_src, A_TILE, full + stage);
          if (cp_b)
            tma_bulk_g2s(sa + A_TILE, p.b_pack + ((size_t)kt * p.NG + (size_t)nb * (p.Nt / 8)) * ATOM,
                         (uint32_t)p.Nt * 256u, full + stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the warp walks the pipeline together, one elected lane issues every tcgen05 instruction =====
    {
      const uint32_t idesc = umma_idesc(KIND, p.Nt);
      const bool all_terms = (p.terms & 7) == 7;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;  // chunks issued so far: TMEM buffer g & 1, its barrier phase (g >> 1) & 1
      bool ok = true;
      const bool tr = (p.terms & 512) && blockIdx.x == 0 && lane == 0;
      int trn = 0;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
        const int mb = tile / p.NB;
        const int t0 = SDDMM ? 0 : __ldg(p.blk_ptr + mb), t1 = SDDMM ? p.KT : __ldg(p.blk_ptr + mb + 1);
        for (int c0 = t0; c0 < t1 && ok; c0 += p.chunk, ++g) {
          const int as = (int)(g & 1u);
          tc_trace(tr, 1, trn, 10);  // MMA lane: starts waiting for a free accumulator
          ok = __all_sync(0xffffffffu, mbar_wait_bounded(acc_empty + as, ((g >> 1) & 1u) ^ 1u, p.ctl + 1, 200 + as));
          if (!ok) break;
          tc_trace(tr, 1, trn, 11);  // accumulator free
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(as * p.Nt);
          uint32_t acc = 0;
          const int c1 = min(c0 + p.chunk, t1);
          for (int t = c0; t < c1; ++t) {
            tc_trace(tr, 1, trn, 12);  // MMA lane: starts waiting for operands
            ok = __all_sync(0xffffffffu, mbar_wait_bounded(full + stage, phase, p.ctl + 1, 300 + stage));
            if (!ok) break;
            tc_trace(tr, 1, trn, 13);  // operands landed
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t sb = sa + A_TILE;
            __syncwarp();
            if (elect_one()) {
              if (all_terms) {  // the shipped configuration: straight-line issue, descriptor arithmetic hoisted
                umma_ktile_all_terms<KIND, 1>(d, sa, idesc, acc);
              } else {
                uint32_t a2 = acc;
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // one MMA consumes 32 bytes of every swizzled 128-byte row
                  const uint64_t a_hi = umma_desc(sa + k * 32, ATOM), a_lo = umma_desc(sa + 1024 + k * 32, ATOM);
                  const uint64_t b_hi = umma_desc(sb + k * 32, ATOM), b_lo = umma_desc(sb + 1024 + k * 32, ATOM);
                  if (p.terms & 2) { umma<KIND>(d, a_hi, b_lo, idesc, a2); a2 = 1; }
                  if (p.terms & 4) { umma<KIND>(d, a_lo, b_hi, idesc, a2); a2 = 1; }
                  if (p.terms & 1) { umma<KIND>(d, a_hi, b_hi, idesc, a2); a2 = 1; }
                }
              }
              umma_commit(empty + stage);  // the smem slot is free once these MMAs have read it
              if (t + 1 == c1) umma_commit(acc_full + as);  // partial sums of this chunk complete -> epilogue warps
            }
            __syncwarp();
            if (p.terms & 7) acc = 1;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM -> fp32 registers (+=) per chunk; at tile end (+bias, ReLU) -> C =====
    const int q = warp & 3;               // TMEM lane quarter this warp may read (hardware rule: warp id % 4)
    const int half = (warp - 2) >> 2;     // which half of the tile's columns
    const int ntw = p.Nt >> 1;            // columns per warp (multiple of 32, <= 128)
    float* stg = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes + 256) + (warp - 2) * (32 * STG_LD);
    uint32_t g = 0;
    bool ok = true;
    const bool etr = (p.terms & 512) && blockIdx.x == 0 && warp == 2 && lane == 0;
    int etrn = 0;
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
      const int mb = tile / p.NB, nb = tile % p.NB;
      const int t0 = SDDMM ? 0 : __ldg(p.blk_ptr + mb), t1 = SDDMM ? p.KT : __ldg(p.blk_ptr + mb + 1);
      if constexpr (SDDMM) {
        if (t1 - t0 <= p.chunk && !(p.terms & 8)) {
          // short contraction (one chunk): nothing to add up, so every group goes TMEM -> transposer -> pattern entries
          // without the 128 accumulator registers: the store loop gets them for 32 iterations in flight
          const int as = (int)(g & 1u);
          ok = mbar_wait_bounded(acc_full + as, (g >> 1) & 1u, p.ctl + 1, 400 + as);
          if (!ok) break;
          tc_fence_after();
          TilePre pre;
          tc_tile_prefetch<KIND, true>(p, mb, nb, q, half, ntw, lane, pre);
          const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.Nt + half * ntw);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c * 32 < ntw) {
              uint32_t v[32];
              tmem_ld32(t_addr + (uint32_t)(c * 32), v);
              tmem_ld_wait();
              if ((c + 1) * 32 >= ntw) {  // last group read: hand the TMEM buffer back before the stores
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + as);
              }
              const int grp = (nb * p.Nt + half * ntw + c * 32) >> 5;
              if (grp < p.groups) {
                float vf[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) vf[j] = __uint_as_float(v[j]);
                tc_sddmm_group_store<32>(p.C, vf, pre.rs, pre.cs[c], pre.msk[c], pre.bs[c], lane, stg);
              }
            }
          }
          ++g;
          continue;
        }
      }
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int c0 = t0; c0 < t1 && ok; c0 += p.chunk, ++g) {
        const int as = (int)(g & 1u);
        tc_trace(etr, 2, etrn, 20);  // epilogue: starts waiting for a chunk's sums
        ok = mbar_wait_bounded(acc_full + as, (g >> 1) & 1u, p.ctl + 1, 400 + as);
        if (!ok) break;
        tc_trace(etr, 2, etrn, 21);  // sums complete, drain starts
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.Nt + half * ntw);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < ntw) {
            uint32_t v[32];
            tmem_ld32(t_addr + (uint32_t)(c * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + as);  // this warp's part of the buffer is drained
        tc_trace(etr, 2, etrn, 22);  // drained
      }
      if (!ok) break;
      tc_trace(etr, 2, etrn, 23);  // tile store starts
      if (!(p.terms & 8)) {
        // (fetching these before the accumulator wait was measured slower: 14 more live registers spill in the drain)
        TilePre pre;
        tc_tile_prefetch<KIND, SDDMM>(p, mb, nb, q, half, ntw, lane, pre);
        tc_trace(etr, 2, etrn, 25);  // scales / bias fetched
        tc_tile_store<KIND, SDDMM>(p, acc, mb, nb, q, half, ntw, lane, stg, pre, etr, &etrn);
      }
      tc_trace(etr, 2, etrn, 24);  // tile stored
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  tc_exit_ticket(p.ctl);
}

// ---- SDDMM, short contractions: transposed product, 16 epilogue warps ----------------------------------------------
// When the whole contraction fits one TMEM accumulation (KT <= chunk: every N of config 5), a tile's time is its
// sampled stores, and in tc_spmm_kernel those are bound by the LSU / MIO instruction rate: per 32 x 32 group 33 STS +
// 64 LDS + 32 SHFL + 32 STG through the transposer (ncu: mio_throttle is the first stall reason, LSU pipe 59 %).
// This variant computes D^T instead: the packed B block (128 pattern COLUMNS) is the MMA's M operand and the packed
// G block (128 pattern ROWS) its N operand — both operands have the same packed layout, so this is a swap of two
// pointers.  A TMEM lane is then a pattern column and a register a pattern row: exactly what the row-wise store needs
// (the lanes of one instruction hold 32 consecutive columns of one row), with NO transposer.  Per group: one STS.128
// (each lane publishes mask / offset / scale of its row), then per row one broadcast LDS.128 and one STG.
// 16 epilogue warps (no accumulator registers needed: 64 registers per thread), tiles of 128 x 128, warp w owns the
// group (TMEM lane quarter w % 4 = 32 columns, register block (w - 2) / 4 = 32 rows), four TMEM buffers, three 64 KB
// operand stages; pattern words and scales of the NEXT tile are fetched before the wait for the current one.
constexpr int SH_EPI_WARPS = 16;
constexpr int SH_THREADS = (2 + SH_EPI_WARPS) * 32;
constexpr int SH_SIDE_BYTES = 32 * 16;  // per warp: (mask, offset, row scale, -) of its 32 rows
// NT = pattern rows per tile (the MMA's N): 128 (three 64 KB stages, four TMEM buffers, one group per warp and tile) or
// 256 (two 96 KB stages, two TMEM buffers, two groups per warp and tile: fewer operand bytes per flop for N > 128)
template <int NT> struct ShCfg {
  static constexpr int STAGES = NT == 128 ? 3 : 2;
  static constexpr int STAGE_BYTES = A_TILE + NT * 256;
  static constexpr int BUFS = 512 / NT;
  static constexpr int GPW = NT / 128;  // groups per warp and tile
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + (size_t)SH_EPI_WARPS * SH_SIDE_BYTES;
};

struct GroupPre {
  float rs, cs;   // scale of this lane's ROW (lane = row while fetching) / of this lane's COLUMN
  uint32_t msk;   // pattern bits of (row = lane, the group's 32 columns)
  int bs;         // position of that row's first entry of the group in the CSR value array
};

// q: which 32 columns of the tile's 128 (TMEM lane quarter), rb: which 32 rows
template <int KIND, int NT>
__device__ __forceinline__ GroupPre tc_group_prefetch(const TcParams& p, int tile, int q, int rb, int lane) {
  const int mb = tile / p.NB, nb = tile % p.NB;
  const int m = mb * NT + rb * 32 + lane;
  const int grp = nb * (BM / 32) + q;
  GroupPre t;
  t.rs = 1.f; t.cs = 1.f; t.msk = 0u; t.bs = 0;
  if (grp < p.groups) {
    if constexpr (KIND == KIND_F16) t.cs = scale_down(scale_exp(__ldg(p.colmax + grp * 32 + lane)));
    if (m < p.M) {
      t.msk = __ldg(p.samp_mask + (size_t)grp * p.M + m);
      t.bs = __ldg(p.samp_off + (size_t)grp * p.M + m);
      if constexpr (KIND == KIND_F16) t.rs = scale_down(scale_exp(__ldg(p.amax + m)));
    }
  }
  return t;
}

template <int KIND, int NT>
__global__ void __launch_bounds__(SH_THREADS, 1) tc_sddmm_short_kernel(const TcParams p) {
  constexpr int SH_NT = NT, SH_STAGES = ShCfg<NT>::STAGES, SH_STAGE_BYTES = ShCfg<NT>::STAGE_BYTES;
  constexpr int SH_BUFS = ShCfg<NT>::BUFS, GPW = ShCfg<NT>::GPW;
  extern __shared__ unsigned char tc_smem_raw[];
  if (*reinterpret_cast<const volatile int*>(p.ctl) != 0) {  // Inf / NaN in a dense operand (uniform over the grid)
    csr_order_sddmm(p.row_ptr, p.col, p.M, p.N, p.G, p.g_sm, p.g_sn, p.B, p.b_sk, p.b_sn, p.C);
    tc_exit_ticket(p.ctl);
    return;
  }
  const uint32_t raw = smem_u32(tc_smem_raw);
  unsigned char* smem = tc_smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)SH_STAGES * SH_STAGE_BYTES);
  uint64_t* empty = full + SH_STAGES;
  uint64_t* acc_full = empty + SH_STAGES;
  uint64_t* acc_empty = acc_full + SH_BUFS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + SH_BUFS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SH_STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < SH_BUFS; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, SH_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const int n_tiles = p.MB * p.NB;  // tile = (block of NT pattern rows mb, block of 128 pattern columns nb)

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: M operand = 128 packed rows of B (pattern columns), N operand = 128 of G =====
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
        const int mb = tile / p.NB, nb = tile % p.NB;
        for (int kt = 0; kt < p.KT; ++kt) {
          ok = mbar_wait_bounded(empty + stage, phase ^ 1u, p.ctl + 1, 100 + stage);
          if (!ok) break;
          unsigned char* sa = smem + (size_t)stage * SH_STAGE_BYTES;
          mbar_expect_tx(full + stage, (uint32_t)SH_STAGE_BYTES);
          tma_bulk_g2s(sa, p.b_pack + ((size_t)kt * p.NG + (size_t)nb * (BM / 8)) * ATOM, A_TILE, full + stage);
          tma_bulk_g2s(sa + A_TILE, p.a_tiles + ((size_t)kt * p.NGa + (size_t)mb * (SH_NT / 8)) * ATOM,
                       (uint32_t)SH_NT * 256u, full + stage);
          if (++stage == SH_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {  // ===== MMA issuer (warp-wide loop, one elected lane issues): a tile's whole contraction goes into one buffer =====
      const uint32_t idesc = umma_idesc(KIND, SH_NT);
      const bool all_terms = (p.terms & 7) == 7;
      int stage = 0;
      uint32_t phase = 0, g = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++g) {
        const uint32_t as = g % SH_BUFS;
        ok = __all_sync(0xffffffffu,
                        mbar_wait_bounded(acc_empty + as, ((g / SH_BUFS) & 1u) ^ 1u, p.ctl + 1, 200 + (int)as));
        if (!ok) break;
        tc_fence_after();
        const uint32_t d = tmem_base + as * SH_NT;
        uint32_t acc = 0;
        for (int kt = 0; kt < p.KT; ++kt) {
          ok = __all_sync(0xffffffffu, mbar_wait_bounded(full + stage, phase, p.ctl + 1, 300 + stage));
          if (!ok) break;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * SH_STAGE_BYTES);
          const uint32_t sb = sa + A_TILE;
          __syncwarp();
          if (elect_one()) {
            if (all_terms) {
              umma_ktile_all_terms<KIND, 1>(d, sa, idesc, acc);
            } else {
              uint32_t a2 = acc;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t a_hi = umma_desc(sa + k * 32, ATOM), a_lo = umma_desc(sa + 1024 + k * 32, ATOM);
                const uint64_t b_hi = umma_desc(sb + k * 32, ATOM), b_lo = umma_desc(sb + 1024 + k * 32, ATOM);
                if (p.terms & 2) { umma<KIND>(d, a_hi, b_lo, idesc, a2); a2 = 1; }
                if (p.terms & 4) { umma<KIND>(d, a_lo, b_hi, idesc, a2); a2 = 1; }
                if (p.terms & 1) { umma<KIND>(d, a_hi, b_hi, idesc, a2); a2 = 1; }
              }
            }
            umma_commit(empty + stage);
            if (kt + 1 == p.KT) umma_commit(acc_full + as);
          }
          __syncwarp();
          if (p.terms & 7) acc = 1;
          if (++stage == SH_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue warps 2..17: one 32 x 32 group of every tile, TMEM -> registers -> entries on A's pattern =====
    const int q = warp & 3;          // TMEM lane quarter this warp may read (hardware rule: warp id % 4): 32 columns
    const int rb = (warp - 2) >> 2;  // which 32 rows of the tile (TMEM columns)
    uint4* side = reinterpret_cast<uint4*>(smem + (size_t)SH_STAGES * SH_STAGE_BYTES + 256) + (warp - 2) * 32;
    const uint32_t lbit = 1u << lane, lt = lbit - 1u;
    const uint32_t side_a = smem_u32(side);
    unsigned long long outp = reinterpret_cast<unsigned long long>(p.C);
    asm volatile("" : "+l"(outp));  // (kept in a register pair: no constant reload per row)
    uint32_t g = 0;
    bool ok = true;
    int tile = blockIdx.x;
    GroupPre nxt[GPW];
#pragma unroll
    for (int h = 0; h < GPW; ++h) {
      nxt[h] = GroupPre{1.f, 1.f, 0u, 0};
      if (tile < n_tiles) nxt[h] = tc_group_prefetch<KIND, NT>(p, tile, q, rb + 4 * h, lane);
    }
    for (; tile < n_tiles && ok; tile += gridDim.x, ++g) {
      GroupPre cur[GPW];
#pragma unroll
      for (int h = 0; h < GPW; ++h) {
        cur[h] = nxt[h];
        if (tile + (int)gridDim.x < n_tiles)
          nxt[h] = tc_group_prefetch<KIND, NT>(p, tile + (int)gridDim.x, q, rb + 4 * h, lane);
      }
      const uint32_t as = g % SH_BUFS;
      ok = mbar_wait_bounded(acc_full + as, (g / SH_BUFS) & 1u, p.ctl + 1, 400 + (int)as);
      if (!ok) break;
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < GPW; ++h) {
        uint32_t v[32];  // v[j] = D^T(column of this lane, row (rb + 4 h) * 32 + j)
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * SH_NT + (uint32_t)((rb + 4 * h) * 32), v);
        tmem_ld_wait();
        if (h == GPW - 1) {  // last read of the buffer: it goes back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + as);
        }
        if (!__any_sync(0xffffffffu, cur[h].msk != 0u)) continue;  // (also: group beyond K, rows beyond M)
        __syncwarp();  // everybody has read `side` before it is overwritten
        side[lane] = make_uint4(cur[h].msk, (uint32_t)cur[h].bs, __float_as_uint(cur[h].rs), 0u);
        __syncwarp();
        const float cs = cur[h].cs;
        // hand-structured (left to itself the compiler emitted a branch, three scalar LDS and a reload of the output
        // pointer per row): one broadcast LDS.128 per row, straight-line ALU work, one predicated STG
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          uint32_t mk, bs, sc, pad;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(mk), "=r"(bs), "=r"(sc), "=r"(pad)
                       : "r"(side_a + (uint32_t)j * 16u));
          const float x = __uint_as_float(v[j]) * cs * __uint_as_float(sc);
          const int idx = (int)bs + __popc(mk & lt);
          asm volatile(
              "{\n\t.reg .pred q;\n\t.reg .b32 t;\n\t.reg .b64 a;\n\t"
              "and.b32 t, %0, %1;\n\t"
              "setp.ne.u32 q, t, 0;\n\t"
              "mad.wide.s32 a, %2, 4, %3;\n\t"
              "@q st.global.f32 [a], %4;\n\t}"
              ::"r"(mk), "r"(lbit), "r"(idx), "l"(outp), "f"(x)
              : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  tc_exit_ticket(p.ctl);
}

// ---- CTA-pair variant: tcgen05.mma.cta_group::2 ------------------------------------------------------------------
// Two CTAs of a cluster (one TPC) multiply a 256 x Nt tile together: each holds its own 128 rows of A and HALF of
// the B tile (Nt / 2 columns), the tensor cores of both SMs read both halves, each SM's TMEM receives its 128 rows of
// the result.  Per stage and CTA: A 32 KB + B Nt * 128 B (64 KB at Nt = 256) instead of 96 KB: one third fewer bytes
// through each SM's L2 port (the limiter of the single-CTA kernel, profiles/r2_ncu_tc.md) and 3 stages instead of 2.
// Protocol (all mbarriers live at the same offsets in both CTAs; rank 0 = leader issues every MMA):
//   full[s]      leader: own TMA bytes + one arrive relayed by the peer (its warp 1 waits for the peer's own full[s],
//                then arrives remotely); peer: own TMA bytes
//   empty[s]     released in BOTH CTAs by the leader's tcgen05.commit ... multicast::cluster
//   acc_full[b]  published in both CTAs the same way; acc_empty[b] lives on the leader and collects the 16 epilogue
//                warps of both CTAs (remote arrives)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// bounded wait with cluster-scope acquire (the barrier may be completed by the other CTA)
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* dbg, int code) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return true;
    if ((it & 1023u) == 1023u) {
      if (*reinterpret_cast<volatile int*>(dbg) != 0) return false;
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {
        atomicCAS(dbg, 0, code);
        return false;
      }
    }
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                      uint32_t accumulate) {
  if constexpr (KIND == KIND_F16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <int KIND, bool SDDMM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) tc_spmm2_kernel(const TcParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  if (*reinterpret_cast<const volatile int*>(p.ctl) != 0) {  // Inf / NaN in a dense operand (uniform over the grid)
    if constexpr (SDDMM) {
      csr_order_sddmm(p.row_ptr, p.col, p.M, p.N, p.G, p.g_sm, p.g_sn, p.B, p.b_sk, p.b_sn, p.C);
    } else {
      const bool ctr = (p.flags & 2u) != 0;
      csr_order_spmm(p.row_ptr, p.col, p.val, p.M, p.N, p.B, p.b_sk, p.b_sn, p.C, ctr ? 1 : p.ldc, ctr ? p.ldc : 1,
                     p.bias, (p.flags & 4u) != 0);
    }
    tc_exit_ticket(p.ctl);
    return;
  }
  const uint32_t rank = cluster_ctarank();
  const uint32_t raw = smem_u32(tc_smem_raw);
  unsigned char* smem = tc_smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int b_bytes = p.Nt * 128;                 // this CTA's half of the B tile (Nt / 2 rows, hi + lo)
  const int stage_bytes = A_TILE + b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full + i, rank == 0 ? 2 : 1);
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, 2 * EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int n_tiles = p.MB2 * p.NB;
  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A tile + own half of the B tile =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = pair; tile < n_tiles && ok; tile += n_pairs) {
        const int mp = tile / p.NB, nb = tile % p.NB;
        const int mb = 2 * mp + (int)rank;
        const int t0 = SDDMM ? 0 : __ldg(p.pair_ptr + mp), t1 = SDDMM ? p.KT : __ldg(p.pair_ptr + mp + 1);
        for (int t = t0; t < t1; ++t) {
          ok = mbar_wait_cluster(empty + stage, phase ^ 1u, p.ctl + 1, 100 + stage);
          if (!ok) break;
          const int kt = SDDMM ? t : __ldg(p.pair_kt + t);
          unsigned char* sa = smem + (size_t)stage * stage_bytes;
          const bool cp_a = !(p.terms & 32), cp_b = !(p.terms & 16);  // (timing studies: skip a copy)
          const uint32_t tx = (cp_a ? (uint32_t)A_TILE : 0u) + (cp_b ? (uint32_t)b_bytes : 0u);
          if (tx) mbar_expect_tx(full + stage, tx); else mbar_arrive(full + stage);
          const unsigned char* a_src =
              SDDMM ? p.a_tiles + ((size_t)kt * p.NGa + (size_t)mb * (BM / 8)) * ATOM
                    : p.a_tiles + (size_t)__ldg(p.pair_slot + (size_t)rank * p.n_pair_tiles + t) * A_TILE;
          if (cp_a) tma_bulk_g2s(sa, a_src, A_TILE, full + stage);
          if (cp_b)
            tma_bulk_g2s(sa + A_TILE,
                         p.b_pack + ((size_t)kt * p.NG + (size_t)nb * (p.Nt / 8) + (size_t)rank * (p.Nt / 16)) * ATOM,
                         (uint32_t)b_bytes, full + stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 1) {
      // ===== relay (peer CTA): my operands of a stage have landed -> tell the leader's full barrier =====
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = pair; tile < n_tiles && ok; tile += n_pairs) {
        const int mp = tile / p.NB;
        const int t0 = SDDMM ? 0 : __ldg(p.pair_ptr + mp), t1 = SDDMM ? p.KT : __ldg(p.pair_ptr + mp + 1);
        for (int t = t0; t < t1; ++t) {
          ok = mbar_wait_cluster(full + stage, phase, p.ctl + 1, 500 + stage);
          if (!ok) break;
          mbar_arrive_remote(full + stage, 0);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (rank == 0) {
      // ===== MMA issuer (leader CTA): the warp walks the pipeline, one elected lane issues every tcgen05 instruction =====
      const uint32_t idesc = (1u << 4) | ((KIND == KIND_F16 ? 0u : 2u) << 7) | ((KIND == KIND_F16 ? 0u : 2u) << 10) |
                             ((uint32_t)(p.Nt >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const bool all_terms = (p.terms & 7) == 7;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;
      bool ok = true;
      for (int tile = pair; tile < n_tiles && ok; tile += n_pairs) {
        const int mp = tile / p.NB;
        const int t0 = SDDMM ? 0 : __ldg(p.pair_ptr + mp), t1 = SDDMM ? p.KT : __ldg(p.pair_ptr + mp + 1);
        for (int c0 = t0; c0 < t1 && ok; c0 += p.chunk, ++g) {
          const int as = (int)(g & 1u);
          ok = __all_sync(0xffffffffu, mbar_wait_cluster(acc_empty + as, ((g >> 1) & 1u) ^ 1u, p.ctl + 1, 200 + as));
          if (!ok) break;
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(as * p.Nt);
          uint32_t acc = 0;
          const int c1 = min(c0 + p.chunk, t1);
          for (int t = c0; t < c1; ++t) {
            ok = __all_sync(0xffffffffu, mbar_wait_cluster(full + stage, phase, p.ctl + 1, 300 + stage));
            if (!ok) break;
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t sb = sa + A_TILE;
            __syncwarp();
            if (elect_one()) {
              if (all_terms) {
                umma_ktile_all_terms<KIND, 2>(d, sa, idesc, acc);
              } else {
                uint32_t a2 = acc;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t a_hi = umma_desc(sa + k * 32, ATOM), a_lo = umma_desc(sa + 1024 + k * 32, ATOM);
                  const uint64_t b_hi = umma_desc(sb + k * 32, ATOM), b_lo = umma_desc(sb + 1024 + k * 32, ATOM);
                  if (p.terms & 2) { umma2<KIND>(d, a_hi, b_lo, idesc, a2); a2 = 1; }
                  if (p.terms & 4) { umma2<KIND>(d, a_lo, b_hi, idesc, a2); a2 = 1; }
                  if (p.terms & 1) { umma2<KIND>(d, a_hi, b_hi, idesc, a2); a2 = 1; }
                }
              }
              umma_commit2(empty + stage);  // frees the slot in both CTAs
              if (t + 1 == c1) umma_commit2(acc_full + as);  // both CTAs' epilogue warps
            }
            __syncwarp();
            if (p.terms & 7) acc = 1;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): own 128 rows of the pair's tile =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ntw = p.Nt >> 1;
    float* stg = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes + 256) + (warp - 2) * (32 * STG_LD);
    uint32_t g = 0;
    bool ok = true;
    for (int tile = pair; tile < n_tiles && ok; tile += n_pairs) {
      const int mp = tile / p.NB, nb = tile % p.NB;
      const int mb = 2 * mp + (int)rank;
      const int t0 = SDDMM ? 0 : __ldg(p.pair_ptr + mp), t1 = SDDMM ? p.KT : __ldg(p.pair_ptr + mp + 1);
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int c0 = t0; c0 < t1 && ok; c0 += p.chunk, ++g) {
        const int as = (int)(g & 1u);
        ok = mbar_wait_cluster(acc_full + as, (g >> 1) & 1u, p.ctl + 1, 400 + as);
        if (!ok) break;
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.Nt + half * ntw);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < ntw) {
            uint32_t v[32];
            tmem_ld32(t_addr + (uint32_t)(c * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(acc_empty + as, 0);  // the leader's barrier collects both CTAs
      }
      if (!ok) break;
      if (mb < p.MB && !(p.terms & 8)) {
        TilePre pre;
        tc_tile_prefetch<KIND, SDDMM>(p, mb, nb, q, half, ntw, lane, pre);
        tc_tile_store<KIND, SDDMM>(p, acc, mb, nb, q, half, ntw, lane, stg, pre);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the other CTA may still signal it or read its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
  tc_exit_ticket(p.ctl);
}

// Generic strided CSR SpMM: warp per row, lanes over 32 dense columns, any element strides for B and C, bias
// and ReLU.  Slow and simple on purpose: it is what runs when the tensor-core path has to stand down (Inf/NaN in
// the dense operand) or when transposed operands are requested for a matrix that is too sparse for it.
__global__ void __launch_bounds__(256)
spmm_strided_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const float* __restrict__ val,
                    int M, int N, const float* __restrict__ B, long long b_sk, long long b_sn, float* __restrict__ C,
                    long long c_sm, long long c_sn, const float* __restrict__ bias, int relu) {
  csr_order_spmm(row_ptr, col, val, M, N, B, b_sk, b_sn, C, c_sm, c_sn, bias, relu != 0);
}

__global__ void __launch_bounds__(256)
sddmm_strided_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int M, int N,
                     const float* __restrict__ G, long long g_sm, long long g_sn, const float* __restrict__ B,
                     long long b_sk, long long b_sn, float* __restrict__ out) {
  csr_order_sddmm(row_ptr, col, M, N, G, g_sm, g_sn, B, b_sk, b_sn, out);
}

}  // namespace

// generic strided SDDMM (any G / B strides): small problems whose operands are held transposed
int launch_sddmm_strided(const int* row_ptr, const int* col, int M, int N, const float* G, long long g_sm,
                         long long g_sn, const float* B, long long b_sk, long long b_sn, float* out, cudaStream_t s) {
  if (M == 0) return B200_OK;
  const int blocks = (int)std::min<long long>(((long long)M + 7) / 8, (long long)device_sm_count() * 16);
  sddmm_strided_kernel<<<blocks, 256, 0, s>>>(row_ptr, col, M, N, G, g_sm, g_sn, B, b_sk, b_sn, out);
  B200_LAUNCH_OK("sddmm_strided_kernel");
  return B200_OK;
}

int launch_spmm_strided(const int* row_ptr, const int* col, const float* val, int M, int N, const float* B,
                        long long b_sk, long long b_sn, float* C, long long c_sm, long long c_sn,
                        const float* bias, bool relu, cudaStream_t s) {
  if (M == 0 || N == 0) return B200_OK;
  const long long warps = (long long)M * ((N + 31) / 32);
  const int blocks = (int)std::min<long long>((warps + 7) / 8, (long long)device_sm_count() * 16);
  spmm_strided_kernel<<<blocks, 256, 0, s>>>(row_ptr, col, val, M, N, B, b_sk, b_sn, C, c_sm, c_sn, bias,
                                              relu ? 1 : 0);
  B200_LAUNCH_OK("spmm_strided_kernel");
  return B200_OK;
}

// ---- host side ------------------------------------------------------------------------------------------------

void tc_store_free(TcStore* st) {
  cudaFree(st->blk_ptr); cudaFree(st->blk_kt); cudaFree(st->scan); cudaFree(st->tiles); cudaFree(st->row_scale);
  cudaFree(st->pair_ptr); cudaFree(st->pair_kt); cudaFree(st->pair_slot);
  *st = TcStore{};
}

static int tc_launch_pack_a(const int* row_ptr, const int* col, const float* val, int nnz, TcStore* st, cudaStream_t s) {
  if (nnz == 0) return B200_OK;
  const unsigned wblocks = (unsigned)(((size_t)st->M * 32 + 255) / 256);
  if (st->kind == KIND_F16)
    tc_pack_a_kernel<KIND_F16><<<wblocks, 256, 0, s>>>(row_ptr, col, val, st->M, st->KT, st->scan, st->tiles, st->row_scale);
  else
    tc_pack_a_kernel<KIND_TF32><<<wblocks, 256, 0, s>>>(row_ptr, col, val, st->M, st->KT, st->scan, st->tiles, st->row_scale);
  B200_LAUNCH_OK("tc_pack_a_kernel");
  return B200_OK;
}

// Densify the non-empty 128 x BK tiles of a CSR matrix into hi/lo swizzled tiles.  Synchronises the stream once
// (the number of non-empty tiles sizes the allocation).
int tc_store_build(const int* row_ptr, const int* col, const float* val, int M, int K, int nnz, int kind, TcStore* st,
                   cudaStream_t s) {
  *st = TcStore{};
  st->M = M; st->K = K; st->kind = kind;
  const int bk = bk_of(kind);
  st->MB = (M + BM - 1) / BM;
  st->KT = (K + bk - 1) / bk;
  const size_t cells = (size_t)st->MB * st->KT;
  if (cells == 0 || cells > ((size_t)1 << 26)) {
    set_error("block-dense store: %zu candidate tiles is outside the supported range", cells);
    return B200_ERR_INVALID;
  }
  int* flags = nullptr;
  B200_CUDA_OK(cudaMalloc(&flags, cells * sizeof(int)));
  B200_CUDA_OK(cudaMalloc(&st->scan, cells * sizeof(int)));
  B200_CUDA_OK(cudaMalloc(&st->blk_ptr, ((size_t)st->MB + 1) * sizeof(int)));
  B200_CUDA_OK(cudaMalloc(&st->row_scale, (size_t)std::max(M, 1) * sizeof(float)));
  B200_CUDA_OK(cudaMemsetAsync(flags, 0, cells * sizeof(int), s));
  const unsigned wblocks = (unsigned)(((size_t)M * 32 + 255) / 256);
  if (nnz > 0) {
    tc_mark_tiles_kernel<<<wblocks, 256, 0, s>>>(row_ptr, col, M, st->KT, bk, flags);
    B200_LAUNCH_OK("tc_mark_tiles_kernel");
  }
  size_t tmp_bytes = 0;
  B200_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flags, st->scan, (int)cells, s));
  void* tmp = nullptr;
  B200_CUDA_OK(cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)));
  B200_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flags, st->scan, (int)cells, s));
  int last[2] = {0, 0};
  B200_CUDA_OK(cudaMemcpyAsync(&last[0], st->scan + cells - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
  B200_CUDA_OK(cudaMemcpyAsync(&last[1], flags + cells - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
  B200_CUDA_OK(cudaStreamSynchronize(s));
  st->n_tiles = last[0] + last[1];
  B200_CUDA_OK(cudaMalloc(&st->blk_kt, std::max<size_t>((size_t)st->n_tiles, 1) * sizeof(int)));
  tc_list_tiles_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(flags, st->scan, st->MB, st->KT, st->blk_ptr,
                                                                       st->blk_kt);
  B200_LAUNCH_OK("tc_list_tiles_kernel");
  const size_t bytes = ((size_t)st->n_tiles + 1) * A_TILE;  // + one all-zero tile (slot n_tiles) for the CTA-pair kernel
  if (cudaMalloc(&st->tiles, bytes) != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMalloc of %zu bytes for the block-dense tiles failed", bytes);
    cudaFree(flags); cudaFree(tmp);
    tc_store_free(st);
    return B200_ERR_ALLOC;
  }
  st->tile_bytes = bytes;
  B200_CUDA_OK(cudaMemsetAsync(st->tiles, 0, bytes, s));
  int rc = tc_launch_pack_a(row_ptr, col, val, nnz, st, s);
  if (rc) return rc;
  {  // lists for the CTA-pair kernel
    st->MB2 = (st->MB + 1) / 2;
    const size_t pcells = (size_t)st->MB2 * st->KT;
    int *pflags = nullptr, *pscan = nullptr;
    B200_CUDA_OK(cudaMalloc(&pflags, pcells * sizeof(int)));
    B200_CUDA_OK(cudaMalloc(&pscan, pcells * sizeof(int)));
    B200_CUDA_OK(cudaMalloc(&st->pair_ptr, ((size_t)st->MB2 + 1) * sizeof(int)));
    tc_pair_flags_kernel<<<(unsigned)((pcells + 255) / 256), 256, 0, s>>>(flags, st->MB, st->MB2, st->KT, pflags);
    B200_LAUNCH_OK("tc_pair_flags_kernel");
    B200_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, pflags, pscan, (int)pcells, s));  // pcells <= cells
    int plast[2] = {0, 0};
    B200_CUDA_OK(cudaMemcpyAsync(&plast[0], pscan + pcells - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaMemcpyAsync(&plast[1], pflags + pcells - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    st->n_pair_tiles = plast[0] + plast[1];
    const size_t npt = std::max<size_t>((size_t)st->n_pair_tiles, 1);
    B200_CUDA_OK(cudaMalloc(&st->pair_kt, npt * sizeof(int)));
    B200_CUDA_OK(cudaMalloc(&st->pair_slot, 2 * npt * sizeof(int)));
    tc_pair_list_kernel<<<(unsigned)((pcells + 255) / 256), 256, 0, s>>>(flags, st->scan, pflags, pscan, st->MB, st->MB2,
                                                                        st->KT, st->n_tiles, (int)npt, st->pair_ptr,
                                                                        st->pair_kt, st->pair_slot);
    B200_LAUNCH_OK("tc_pair_list_kernel");
    B200_CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(pflags);
    cudaFree(pscan);
  }
  B200_CUDA_OK(cudaStreamSynchronize(s));
  cudaFree(flags);
  cudaFree(tmp);
  return B200_OK;
}

// values changed in place, pattern unchanged: rewrite the stored values (zeros stay zeros)
int tc_store_refresh(const int* row_ptr, const int* col, const float* val, int nnz, TcStore* st, cudaStream_t s) {
  return tc_launch_pack_a(row_ptr, col, val, nnz, st, s);
}

// Tile width: the Nt (multiple of 64, <= 256) that minimises waves x per-tile cost on `sms` persistent CTAs.
// Narrow tiles re-read A more often and are shared-memory-bound (A 4 KB + B Nt*32 B per MMA), so their
// per-column cost is higher: cost(Nt) ~ Nt + 64.
int tc_pick_nt(int MB, int N, int sms) {
  int best = 64;
  double best_cost = 1e300;
  for (int nt = 64; nt <= 256; nt += 64) {
    const long long tiles = (long long)MB * ((N + nt - 1) / nt);
    const long long waves = (tiles + sms - 1) / sms;
    const double cost = (double)waves * (nt + 64);
    if (cost < best_cost - 1e-9 || (fabs(cost - best_cost) <= 1e-9 && nt > best)) {
      best_cost = cost;
      best = nt;
    }
  }
  return best;
}

// SDDMM: tile width the operands are padded for.  When the contraction fits one TMEM accumulation the transposed
// kernel (tc_sddmm_short_kernel) walks the pattern in blocks of 128 columns, so the padding must be a multiple of 128.
int tc_pick_nt_sddmm(int MB, int K, int N, int kind, int sms) {
  const int nt = tc_pick_nt(MB, K, sms);
  const int KT = (N + bk_of(kind) - 1) / bk_of(kind);
  if (options().tc_sddmm_short != 0 && KT <= std::max(1, options().tc_chunk)) return nt <= 128 ? 128 : 256;
  return nt;
}

// workspace: [control words, 256 B: zero between launches][column maxima, one per padded column][packed B]
static size_t tc_colmax_bytes(int n_pad) { return (((size_t)n_pad * 4 + 255) / 256) * 256; }
size_t tc_pack_b_bytes(int K, int N, int nt, int kind) {
  const int bk = bk_of(kind);
  const size_t KT = (size_t)(K + bk - 1) / bk;
  const size_t NG = (size_t)((N + nt - 1) / nt) * (nt / 8);
  return 256 + tc_colmax_bytes((int)NG * 8) + KT * NG * ATOM;
}

// Pack one dense operand for the tiles: `rows` rows (the operand's non-contracted index, padded to NG groups of 8)
// by `depth` contraction values.  by_depth = true: memory is [depth][rows] (each contraction index is a row of
// memory, e.g. B [K, N] for the forward); false: memory is [rows][depth] (e.g. B^T, or G [M, N] for the SDDMM).
// fp16 kind: row maxima -> rmax (then the scaling), wide operands in one fused launch.
template <int KIND>
static int tc_pack_operand(const float* X, long long ld, int rows, int depth, bool by_depth, int KT, int NG,
                           unsigned char* pack, uint32_t* rmax, int* ctl, cudaStream_t s) {
  const int sms = device_sm_count();
  const int total = KT * NG;
  const int blocks = std::max(1, std::min((total + 7) / 8, sms * 8));
  if (KIND == KIND_F16 && NG >= sms) {
    if (!by_depth) tc_prepack_b_kernel<true><<<NG, 256, 0, s>>>(X, ld, depth, rows, KT, NG, pack, rmax, ctl);
    else tc_prepack_b_kernel<false><<<NG, 256, 0, s>>>(X, ld, depth, rows, KT, NG, pack, rmax, ctl);
    B200_LAUNCH_OK("tc_prepack_b_kernel");
    return B200_OK;
  }
  if (KIND == KIND_F16) {
    B200_CUDA_OK(cudaMemsetAsync(rmax, 0, tc_colmax_bytes(NG * 8), s));
    if (!by_depth) {
      tc_colmax_kernel<true><<<(unsigned)(((size_t)rows * 32 + 255) / 256), 256, 0, s>>>(X, ld, depth, rows, rmax, ctl);
    } else {
      dim3 grid((unsigned)((rows + 255) / 256),
                (unsigned)std::max(1, std::min(depth, (sms * 8) / std::max(1, (rows + 255) / 256))));
      tc_colmax_kernel<false><<<grid, 256, 0, s>>>(X, ld, depth, rows, rmax, ctl);
    }
    B200_LAUNCH_OK("tc_colmax_kernel");
  }
  if (!by_depth) tc_pack_b_kernel<KIND, true><<<blocks, 256, 0, s>>>(X, ld, depth, rows, KT, NG, pack, rmax, ctl);
  else tc_pack_b_kernel<KIND, false><<<blocks, 256, 0, s>>>(X, ld, depth, rows, KT, NG, pack, rmax, ctl);
  B200_LAUNCH_OK("tc_pack_b_kernel");
  return B200_OK;
}

// debug timeline read-back (tools/tc_trace.py)
static unsigned long long g_tc_trace_host[3 * TRACE_REGION];
int tc_trace_sync() {
  return cudaMemcpyFromSymbol(g_tc_trace_host, g_tc_trace, sizeof(g_tc_trace_host)) == cudaSuccess ? 0 : -1;
}
int tc_trace_clear() {
  memset(g_tc_trace_host, 0, sizeof(g_tc_trace_host));
  return cudaMemcpyToSymbol(g_tc_trace, g_tc_trace_host, sizeof(g_tc_trace_host)) == cudaSuccess ? 0 : -1;
}
long long tc_trace_word(int i) { return (i >= 0 && i < 3 * TRACE_REGION) ? (long long)g_tc_trace_host[i] : -1; }

int g_tc_last_kernel = 0;    // 1 = one CTA per tile, 2 = CTA pairs (diagnostics: b200_get_option("tc_last_kernel"))
int g_tc_pair_capacity = -1; // co-resident CTA pairs reported by the occupancy query

// co-resident CTA pairs of the pair kernel on this device (0 = cluster launch not possible)
template <int KIND, bool SDDMM>
static int tc_pair_capacity(size_t smem) {
  static int cached[64] = {0};
  static bool known[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (!known[dev & 63]) {
    int n = 0;
    if (cudaFuncSetAttribute(tc_spmm2_kernel<KIND, SDDMM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) ==
        cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * 74, 1, 1);
      cfg.blockDim = dim3(TC_THREADS, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, tc_spmm2_kernel<KIND, SDDMM>, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    cached[dev & 63] = n;
    known[dev & 63] = true;
  }
  return cached[dev & 63];
}

template <int KIND, bool SDDMM>
static int tc_run(TcParams& p, int tiles, cudaStream_t s) {
  p.chunk = std::max(1, options().tc_chunk);
  p.terms = options().tc_terms;
  if constexpr (SDDMM) {
    // short contraction: the 16-epilogue-warp kernel over 128 x 128 tiles (the packed operands are laid out in
    // groups of 8 rows, so the tile width is free to differ from the nt they were padded for)
    // (mode 1: 256-row tiles — fewer operand bytes per flop and per stored entry — unless the contraction is one
    // k-tile deep AND the pattern is dense, where the 128-row tiles' four TMEM buffers hide the longer store phase
    // better: measured 0.200 / 0.204 / 0.226 / 0.312 ms (256) vs 0.224 / 0.228 / 0.271 / 0.433 ms (128) at N = 16 / 64 /
    // 128 / 256 on 16384^2 at 90 % sparsity, 0.261 / 0.257 vs 0.233 / 0.237 ms at 50 % and N <= 64; 2 / 3 force a shape)
    const int mode = options().tc_sddmm_short;
    if (mode != 0 && p.KT <= p.chunk && (p.terms & ~7) == 0 && p.Nt % 128 == 0) {
      const bool wide = mode == 3 || (mode == 1 && p.MB >= 2 && (p.KT > 1 || !p.dense_pattern));
      p.NB = p.NB * (p.Nt / 128);            // blocks of 128 pattern columns
      p.Nt = wide ? 256 : 128;
      if (wide) p.MB = (p.MB + 1) / 2;       // blocks of 256 pattern rows (packed G is padded to an even block count)
      const int grid = std::max(1, std::min(device_sm_count(), p.MB * p.NB));
      g_tc_last_kernel = wide ? 4 : 3;
      static bool attr_short[2][64] = {{false}};
      int dev = 0;
      B200_CUDA_OK(cudaGetDevice(&dev));
      if (!attr_short[wide][dev & 63]) {
        if (wide) B200_CUDA_OK(cudaFuncSetAttribute(tc_sddmm_short_kernel<KIND, 256>,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ShCfg<256>::SMEM));
        else B200_CUDA_OK(cudaFuncSetAttribute(tc_sddmm_short_kernel<KIND, 128>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ShCfg<128>::SMEM));
        attr_short[wide][dev & 63] = true;
      }
      if (wide) tc_sddmm_short_kernel<KIND, 256><<<grid, SH_THREADS, ShCfg<256>::SMEM, s>>>(p);
      else tc_sddmm_short_kernel<KIND, 128><<<grid, SH_THREADS, ShCfg<128>::SMEM, s>>>(p);
      B200_LAUNCH_OK("tc_sddmm_short_kernel");
      return B200_OK;
    }
  }
  // CTA pairs (cta_group::2) when there are at least two row blocks: each CTA stages A + half of the B tile.
  // tc_pair = 2 (default, auto): only where they were measured faster — 256-wide tiles on a long contraction, where one
  // CTA's operand stream (96 KB per k-tile) runs into the L2 -> SM rate of its SM and only two stages fit: config 5
  // x 256 0.375 / 0.363 -> 0.339 / 0.331 ms at 50 / 90 % sparsity; narrower tiles and config 3's short K are slower
  // on pairs (profiles/r2_experiments.md).
  const int pair_mode = options().tc_pair;
  const bool want_pair = pair_mode == 1 || (pair_mode == 2 && !SDDMM && p.Nt == 256 && p.KT >= 128);
  if (want_pair && p.MB >= 2 && p.pair_ptr_ok) {
    const int stage2 = A_TILE + p.Nt * 128;
    p.stages = std::min(MAX_STAGES, (int)((TC_SMEM_MAX - 1024 - 256 - STG_BYTES) / stage2));
    const size_t smem2 = (size_t)p.stages * stage2 + 1024 + 256 + STG_BYTES;
    const int cap = tc_pair_capacity<KIND, SDDMM>(smem2);
    g_tc_pair_capacity = cap;
    if (cap > 0) {
      g_tc_last_kernel = 2;
      const int pairs = std::max(1, std::min(cap, p.MB2 * p.NB));
      tc_spmm2_kernel<KIND, SDDMM><<<2 * pairs, TC_THREADS, smem2, s>>>(p);
      B200_LAUNCH_OK(SDDMM ? "tc_sddmm2_kernel" : "tc_spmm2_kernel");
      return B200_OK;
    }
  }
  const int stage_bytes = A_TILE + p.Nt * 256;
  p.stages = std::min(MAX_STAGES, (int)((TC_SMEM_MAX - 1024 - 256 - STG_BYTES) / stage_bytes));
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256 + STG_BYTES;  // ring | barriers | transposers
  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    B200_CUDA_OK(cudaFuncSetAttribute(tc_spmm_kernel<KIND, SDDMM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024));
    attr_set[dev & 63] = true;
  }
  const int grid = std::max(1, std::min(device_sm_count(), tiles));
  g_tc_last_kernel = 1;
  tc_spmm_kernel<KIND, SDDMM><<<grid, TC_THREADS, smem, s>>>(p);
  B200_LAUNCH_OK(SDDMM ? "tc_sddmm_kernel" : "tc_spmm_kernel");
  return B200_OK;
}

template <int KIND>
static int tc_launch(const TcStore& st, const int* row_ptr, const int* col, const float* val, const float* B,
                     long long ldb, int N, float* C, long long ldc, const float* bias, unsigned flags, void* ws, int nt,
                     cudaStream_t s) {
  int* ctl = reinterpret_cast<int*>(ws);  // zero between launches (the kernel re-arms it on exit)
  const int NB = (N + nt - 1) / nt, NG = NB * (nt / 8);
  uint32_t* colmax = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(ws) + 256);
  unsigned char* b_pack = reinterpret_cast<unsigned char*>(ws) + 256 + tc_colmax_bytes(NG * 8);
  const bool bt = (flags & 1u) != 0;
  int rc = tc_pack_operand<KIND>(B, ldb, N, st.K, /*by_depth=*/!bt, st.KT, NG, b_pack, colmax, ctl, s);
  if (rc) return rc;
  TcParams p{};
  p.a_tiles = st.tiles; p.blk_ptr = st.blk_ptr; p.blk_kt = st.blk_kt;
  p.b_pack = b_pack; p.NG = NG;
  p.C = C; p.ldc = ldc; p.M = st.M; p.N = N; p.bias = bias; p.flags = flags;
  p.row_scale = st.row_scale; p.colmax = colmax;
  p.MB = st.MB; p.NB = NB; p.Nt = nt; p.KT = st.KT;  // (KT: only read by the kernel-shape choice in tc_run)
  p.pair_ptr = st.pair_ptr; p.pair_kt = st.pair_kt; p.pair_slot = st.pair_slot;
  p.n_pair_tiles = std::max(st.n_pair_tiles, 1); p.MB2 = st.MB2; p.pair_ptr_ok = st.pair_ptr != nullptr;
  p.ctl = ctl;
  p.row_ptr = row_ptr; p.col = col; p.val = val; p.B = B;
  p.b_sk = bt ? 1 : ldb; p.b_sn = bt ? ldb : 1;
  return tc_run<KIND, false>(p, st.MB * NB, s);
}

// C = A_tiles * B (+ bias, ReLU).  ws: the plan's packed-operand workspace, its first 256 bytes zeroed once when
// it was allocated.  When the packer finds Inf/NaN in B the same launch computes C in CSR order instead.
int launch_tc_forward(const TcStore& st, const int* row_ptr, const int* col, const float* val, const float* B,
                      long long ldb, int N, float* C, long long ldc, const float* bias, unsigned flags, void* ws,
                      size_t ws_bytes, int nt, cudaStream_t s) {
  if (st.M == 0 || N == 0) return B200_OK;
  if (nt <= 0) nt = tc_pick_nt(st.MB, N, device_sm_count());
  if (ws_bytes < tc_pack_b_bytes(st.K, N, nt, st.kind)) {
    set_error("block-dense path: workspace too small");
    return B200_ERR_INVALID;
  }
  return st.kind == KIND_F16
             ? tc_launch<KIND_F16>(st, row_ptr, col, val, B, ldb, N, C, ldc, bias, flags, ws, nt, s)
             : tc_launch<KIND_TF32>(st, row_ptr, col, val, B, ldb, N, C, ldc, bias, flags, ws, nt, s);
}

// ---- SDDMM on the tensor cores -------------------------------------------------------------------------------
namespace {
__global__ void tc_sample_mask_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int M,
                                      uint32_t* __restrict__ mask) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const int rs = __ldg(row_ptr + row), re = __ldg(row_ptr + row + 1);
  for (int p = rs + lane; p < re; p += 32) {
    const int c = __ldg(col + p);
    atomicOr(mask + (size_t)(c >> 5) * M + row, 1u << (c & 31));
  }
}
__global__ void tc_sample_off_kernel(const int* __restrict__ row_ptr, const uint32_t* __restrict__ mask, int M,
                                     int groups, int* __restrict__ off) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  int q = __ldg(row_ptr + row);
  for (int g = 0; g < groups; ++g) {  // [group][row]: coalesced across the rows of a warp
    off[(size_t)g * M + row] = q;
    q += __popc(mask[(size_t)g * M + row]);
  }
}
}  // namespace

void tc_sample_free(TcSample* sm) {
  cudaFree(sm->mask); cudaFree(sm->off);
  *sm = TcSample{};
}

// Pattern of a CSR matrix as per-(32-column group, row) bit masks + value offsets (rows sorted, no duplicates).
int tc_sample_build(const int* row_ptr, const int* col, int M, int K, int nnz, TcSample* sm, cudaStream_t s) {
  *sm = TcSample{};
  sm->M = M; sm->K = K; sm->nnz = nnz;
  sm->groups = (K + 31) / 32;
  const size_t cells = (size_t)sm->groups * (size_t)std::max(M, 1);
  if (cells * 8 > ((size_t)4 << 30)) {
    set_error("tensor-core SDDMM: pattern masks would take %zu bytes", cells * 8);
    return B200_ERR_ALLOC;
  }
  B200_CUDA_OK(cudaMalloc(&sm->mask, cells * sizeof(uint32_t)));
  B200_CUDA_OK(cudaMalloc(&sm->off, cells * sizeof(int)));
  B200_CUDA_OK(cudaMemsetAsync(sm->mask, 0, cells * sizeof(uint32_t), s));
  if (M > 0 && nnz > 0) {
    tc_sample_mask_kernel<<<(unsigned)(((size_t)M * 32 + 255) / 256), 256, 0, s>>>(row_ptr, col, M, sm->mask);
    B200_LAUNCH_OK("tc_sample_mask_kernel");
  }
  if (M > 0) {
    tc_sample_off_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(row_ptr, sm->mask, M, sm->groups, sm->off);
    B200_LAUNCH_OK("tc_sample_off_kernel");
  }
  return B200_OK;
}

// workspace of the SDDMM: [control 256 B][row maxima of G][row maxima of B][packed G][packed B]
size_t tc_sddmm_ws_bytes(int M, int K, int N, int nt, int kind) {
  const size_t KT = (size_t)(N + bk_of(kind) - 1) / bk_of(kind);
  const size_t NGa = (size_t)(((M + BM - 1) / BM + 1) / 2 * 2) * (BM / 8);  // an even number of 128-row blocks (CTA pairs)
  const size_t NGb = (size_t)((K + nt - 1) / nt) * (nt / 8);
  return 256 + tc_colmax_bytes((int)NGa * 8) + tc_colmax_bytes((int)NGb * 8) + KT * (NGa + NGb) * ATOM;
}

template <int KIND>
static int tc_sddmm(const TcSample& sm, const int* row_ptr, const int* col, const float* G, long long ldg,
                    const float* B, long long ldb, int N, bool transposed, float* out, void* ws, int nt,
                    cudaStream_t s) {
  const int M = sm.M, K = sm.K;
  const int KT = (N + bk_of(KIND) - 1) / bk_of(KIND);
  const int MB = (M + BM - 1) / BM, MB2 = (MB + 1) / 2, NGa = 2 * MB2 * (BM / 8);
  const int NB = (K + nt - 1) / nt, NGb = NB * (nt / 8);
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  int* ctl = reinterpret_cast<int*>(base);
  uint32_t* amax = reinterpret_cast<uint32_t*>(base + 256);
  uint32_t* bmax = reinterpret_cast<uint32_t*>(base + 256 + tc_colmax_bytes(NGa * 8));
  unsigned char* a_pack = base + 256 + tc_colmax_bytes(NGa * 8) + tc_colmax_bytes(NGb * 8);
  unsigned char* b_pack = a_pack + (size_t)KT * NGa * ATOM;
  // contraction over the dense column index n: G [M, N] / B [K, N] are [rows][depth]; the transposed forms
  // (dy [N, M], x [N, K] of an FC layer) are [depth][rows]
  int rc = tc_pack_operand<KIND>(G, ldg, M, N, /*by_depth=*/transposed, KT, NGa, a_pack, amax, ctl, s);
  if (rc) return rc;
  rc = tc_pack_operand<KIND>(B, ldb, K, N, /*by_depth=*/transposed, KT, NGb, b_pack, bmax, ctl, s);
  if (rc) return rc;
  TcParams p{};
  p.a_tiles = a_pack; p.NGa = NGa; p.KT = KT;
  p.b_pack = b_pack; p.NG = NGb;
  p.C = out; p.M = M; p.N = N;
  p.amax = amax; p.colmax = bmax;
  p.samp_mask = sm.mask; p.samp_off = sm.off; p.groups = sm.groups;
  p.MB = MB; p.NB = NB; p.Nt = nt;
  p.MB2 = MB2; p.pair_ptr_ok = 1;
  p.dense_pattern = (double)sm.nnz >= 0.3 * (double)M * (double)K ? 1 : 0;
  p.ctl = ctl;
  p.row_ptr = row_ptr; p.col = col;
  p.G = G; p.g_sm = transposed ? 1 : ldg; p.g_sn = transposed ? ldg : 1;
  p.B = B; p.b_sk = transposed ? 1 : ldb; p.b_sn = transposed ? ldb : 1;
  return tc_run<KIND, true>(p, MB * NB, s);
}

// out[p] = sum_n G[row(p), n] * B[col(p), n] on the tensor cores (D = G * B^T sampled on the pattern).
// transposed: G and B are given as [N, M] / [N, K] (gradient and activations of an FC layer as they are).
int launch_tc_sddmm(const TcSample& sm, const int* row_ptr, const int* col, const float* G, long long ldg,
                    const float* B, long long ldb, int N, bool transposed, float* out, void* ws, size_t ws_bytes,
                    int nt, int kind, cudaStream_t s) {
  if (sm.M == 0 || sm.K == 0) return B200_OK;
  if (ws_bytes < tc_sddmm_ws_bytes(sm.M, sm.K, N, nt, kind)) {
    set_error("tensor-core SDDMM: workspace too small");
    return B200_ERR_INVALID;
  }
  return kind == KIND_F16 ? tc_sddmm<KIND_F16>(sm, row_ptr, col, G, ldg, B, ldb, N, transposed, out, ws, nt, s)
                          : tc_sddmm<KIND_TF32>(sm, row_ptr, col, G, ldg, B, ldb, N, transposed, out, ws, nt, s);
}

}  // namespace b200
