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

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace b200 {

namespace {

constexpr int BM = 128;           // rows of A per tile = UMMA M
constexpr int BK = 32;            // tf32 per 128-byte swizzle row = k extent of a stage
constexpr int ATOM = 2048;        // bytes of one 8-row group: hi atom (1 KB) then lo atom (1 KB)
constexpr int A_TILE = (BM / 8) * ATOM;  // 32 KB
constexpr int TC_THREADS = 192;
constexpr int MAX_STAGES = 4;

// byte offset of element (r, c) inside a 1 KB K-major SWIZZLE_128B atom (r < 8 rows of 128 B, c < 32 floats)
__host__ __device__ __forceinline__ int sw128(int r, int c) {
  return r * 128 + ((((c >> 2) ^ r) & 7) << 4) + ((c & 3) << 2);
}

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u & 0xffffe000u);
}

// ---- plan-time packing of A --------------------------------------------------------------------------------
__global__ void tc_mark_tiles_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int M, int KT,
                                     int* __restrict__ flags) {
  // one warp per row: mark the k-tiles this row touches
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int mb = row / BM;
  const int rs = __ldg(row_ptr + row), re = __ldg(row_ptr + row + 1);
  for (int p = rs + lane; p < re; p += 32) flags[(size_t)mb * KT + (__ldg(col + p) / BK)] = 1;
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

__global__ void tc_pack_a_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col,
                                 const float* __restrict__ val, int M, int KT, const int* __restrict__ scan,
                                 unsigned char* __restrict__ tiles) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int mb = row / BM, r = row % BM;
  const int rs = __ldg(row_ptr + row), re = __ldg(row_ptr + row + 1);
  for (int p = rs + lane; p < re; p += 32) {
    const int c = __ldg(col + p);
    const float x = __ldg(val + p);
    const int slot = scan[(size_t)mb * KT + c / BK];
    unsigned char* atom = tiles + (size_t)slot * A_TILE + (size_t)(r >> 3) * ATOM;
    const int off = sw128(r & 7, c % BK);
    const float hi = tf32_rn(x);
    *reinterpret_cast<float*>(atom + off) = hi;
    *reinterpret_cast<float*>(atom + 1024 + off) = tf32_rn(x - hi);
  }
}

// ---- per-call packing of the dense operand ----------------------------------------------------------------
// Packed layout: [kt][n group of 8][hi atom | lo atom]; atom row = one n, its 32 k values (128 B, swizzled).
// TRANS = false: B is [K, N] row-major (the reference's layout, src/custom_mm.cpp:203-217): a warp reads, for
//   each of the 32 k rows of the tile, the 32-byte sector holding 8 consecutive n (lane = k) and writes full
//   128-byte atom rows (the 32 lanes of a store cover one row).
// TRANS = true: the operand is given as B^T [N, K] row-major (activations [tokens, features] of an FC layer):
//   atom rows are contiguous in memory, float4 in, float4 out.
// Rows k >= K and columns n >= N are written as zeros.  Non-finite inputs raise *nonfinite (the dense
// contraction would turn 0 * Inf into NaN where the sparse product skips the term): the caller falls back.
template <bool TRANS>
__global__ void __launch_bounds__(256)
tc_pack_b_kernel(const float* __restrict__ B, long long ld, int K, int N, int KT, int NG,
                 unsigned char* __restrict__ out, int* __restrict__ nonfinite) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = KT * NG;
  bool bad = false;
  for (int w = warp; w < total; w += (gridDim.x * blockDim.x) >> 5) {
    const int kt = w / NG, g = w % NG;
    unsigned char* atom = out + ((size_t)kt * NG + g) * ATOM;
    const int n0 = g * 8, k0 = kt * BK;
    if constexpr (!TRANS) {
      const int k = k0 + lane;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = 0.f;
      if (k < K) {
        const float* src = B + (long long)k * ld + n0;
        if (n0 + 8 <= N && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src));
          const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
          x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + j < N) x[j] = __ldg(src + j);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bad |= !(fabsf(x[j]) <= 3.0e38f);
        const float hi = tf32_rn(x[j]);
        const int off = sw128(j, lane);
        *reinterpret_cast<float*>(atom + off) = hi;
        *reinterpret_cast<float*>(atom + 1024 + off) = tf32_rn(x[j] - hi);
      }
    } else {
      // lane -> (row rr = lane / 4, 16-byte chunks q = lane % 4 and q + 4)
      const int rr = lane >> 2, n = n0 + rr;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = (lane & 3) + 4 * h;
        const int k = k0 + 4 * q;
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < N) {
          const float* src = B + (long long)n * ld + k;
          if (k + 4 <= K && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src));
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (k + j < K) x[j] = __ldg(src + j);
          }
        }
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bad |= !(fabsf(x[j]) <= 3.0e38f);
          hi[j] = tf32_rn(x[j]);
          lo[j] = tf32_rn(x[j] - hi[j]);
        }
        const int off = rr * 128 + (((q ^ rr) & 7) << 4);
        *reinterpret_cast<float4*>(atom + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(atom + 1024 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
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
// D[tmem] (+)= A[smem] * B[smem]^T, both K-major, one elected thread issues.  SASS: UTCHMMA (tf32 flavour).
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed (implies fence::before)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive columns of 32-bit TMEM -> 32 registers per thread (thread = lane = row).  SASS: LDTM
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups `sbo` bytes apart (sm_100 version bit)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;                        // leading byte offset: unused for swizzled K-major (canonical 1)
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;   // stride byte offset
  d |= (uint64_t)1 << 46;                        // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::tf32: D fp32, A/B tf32, both K-major, M = 128, N = nt
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32(int nt) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
  unsigned flags;
  int MB, NB, Nt;
  int stages;
  int split_acc;   // 1: correction terms accumulate in their own TMEM columns, added in the epilogue
  int terms;       // bit 0 hi*hi, bit 1 hi*lo, bit 2 lo*hi (debug / accuracy studies; 7 = all)
  const int* skip_if;  // the packer's non-finite flag: when set the kernel leaves C to the gather path
  int* dbg;        // [0] = first wait that timed out (0 = none)
};

__global__ void __launch_bounds__(TC_THREADS, 1) tc_spmm_kernel(const TcParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  if (p.skip_if != nullptr && *reinterpret_cast<const volatile int*>(p.skip_if) != 0) return;

  const uint32_t raw = smem_u32(tc_smem_raw);
  unsigned char* smem = tc_smem_raw + (((raw + 1023u) & ~1023u) - raw);  // swizzle atoms need 1 KB alignment
  const int stage_bytes = A_TILE + p.Nt * 256;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int acc_cols = p.Nt * (p.split_acc ? 2 : 1);
  const int acc_stages = (2 * acc_cols <= 512) ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, 4);  // one arrive per epilogue warp
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
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
        const int mb = tile / p.NB, nb = tile % p.NB;
        const int t0 = __ldg(p.blk_ptr + mb), t1 = __ldg(p.blk_ptr + mb + 1);
        for (int t = t0; t < t1; ++t) {
          ok = mbar_wait_bounded(empty + stage, phase ^ 1u, p.dbg, 100 + stage);
          if (!ok) break;
          const int kt = __ldg(p.blk_kt + t);
          unsigned char* sa = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(full + stage, (uint32_t)stage_bytes);
          tma_bulk_g2s(sa, p.a_tiles + (size_t)t * A_TILE, A_TILE, full + stage);
          tma_bulk_g2s(sa + A_TILE, p.b_pack + ((size_t)kt * p.NG + (size_t)nb * (p.Nt / 8)) * ATOM,
                       (uint32_t)p.Nt * 256u, full + stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one lane issues every tcgen05.mma of this CTA =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(p.Nt);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
        const int mb = tile / p.NB;
        const int as = it % acc_stages;
        const uint32_t aphase = (uint32_t)(it / acc_stages) & 1u;
        ok = mbar_wait_bounded(acc_empty + as, aphase ^ 1u, p.dbg, 200 + as);
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(as * acc_cols);
        const uint32_t d_corr = p.split_acc ? d_main + (uint32_t)p.Nt : d_main;
        const int t0 = __ldg(p.blk_ptr + mb), t1 = __ldg(p.blk_ptr + mb + 1);
        uint32_t acc_main = 0, acc_corr = 0;
        for (int t = t0; t < t1; ++t) {
          ok = mbar_wait_bounded(full + stage, phase, p.dbg, 300 + stage);
          if (!ok) break;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + A_TILE;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {  // UMMA K = 8 tf32 = 32 bytes along the swizzled row
            const uint64_t a_hi = umma_desc(sa + k * 32, ATOM), a_lo = umma_desc(sa + 1024 + k * 32, ATOM);
            const uint64_t b_hi = umma_desc(sb + k * 32, ATOM), b_lo = umma_desc(sb + 1024 + k * 32, ATOM);
            uint32_t& acc_c = p.split_acc ? acc_corr : acc_main;  // the enable-input-D bit of the target columns
            if (p.terms & 2) { umma_tf32(d_corr, a_hi, b_lo, idesc, acc_c); acc_c = 1; }
            if (p.terms & 4) { umma_tf32(d_corr, a_lo, b_hi, idesc, acc_c); acc_c = 1; }
            if (p.terms & 1) { umma_tf32(d_main, a_hi, b_hi, idesc, acc_main); acc_main = 1; }
          }
          umma_commit(empty + stage);  // the smem slot is free once these MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(acc_full + as);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM -> registers -> (+bias, ReLU) -> C =====
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const bool relu = (p.flags & 4u) != 0, ct = (p.flags & 2u) != 0;
    const bool vec4 = !ct && (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    int it = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
      const int mb = tile / p.NB, nb = tile % p.NB;
      const int as = it % acc_stages;
      const uint32_t aphase = (uint32_t)(it / acc_stages) & 1u;
      const bool empty_block = __ldg(p.blk_ptr + mb) == __ldg(p.blk_ptr + mb + 1);
      ok = mbar_wait_bounded(acc_full + as, aphase, p.dbg, 400 + as);
      if (!ok) break;
      tc_fence_after();
      const int m = mb * BM + q * 32 + lane;
      const float bias = (p.bias != nullptr && m < p.M) ? __ldg(p.bias + m) : 0.f;
      const uint32_t t_main = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * acc_cols);
      for (int c0 = 0; c0 < p.Nt; c0 += 32) {
        const int n0 = nb * p.Nt + c0;
        if (n0 >= p.N) break;
        float v[32];
        if (empty_block) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        } else {
          tmem_ld32(t_main + (uint32_t)c0, v);
          if (p.split_acc) {
            float w[32];
            tmem_ld32(t_main + (uint32_t)(p.Nt + c0), w);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += w[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] += bias;
          if (relu) v[j] = fmaxf(v[j], 0.f);
        }
        if (ct) {  // C^T [N, M]: a register is one n, the 32 lanes are 32 consecutive m -> coalesced
          if (m < p.M) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) p.C[(long long)(n0 + j) * p.ldc + m] = v[j];
          }
        } else if (m < p.M) {
          float* dst = p.C + (long long)m * p.ldc + n0;
          if (vec4 && n0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              __stcs(reinterpret_cast<float4*>(dst + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) dst[j] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + as);  // this warp's quarter of the accumulator is drained
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Generic strided CSR SpMM: warp per row, lanes over 32 dense columns, any element strides for B and C, bias
// and ReLU.  Slow and simple on purpose: it is what runs when the tensor-core path has to stand down (Inf/NaN in
// the dense operand) or when transposed operands are requested for a matrix that is too sparse for it.
__global__ void __launch_bounds__(256)
spmm_strided_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const float* __restrict__ val,
                    int M, int N, const float* __restrict__ B, long long b_sk, long long b_sn, float* __restrict__ C,
                    long long c_sm, long long c_sn, const float* __restrict__ bias, int relu,
                    const int* __restrict__ run_if) {
  if (run_if != nullptr && *reinterpret_cast<const volatile int*>(run_if) == 0) return;
  const int lane = threadIdx.x & 31;
  const int slabs = (N + 31) / 32;
  const long long total = (long long)M * slabs;
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total;
       w += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int row = (int)(w / slabs), n = (int)(w % slabs) * 32 + lane;
    const int rs = __ldg(row_ptr + row), re = __ldg(row_ptr + row + 1);
    float acc = 0.f;
    if (n < N) {
      for (int p = rs; p < re; ++p)
        acc = fmaf(__ldg(val + p), __ldg(B + (long long)__ldg(col + p) * b_sk + (long long)n * b_sn), acc);
      if (bias != nullptr) acc += __ldg(bias + row);
      if (relu) acc = fmaxf(acc, 0.f);
      C[(long long)row * c_sm + (long long)n * c_sn] = acc;
    }
  }
}

}  // namespace

int launch_spmm_strided(const int* row_ptr, const int* col, const float* val, int M, int N, const float* B,
                        long long b_sk, long long b_sn, float* C, long long c_sm, long long c_sn,
                        const float* bias, bool relu, const int* run_if, cudaStream_t s) {
  if (M == 0 || N == 0) return B200_OK;
  const long long warps = (long long)M * ((N + 31) / 32);
  const int blocks = (int)std::min<long long>((warps + 7) / 8, (long long)device_sm_count() * 16);
  spmm_strided_kernel<<<blocks, 256, 0, s>>>(row_ptr, col, val, M, N, B, b_sk, b_sn, C, c_sm, c_sn, bias,
                                              relu ? 1 : 0, run_if);
  B200_LAUNCH_OK("spmm_strided_kernel");
  return B200_OK;
}

// ---- host side ------------------------------------------------------------------------------------------------

void tc_store_free(TcStore* st) {
  cudaFree(st->blk_ptr); cudaFree(st->blk_kt); cudaFree(st->scan); cudaFree(st->tiles);
  *st = TcStore{};
}

// Densify the non-empty 128 x 32 tiles of a CSR matrix into hi/lo swizzled tiles.  Synchronises the stream once
// (the number of non-empty tiles sizes the allocation).
int tc_store_build(const int* row_ptr, const int* col, const float* val, int M, int K, int nnz, TcStore* st,
                   cudaStream_t s) {
  *st = TcStore{};
  st->M = M; st->K = K;
  st->MB = (M + BM - 1) / BM;
  st->KT = (K + BK - 1) / BK;
  const size_t cells = (size_t)st->MB * st->KT;
  if (cells == 0 || cells > ((size_t)1 << 26)) {
    set_error("block-dense store: %zu candidate tiles is outside the supported range", cells);
    return B200_ERR_INVALID;
  }
  int* flags = nullptr;
  B200_CUDA_OK(cudaMalloc(&flags, cells * sizeof(int)));
  B200_CUDA_OK(cudaMalloc(&st->scan, cells * sizeof(int)));
  B200_CUDA_OK(cudaMalloc(&st->blk_ptr, ((size_t)st->MB + 1) * sizeof(int)));
  B200_CUDA_OK(cudaMemsetAsync(flags, 0, cells * sizeof(int), s));
  const unsigned wblocks = (unsigned)(((size_t)M * 32 + 255) / 256);
  if (nnz > 0) {
    tc_mark_tiles_kernel<<<wblocks, 256, 0, s>>>(row_ptr, col, M, st->KT, flags);
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
  const size_t bytes = std::max<size_t>((size_t)st->n_tiles, 1) * A_TILE;
  if (cudaMalloc(&st->tiles, bytes) != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMalloc of %zu bytes for the block-dense tiles failed", bytes);
    cudaFree(flags); cudaFree(tmp);
    tc_store_free(st);
    return B200_ERR_ALLOC;
  }
  st->tile_bytes = bytes;
  B200_CUDA_OK(cudaMemsetAsync(st->tiles, 0, bytes, s));
  if (nnz > 0) {
    tc_pack_a_kernel<<<wblocks, 256, 0, s>>>(row_ptr, col, val, M, st->KT, st->scan, st->tiles);
    B200_LAUNCH_OK("tc_pack_a_kernel");
  }
  B200_CUDA_OK(cudaStreamSynchronize(s));
  cudaFree(flags);
  cudaFree(tmp);
  return B200_OK;
}

// values changed in place, pattern unchanged: rewrite the stored values (zeros stay zeros)
int tc_store_refresh(const int* row_ptr, const int* col, const float* val, int nnz, TcStore* st, cudaStream_t s) {
  if (nnz == 0) return B200_OK;
  const unsigned wblocks = (unsigned)(((size_t)st->M * 32 + 255) / 256);
  tc_pack_a_kernel<<<wblocks, 256, 0, s>>>(row_ptr, col, val, st->M, st->KT, st->scan, st->tiles);
  B200_LAUNCH_OK("tc_pack_a_kernel");
  return B200_OK;
}

// Tile width: the Nt (multiple of 32, <= 256) that minimises waves x per-tile cost on `sms` persistent CTAs.
// Narrow tiles re-read A more often and are shared-memory-bound (A 4 KB + B Nt*32 B per MMA), so their
// per-column cost is higher: cost(Nt) ~ Nt + 64.
int tc_pick_nt(int MB, int N, int sms) {
  int best = 32;
  double best_cost = 1e300;
  for (int nt = 32; nt <= 256; nt += 32) {
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

size_t tc_pack_b_bytes(int K, int N, int nt) {
  const size_t KT = (size_t)(K + BK - 1) / BK;
  const size_t NG = (size_t)((N + nt - 1) / nt) * (nt / 8);
  return KT * NG * ATOM + 256;  // + the non-finite flag and the debug word
}

// C = A_tiles * B (+ bias, ReLU).  ws holds [flag 4 B | dbg 4 B | pad to 256][packed B].  The kernel is skipped
// on the device when B holds Inf/NaN (flag set by the packer): the caller then runs the gather path
// conditionally on the same flag.
int launch_tc_forward(const TcStore& st, const float* B, long long ldb, int N, float* C, long long ldc,
                      const float* bias, unsigned flags, void* ws, size_t ws_bytes, int nt, cudaStream_t s) {
  if (st.M == 0 || N == 0) return B200_OK;
  const int sms = device_sm_count();
  if (nt <= 0) nt = tc_pick_nt(st.MB, N, sms);
  const int terms_opt = options().tc_terms, split_opt = options().tc_split_acc;
  if (ws_bytes < tc_pack_b_bytes(st.K, N, nt)) {
    set_error("block-dense path: workspace too small");
    return B200_ERR_INVALID;
  }
  int* flag = reinterpret_cast<int*>(ws);
  unsigned char* b_pack = reinterpret_cast<unsigned char*>(ws) + 256;
  const int NB = (N + nt - 1) / nt, NG = NB * (nt / 8);
  B200_CUDA_OK(cudaMemsetAsync(flag, 0, 8, s));
  {
    const int total = st.KT * NG;
    const int blocks = std::min((total + 7) / 8, sms * 8);
    if (flags & 1u) tc_pack_b_kernel<true><<<blocks, 256, 0, s>>>(B, ldb, st.K, N, st.KT, NG, b_pack, flag);
    else tc_pack_b_kernel<false><<<blocks, 256, 0, s>>>(B, ldb, st.K, N, st.KT, NG, b_pack, flag);
    B200_LAUNCH_OK("tc_pack_b_kernel");
  }
  TcParams p{};
  p.a_tiles = st.tiles; p.blk_ptr = st.blk_ptr; p.blk_kt = st.blk_kt;
  p.b_pack = b_pack; p.NG = NG;
  p.C = C; p.ldc = ldc; p.M = st.M; p.N = N; p.bias = bias; p.flags = flags;
  p.MB = st.MB; p.NB = NB; p.Nt = nt;
  p.split_acc = split_opt;
  p.terms = terms_opt;
  p.skip_if = flag; p.dbg = flag + 1;
  const int stage_bytes = A_TILE + nt * 256;
  p.stages = std::min(MAX_STAGES, (int)((220 * 1024) / stage_bytes));
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    B200_CUDA_OK(cudaFuncSetAttribute(tc_spmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev & 63] = true;
  }
  const int grid = std::min(sms, st.MB * NB);
  tc_spmm_kernel<<<grid, TC_THREADS, smem, s>>>(p);
  B200_LAUNCH_OK("tc_spmm_kernel");
  return B200_OK;
}

}  // namespace b200
