// common.cuh — shared helpers of the sm_100a SpMM engine: status handling, PTX
// wrappers (mbarrier + TMA bulk copy, cache-hinted loads/stores) and the
// merge-path task boundary that every kernel and the host planner agree on.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/spmm_b200.h"

namespace b200 {

// ---- status ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define B200_CUDA_OK(expr)                                                            \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                        __FILE__, __LINE__);                                          \
      return B200_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

#define B200_REQUIRE(cond, msg)                                                       \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ::b200::set_error("invalid argument: %s (%s)", msg, #cond);                     \
      return B200_ERR_INVALID;                                                        \
    }                                                                                 \
  } while (0)

#define B200_LAUNCH_OK(name)                                                          \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ::b200::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
      return B200_ERR_CUDA;                                                           \
    }                                                                                 \
    ::b200::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
  } while (0)

// process-wide options (capi.cu)
struct Options {
  int stage_mode = 1;      // 1: TMA bulk staging of col/val, 0: direct global loads
  int items_per_task = 0;  // 0 = auto
  int flat_min_n = 24;     // smallest N served by the flat (one lane group) kernel
  int pipeline = 0;        // flat kernel variant: 0 = shipped; 1 = 8 warps/CTA; 2, 3 = 16 / 24 gathers per warp
                           // (the alternatives of profiles/r1_variant_sweep.md, kept for re-measurement)
  int slab_mode = 1;       // narrow-N shared-memory slab kernel: 0 = never, 1 = when slab_wanted() says so (sliced stream, v2), 2 = same with the first kernel (CSR walked in place), 3 = v2 whenever the shape allows it (probing)
  int host_copy_mode = 0;  // b200_plan_forward_host panel copies: 0 = copy engine (cudaMemcpy2DAsync), 1 = SM copy kernel
  int host_copy_ctas = 32; // CTAs of one SM-driven panel copy
  int host_copy_split = 1; // row chunks (= copy engines) per panel copy of the host pipeline
  int tail_tasks = 0;      // short tasks (W/4) at the end of a forward launch; measured no gain on config 4 and its shards (profiles/r2_scaling.md): off
  int d2c_warps = 0;       // dense->CSR one-pass: count warps per CTA (16, 20, 24; 0 = default 24), the other warps: offset + fill
  int d2c_lag = 0;         // dense->CSR one-pass experiments: low byte 1 = small pair rings; bit 8 = count side alone (wrong results)
  // block-dense tensor-core path (tc_gemm.cu)
  int tc_mode = 1;         // 0 = never, 1 = auto (density / N rule), 2 = whenever the shape allows it
  int tc_nt = 0;           // tile width override (multiple of 32, <= 256); 0 = auto
  int tc_terms = 7;        // which split-precision products are issued: 1 hi*hi, 2 hi*lo, 4 lo*hi
  int tc_chunk = 8;        // k-tiles accumulated in TMEM between two fp32 register accumulations (<= 96 truncating adds)
  int tc_pair = 2;         // CTA pairs (cta_group::2: tc_spmm2_kernel): 0 = never, 1 = whenever the shape allows it, 2 = auto (256-wide tiles on a long contraction: measured 10 % faster there, slower elsewhere; profiles/r2_experiments.md)
  int tc_sddmm_short = 1;  // tensor-core SDDMM whose contraction fits one TMEM accumulation: 1 = the 16-epilogue-warp kernel (tc_sddmm_short_kernel), 0 = the general kernel
  int tc_kind = 1;         // 0 = tf32 hi/lo split, 1 = fp16 hi/lo split with row / column scaling (default)
  int tc_min_density_pct = 2;  // auto rule: never below this density (percent of M*K); above it a cost model decides
};
Options& options();

int device_sm_count();

// ---- merge-path boundary ----------------------------------------------------------
// Items are consumed in the order: non-zeros of row 0, row-end 0, non-zeros of
// row 1, row-end 1, ...  Item index of row-end r is row_ptr[r+1] + r.
// boundary(d) = (i, j): rows < i are complete, j non-zeros are consumed.
// If the boundary falls inside row i and that row is "light" (length <= heavy),
// it is snapped down to the row start so light rows are never split.
struct Boundary {
  int row;
  int nz;
};

__host__ __device__ __forceinline__ Boundary merge_path_boundary(
    const int* __restrict__ row_ptr, int M, int nnz, long long d, int heavy) {
  long long total = (long long)M + (long long)nnz;
  if (d >= total) return Boundary{M, nnz};
  if (d <= 0) return Boundary{0, 0};
  // smallest i in [0, M) with row_ptr[i+1] + i >= d   (exists because d < total)
  int lo = 0, hi = M - 1;
  while (lo < hi) {
    int mid = (int)(((long long)lo + hi) >> 1);
#if defined(__CUDA_ARCH__)
    long long g = (long long)__ldg(row_ptr + mid + 1) + mid;
#else
    long long g = (long long)row_ptr[mid + 1] + mid;
#endif
    if (g >= d) hi = mid; else lo = mid + 1;
  }
  int i = lo;
  int j = (int)(d - i);
#if defined(__CUDA_ARCH__)
  int rs = __ldg(row_ptr + i), re = __ldg(row_ptr + i + 1);
#else
  int rs = row_ptr[i], re = row_ptr[i + 1];
#endif
  // all non-zeros of row i consumed, only its row-end item left: the row is complete.  (A row's last
  // non-zero and its row end are therefore always in the same task.)
  if (j >= re) return Boundary{i + 1, re};
  if (j > rs && (re - rs) <= heavy) j = rs;
  return Boundary{i, j};
}

#if defined(__CUDACC__)
// ---- PTX wrappers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.  SASS: UBLKCP.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// multimem.st: one store to a multicast address, replicated by the switch (NVLS)
__device__ __forceinline__ void multimem_st(float* p, float a, float b, float c, float d) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void multimem_st(float* p, float a, float b) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void multimem_st(float* p, float a) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

// Same bulk copy with an L2 evict-first hint: the col/val stream is read exactly once, it should not push the
// re-used B rows out of L2.
__device__ __forceinline__ void tma_bulk_g2s_stream(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                                    uint64_t* bar) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// read-only 128/64/32-bit loads of B (gathered rows; keep in L1/L2)
template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void load_cg(const float* p) {  // L2 only: data written by other SMs in this launch
    float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store_stream(float* p) const {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ __forceinline__ void store_mc(float* p) const { multimem_st(p, v[0], v[1], v[2], v[3]); }
};
template <>
struct Vec<2> {
  float v[2];
  __device__ __forceinline__ void load(const float* p) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  }
  __device__ __forceinline__ void load_cg(const float* p) {
    float2 t = __ldcg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  }
  __device__ __forceinline__ void store_stream(float* p) const {
    __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
  __device__ __forceinline__ void store_mc(float* p) const { multimem_st(p, v[0], v[1]); }
};
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void load_cg(const float* p) { v[0] = __ldcg(p); }
  __device__ __forceinline__ void store_stream(float* p) const { __stcs(p, v[0]); }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
  __device__ __forceinline__ void store_mc(float* p) const { multimem_st(p, v[0]); }
};
#endif  // __CUDACC__

// ---- internal launchers (one per .cu) -------------------------------------------------
struct FwdParams {
  const int* row_ptr;
  const int* col;
  const float* val;
  int M;
  int nnz;
  const float* B;
  long long ldb;
  int N;
  float* C;
  long long ldc;
  float* pieces;   // workspace: [2 * n_tasks][N] partial rows of split heavy rows
  int* counters;   // [n_tasks * column slabs] arrival counters of the pieces, zero between launches; null = the
                   // launcher zeroes a set inside the workspace
  int stage;       // 1 / 0 force TMA staging on / off, -1 = process option
  const float* bias;  // optional [M]: added to every column of the finished row
  int relu;           // max(x, 0) on the finished row
  // two task sizes: the first t1 tasks hold W items each, the rest wf <= W items (the tail of the launch drains in
  // short tasks; set from FwdPick by the launcher)
  int t1, wf;
};
// first merge-path item of task t / task that holds item i, for W-item tasks followed by wf-item tasks
__host__ __device__ __forceinline__ long long task_first_item(long long t, int W, int t1, int wf) {
  return t <= t1 ? t * W : (long long)t1 * W + (t - t1) * wf;
}
__host__ __device__ __forceinline__ long long task_of_item(long long i, int W, int t1, int wf) {
  const long long coarse = (long long)t1 * W;
  return i < coarse ? i / W : t1 + (i - coarse) / wf;
}
// multi-GPU epilogue (row-sharded A): every finished row of C is also stored into the same row block of
// the other ranks' buffers — plain P2P stores to mapped peer pointers, or one multimem.st to an NVLS
// multicast address that the NVSwitch replicates to every rank (including this one).  Kept out of
// FwdParams so that the single-GPU kernels carry none of it.
struct PeerSet {
  float* peer[7];
  int n_peers;
  float* mc;
};
template <bool PEERS>
struct PeerArg;
template <>
struct PeerArg<true> {
  const PeerSet* s;  // in device memory: the kernel's constant bank stays below 1 KB (measurably faster)
};
template <>
struct PeerArg<false> {
  char unused;
};
// lane layout + task size chosen for a problem (spmm_fwd.cu)
struct FwdPick {
  int vec, lpr, W, heavy, n_tasks, slabs;
  int t1, wf;  // the first t1 tasks hold W items, the remaining ones wf
  bool may_split;  // false when no row is longer than `heavy`: no pieces, no fix-up launch
};
FwdPick fwd_pick(long long M, long long nnz, long long N, bool aligned4, bool aligned2, int max_row);
size_t fwd_workspace_bytes(long long M, long long nnz, long long N);
size_t fwd_bounds_bytes(int n_tasks);
size_t fwd_counter_bytes(const FwdPick& k);
int launch_partition(const int* row_ptr, int M, int nnz, const FwdPick& k, void* part, cudaStream_t s);
// cached_bnd / cached_pick: boundaries a plan computed earlier for the same (W, heavy, n_tasks); may be null
int launch_spmm_forward(const FwdParams& p, void* workspace, size_t workspace_bytes,
                        const void* cached_part, const FwdPick* cached_pick, cudaStream_t s,
                        const PeerSet* d_peers = nullptr);
int sddmm_n_tasks(long long M, long long nnz);
size_t sddmm_bounds_bytes(long long M, long long nnz);
int launch_sddmm_partition(const int* row_ptr, int M, int nnz, Boundary* out, cudaStream_t s);
// cached_bnd: boundaries from launch_sddmm_partition for the same row_ptr (a plan keeps them); may be null
int launch_sddmm(const int* row_ptr, const int* col, int M, int nnz, const float* G, long long ldg,
                 const float* B, long long ldb, int N, float* out, const Boundary* cached_bnd, cudaStream_t s);
int launch_transpose(const int* row_ptr, const int* col, const float* val, int M, int K, int nnz,
                     int* t_row_ptr, int* t_col, float* t_val, int* t_perm, cudaStream_t s);
int launch_gather(const float* src, const int* perm, long long n, float* dst, cudaStream_t s);
int launch_dense_count(const float* dense, int M, int K, long long ld, int* row_ptr, cudaStream_t s);
int launch_dense_fill(const float* dense, int M, int K, long long ld, const int* row_ptr, int* col,
                      float* val, cudaStream_t s);
size_t dense_onepass_scratch_bytes(int M, int K);
bool dense_onepass_supported(const float* dense, int M, int K, long long ld);
int launch_dense_onepass(const float* dense, int M, int K, long long ld, int* row_ptr, int* col, float* val,
                         long long capacity, void* scratch, cudaStream_t s);
int launch_row_stats(const int* row_ptr, int M, int* d_out2 /* max_len, empty */, cudaStream_t s);
int launch_block_diag(const int* row_ptr, int* col, int rows, int M, int K, cudaStream_t s);

// block-dense tensor-core path (tc_gemm.cu): the non-empty 128 x 32 tiles of a CSR matrix, densified
struct TcStore {
  int M = 0, K = 0, MB = 0, KT = 0, n_tiles = 0;
  int kind = 0;                   // element kind of the split (0 tf32, 1 fp16)
  float* row_scale = nullptr;     // [M] fp16 kind: factor that undoes the per-row scaling of A
  // CTA-pair kernel: per pair of 128-row blocks the union of stored k-tiles + each block's slot (n_tiles = zero tile)
  int MB2 = 0, n_pair_tiles = 0;
  int* pair_ptr = nullptr;
  int* pair_kt = nullptr;
  int* pair_slot = nullptr;
  int* blk_ptr = nullptr;         // [MB + 1] first stored tile of every 128-row block
  int* blk_kt = nullptr;          // [n_tiles] k-tile index of every stored tile
  int* scan = nullptr;            // [MB * KT] slot of tile (mb, kt) (exclusive scan of the occupancy flags)
  unsigned char* tiles = nullptr; // n_tiles x 32 KB (hi/lo swizzled atoms)
  size_t tile_bytes = 0;
};
int tc_store_build(const int* row_ptr, const int* col, const float* val, int M, int K, int nnz, int kind,
                   TcStore* st, cudaStream_t s);
int tc_store_refresh(const int* row_ptr, const int* col, const float* val, int nnz, TcStore* st, cudaStream_t s);
void tc_store_free(TcStore* st);
int tc_pick_nt(int MB, int N, int sms);
size_t tc_pack_b_bytes(int K, int N, int nt, int kind);
int launch_tc_forward(const TcStore& st, const int* row_ptr, const int* col, const float* val, const float* B,
                      long long ldb, int N, float* C, long long ldc, const float* bias, unsigned flags, void* ws,
                      size_t ws_bytes, int nt, cudaStream_t s);
// narrow-N forward over shared-memory slabs of B (spmm_slab.cu)
struct SlabStore {
  int N = 0, M = 0, K = 0, S = 0, rows_per_slab = 0, G = 0;
  int* slab_ptr = nullptr;  // [M][S + 1] first non-zero of a row at or beyond every slab boundary
  int* cta_rows = nullptr;  // v1: [G + 1] first row of every persistent CTA
  // v2 (default): a slab-major, sliced copy of A (spmm_slab.cu)
  bool v2 = false;
  int HR = 0, pass_rows = 0, n_slices = 0, total_steps = 0;
  int* vis_off = nullptr;     // [n_slices * S + 1] exclusive prefix of the blocks of every visit, consumption order
  int* cell_gv = nullptr;     // [n_slices * S] visit index of every (slice, slab)
  int* warp_vis = nullptr;    // [G * 32 + 1] first visit of every warp
  int* cta_slices = nullptr;  // [G + 1] first slice of every persistent CTA
  uint4* stream = nullptr;    // [total_steps / 2][32 / (N / 4)] blocks of 2 steps: (col0, val0, col1, val1)
};
bool slab_wanted(int M, int K, long long nnz, int N, const float* B, long long ldb, const float* C, long long ldc);
int slab_store_build(const int* row_ptr, const int* col, const float* val, int M, int K, int nnz, int N, SlabStore* st,
                     cudaStream_t s);
int slab_store_refresh(const int* col, const float* val, SlabStore* st, cudaStream_t s);
void slab_store_free(SlabStore* st);
int launch_spmm_slab(const SlabStore& st, const int* col, const float* val, const float* B, float* C, long long ldc,
                     const float* bias, int relu, cudaStream_t s);
int tc_trace_sync();
int tc_trace_clear();
long long tc_trace_word(int i);
extern int g_tc_last_kernel, g_tc_pair_capacity;  // diagnostics of the last tensor-core launch (tc_gemm.cu)
// pattern of a CSR matrix for the tensor-core SDDMM's sampling epilogue
struct TcSample {
  int M = 0, K = 0, groups = 0, nnz = 0;
  uint32_t* mask = nullptr;  // [groups][M] bit j of (g, m): (m, 32 g + j) is a non-zero
  int* off = nullptr;        // [groups][M] position (in the CSR value array) of the group's first non-zero
};
int tc_sample_build(const int* row_ptr, const int* col, int M, int K, int nnz, TcSample* sm, cudaStream_t s);
void tc_sample_free(TcSample* sm);
size_t tc_sddmm_ws_bytes(int M, int K, int N, int nt, int kind);
int tc_pick_nt_sddmm(int MB, int K, int N, int kind, int sms);
int launch_tc_sddmm(const TcSample& sm, const int* row_ptr, const int* col, const float* G, long long ldg,
                    const float* B, long long ldb, int N, bool transposed, float* out, void* ws, size_t ws_bytes,
                    int nt, int kind, cudaStream_t s);
int launch_sddmm_strided(const int* row_ptr, const int* col, int M, int N, const float* G, long long g_sm,
                         long long g_sn, const float* B, long long b_sk, long long b_sn, float* out, cudaStream_t s);
// generic strided CSR SpMM (any B / C strides, bias, ReLU): the correctness fallback of the extended forward
int launch_spmm_strided(const int* row_ptr, const int* col, const float* val, int M, int N, const float* B,
                        long long b_sk, long long b_sn, float* C, long long c_sm, long long c_sn,
                        const float* bias, bool relu, cudaStream_t s);

// cached scratch per (device, stream), grown on demand
int get_scratch(size_t bytes, void** ptr, cudaStream_t stream);

}  // namespace b200
