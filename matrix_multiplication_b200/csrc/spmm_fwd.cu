// spmm_fwd.cu — forward CSR SpMM  C[M,N] = A_csr[M,K] * B[K,N]  (fp32, row-major B/C).
//
// Replaces cusparse_mm_wrapper (reference src/baseline_mm.cu:167-216) and
// spmm_kernel / naive_spmm_wrapper (src/naive_sparse_mm.cu:24-136).  Not a port:
//
//  * Work is cut with a merge path over (row ends + non-zeros): every task owns
//    `W` consecutive items, so short rows, empty rows and heavy rows all cost the
//    same.  Task boundaries are snapped to row starts for light rows (<= W nnz);
//    heavy rows are split exactly and their partial sums go to a scratch "piece"
//    buffer that a tiny fix-up kernel adds up in task order (deterministic, no
//    atomics, C is written exactly once per row).
//  * A task is executed by a virtual warp of LPR lanes laid along the dense
//    dimension (LPR * VEC columns = one slab): one 128-bit load of a B row per
//    lane per non-zero, accumulation in CSR order.
//  * The CTA's contiguous col/val segment is staged into shared memory with two
//    TMA 1-D bulk copies (cp.async.bulk + mbarrier; SASS UBLKCP) — the index
//    fetch is off the critical path of the B gather.  Lanes then read the
//    segment 32-at-a-time and broadcast with shuffles, issuing `U` independent
//    B-row gathers before the FMAs (memory-level parallelism).
#include "common.cuh"

namespace b200 {

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <int LPR>
struct Cfg {
  static constexpr int GROUPS = 32 / LPR;               // virtual warps per warp
  static constexpr int WARPS = (LPR >= 16) ? 8 : 4;     // warps per CTA
  static constexpr int TASKS = WARPS * GROUPS;          // tasks per CTA
  static constexpr int U = (LPR < 8) ? LPR : 8;         // gathers in flight per batch
};

__host__ __device__ inline int smem_capacity(int tasks, int W, int heavy) {
  return ((tasks * W + heavy + 8) + 3) & ~3;  // non-zeros a CTA may have to stage
}

template <int LPR, int VEC, bool STAGE>
__global__ void __launch_bounds__(Cfg<LPR>::WARPS * 32)
spmm_fwd_kernel(FwdParams p, int W, int heavy, int n_tasks) {
  using C = Cfg<LPR>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [mbar 16B][Boundary (TASKS+1)][pad16][s_col CAP][s_val CAP]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  Boundary* bnd_s = reinterpret_cast<Boundary*>(smem_raw + 16);
  constexpr int BND_BYTES = ((C::TASKS + 1) * (int)sizeof(Boundary) + 15) & ~15;
  const int cap = smem_capacity(C::TASKS, W, heavy);
  int* s_col = reinterpret_cast<int*>(smem_raw + 16 + BND_BYTES);
  float* s_val = reinterpret_cast<float*>(s_col + cap);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int grp = lane / LPR, gl = lane % LPR;
  const long long task0 = (long long)blockIdx.x * C::TASKS;

  if (tid <= C::TASKS) {
    long long t = task0 + tid;
    if (t > n_tasks) t = n_tasks;
    bnd_s[tid] = merge_path_boundary(p.row_ptr, p.M, p.nnz, t * (long long)W, heavy);
  }
  if (STAGE && tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int jb = bnd_s[0].nz, je = bnd_s[C::TASKS].nz;
  const int a0 = jb & ~3;  // staged window starts 16-byte aligned
  if (STAGE) {
    const int n_bulk = (je - a0) & ~3;
    if (tid == 0 && n_bulk > 0) {
      mbar_expect_tx(bar, (uint32_t)n_bulk * 8u);
      tma_bulk_g2s(s_col, p.col + a0, (uint32_t)n_bulk * 4u, bar);
      tma_bulk_g2s(s_val, p.val + a0, (uint32_t)n_bulk * 4u, bar);
    }
    for (int q = a0 + n_bulk + tid; q < je; q += blockDim.x) {  // < 4 fringe elements
      s_col[q - a0] = __ldg(p.col + q);
      s_val[q - a0] = __ldg(p.val + q);
    }
    __syncthreads();
    if (n_bulk > 0) mbar_wait(bar, 0);
  }

  // ---- this virtual warp's task -----------------------------------------------------
  const int tl = warp * C::GROUPS + grp;
  const long long t = task0 + tl;
  const Boundary b0 = bnd_s[tl], b1 = bnd_s[tl + 1];
  const int i0 = b0.row, j0 = b0.nz, i1 = b1.row, j1 = b1.nz;

  const int col0 = blockIdx.y * (LPR * VEC) + gl * VEC;
  const bool active = col0 < p.N;
  const float* Bc = p.B + col0;

  int row = i0;
  bool head = (i0 < p.M) && (j0 > __ldg(p.row_ptr + i0));  // first row continues a split row
  int row_end = (row < p.M) ? __ldg(p.row_ptr + row + 1) : 0x7fffffff;
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;

  auto flush = [&](int r) {
    if (active) {
      Vec<VEC> o;
#pragma unroll
      for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
      if (head) {
        o.store(p.pieces + (size_t)(2 * t) * p.N + col0);
        if (gl == 0) *p.split_flag = 1;
      } else {
        o.store_stream(p.C + (long long)r * p.ldc + col0);
      }
    }
    head = false;
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  };

  int n_iter = (j1 - j0 + LPR - 1) / LPR;
  if (LPR < 32) {
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) n_iter = max(n_iter, __shfl_xor_sync(FULL, n_iter, o));
  }

  for (int it = 0; it < n_iter; ++it) {
    const int chunk = j0 + it * LPR;
    const int idx = chunk + gl;
    int creg = 0;
    float vreg = 0.f;
    if (idx < j1) {
      if (STAGE) {
        creg = s_col[idx - a0];
        vreg = s_val[idx - a0];
      } else {
        creg = __ldg(p.col + idx);
        vreg = __ldg(p.val + idx);
      }
    }
#pragma unroll
    for (int b = 0; b < LPR; b += C::U) {
      Vec<VEC> bv[C::U];
      float vv[C::U];
      // phase 1: U independent gathers of B rows
#pragma unroll
      for (int k = 0; k < C::U; ++k) {
        const int c = __shfl_sync(FULL, creg, b + k, LPR);
        vv[k] = __shfl_sync(FULL, vreg, b + k, LPR);
        if (active && (chunk + b + k) < j1) {
          bv[k].load(Bc + (long long)c * p.ldb);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) bv[k].v[v] = 0.f;
        }
      }
      // phase 2: consume in CSR order, closing rows as their ends are reached
#pragma unroll
      for (int k = 0; k < C::U; ++k) {
        const int pp = chunk + b + k;
        if (pp < j1) {
          while (pp == row_end) {
            flush(row);
            ++row;
            row_end = __ldg(p.row_ptr + row + 1);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] = fmaf(vv[k], bv[k].v[v], acc[v]);
        }
      }
    }
  }

  // rows that end at or before the task end (the current one, then empty rows)
  while (row < i1) {
    flush(row);
    ++row;
  }
  // partial last row (only when a heavy row is split at the task end)
  if (i1 < p.M && j1 > j0 && j1 > __ldg(p.row_ptr + i1)) {
    if (active) {
      Vec<VEC> o;
#pragma unroll
      for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
      o.store(p.pieces + (size_t)(2 * t + (head ? 0 : 1)) * p.N + col0);
      if (gl == 0) *p.split_flag = 1;
    }
  }
}

// Adds up the pieces of every split row in task order and writes the row of C.
// One virtual warp per interior task boundary; exits at once when no row was split.
template <int LPR, int VEC>
__global__ void __launch_bounds__(256)
spmm_fixup_kernel(FwdParams p, int W, int heavy, int n_tasks) {
  if (*reinterpret_cast<volatile int*>(p.split_flag) == 0) return;
  constexpr int GROUPS_PER_CTA = 256 / LPR;
  const int grp = threadIdx.x / LPR, gl = threadIdx.x % LPR;
  const long long b = 1 + (long long)blockIdx.x * GROUPS_PER_CTA + grp;
  if (b >= n_tasks) return;
  const Boundary bd = merge_path_boundary(p.row_ptr, p.M, p.nnz, b * (long long)W, heavy);
  const int i = bd.row;
  if (i >= p.M) return;
  const int rs = __ldg(p.row_ptr + i);
  if (bd.nz <= rs) return;  // boundary is at a row start: nothing split here
  const Boundary nb = merge_path_boundary(p.row_ptr, p.M, p.nnz, (b + 1) * (long long)W, heavy);
  if (nb.row == i) return;  // the row goes on; the boundary where it ends does the sum
  const long long t_first = ((long long)rs + i) / W;
  const int col0 = blockIdx.y * (LPR * VEC) + gl * VEC;
  if (col0 >= p.N) return;
  Vec<VEC> sum, x;
  sum.load(p.pieces + (size_t)(2 * t_first + 1) * p.N + col0);
  for (long long t = t_first + 1; t <= b; ++t) {
    x.load(p.pieces + (size_t)(2 * t) * p.N + col0);
#pragma unroll
    for (int v = 0; v < VEC; ++v) sum.v[v] += x.v[v];
  }
  sum.store_stream(p.C + (long long)i * p.ldc + col0);
}

struct Pick {
  int vec, lpr, W, heavy, n_tasks, slabs;
};

Pick pick(long long M, long long nnz, long long N, bool aligned4, bool aligned2) {
  Pick k;
  k.vec = (N % 4 == 0 && aligned4) ? 4 : ((N % 2 == 0 && aligned2) ? 2 : 1);
  if (k.vec == 4) {
    k.lpr = 4;
    while (k.lpr < 32 && (long long)k.lpr * 4 < N) k.lpr <<= 1;
  } else {
    k.lpr = 32;
  }
  int W = options().items_per_task;
  if (W <= 0) W = (k.lpr == 32) ? 256 : 128;
  W = (W + 31) & ~31;
  if (W > 2048) W = 2048;
  {  // keep the staged segment of a CTA within ~160 KB of shared memory
    const int groups = 32 / k.lpr, warps = (k.lpr >= 16) ? 8 : 4;
    while (W > 128 && (size_t)smem_capacity(warps * groups, W, W) * 8 > 160 * 1024) W >>= 1;
  }
  k.W = W;
  k.heavy = W;
  long long items = M + nnz;
  k.n_tasks = (int)((items + W - 1) / W);
  if (k.n_tasks < 1) k.n_tasks = 1;
  k.slabs = (int)((N + (long long)k.lpr * k.vec - 1) / ((long long)k.lpr * k.vec));
  return k;
}

template <int LPR, int VEC, bool STAGE>
int launch_one(const FwdParams& p, const Pick& k, cudaStream_t s) {
  using C = Cfg<LPR>;
  const int cap = smem_capacity(C::TASKS, k.W, k.heavy);
  constexpr int BND_BYTES = ((C::TASKS + 1) * (int)sizeof(Boundary) + 15) & ~15;
  size_t smem = 16 + BND_BYTES + (STAGE ? (size_t)cap * 8 : 0);
  auto kern = spmm_fwd_kernel<LPR, VEC, STAGE>;
  if (smem > 48 * 1024) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((unsigned)((k.n_tasks + C::TASKS - 1) / C::TASKS), (unsigned)k.slabs);
  kern<<<grid, C::WARPS * 32, smem, s>>>(p, k.W, k.heavy, k.n_tasks);
  B200_LAUNCH_OK("spmm_fwd_kernel");
  if (k.n_tasks > 1) {
    constexpr int GROUPS_PER_CTA = 256 / LPR;
    dim3 fgrid((unsigned)((k.n_tasks - 1 + GROUPS_PER_CTA - 1) / GROUPS_PER_CTA), (unsigned)k.slabs);
    spmm_fixup_kernel<LPR, VEC><<<fgrid, 256, 0, s>>>(p, k.W, k.heavy, k.n_tasks);
    B200_LAUNCH_OK("spmm_fixup_kernel");
  }
  return B200_OK;
}

template <bool STAGE>
int dispatch(const FwdParams& p, const Pick& k, cudaStream_t s) {
  if (k.vec == 4) {
    switch (k.lpr) {
      case 4: return launch_one<4, 4, STAGE>(p, k, s);
      case 8: return launch_one<8, 4, STAGE>(p, k, s);
      case 16: return launch_one<16, 4, STAGE>(p, k, s);
      default: return launch_one<32, 4, STAGE>(p, k, s);
    }
  }
  if (k.vec == 2) return launch_one<32, 2, STAGE>(p, k, s);
  return launch_one<32, 1, STAGE>(p, k, s);
}

inline bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

__global__ void zero_rows_kernel(float* C, long long ldc, int M, int N) {
  long long n = (long long)M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    C[(i / N) * ldc + (i % N)] = 0.f;
  }
}

}  // namespace

size_t fwd_workspace_bytes(long long M, long long nnz, long long N) {
  // sized for the smallest task size the launcher may pick (W >= 32 after rounding,
  // but auto picks >= 128); keep in sync with pick().
  int W = options().items_per_task;
  if (W <= 0 || W > 128) W = 128;  // pick() never goes below min(128, option)
  W = (W + 31) & ~31;
  long long n_tasks = (M + nnz + W - 1) / W;
  if (n_tasks < 1) n_tasks = 1;
  return 256 + (size_t)2 * (size_t)n_tasks * (size_t)N * sizeof(float);
}

int launch_spmm_forward(const FwdParams& pin, void* workspace, size_t workspace_bytes,
                        cudaStream_t s) {
  FwdParams p = pin;
  if (p.M == 0 || p.N == 0) return B200_OK;
  if (p.nnz == 0) {  // all rows empty: C = 0
    int blocks = (int)std::min<long long>(((long long)p.M * p.N + 255) / 256, 148 * 8);
    zero_rows_kernel<<<blocks, 256, 0, s>>>(p.C, p.ldc, p.M, p.N);
    B200_LAUNCH_OK("zero_rows_kernel");
    return B200_OK;
  }
  const bool a4 = is_aligned(p.B, 16) && is_aligned(p.C, 16) && (p.ldb % 4 == 0) && (p.ldc % 4 == 0);
  const bool a2 = is_aligned(p.B, 8) && is_aligned(p.C, 8) && (p.ldb % 2 == 0) && (p.ldc % 2 == 0);
  const Pick k = pick(p.M, p.nnz, p.N, a4, a2);
  const size_t need = 256 + (size_t)2 * (size_t)k.n_tasks * (size_t)p.N * sizeof(float);
  if (workspace == nullptr) {
    void* w = nullptr;
    int rc = get_scratch(need, &w);
    if (rc != B200_OK) return rc;
    workspace = w;
    workspace_bytes = need;
  }
  if (workspace_bytes < need) {
    set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
    return B200_ERR_INVALID;
  }
  if (!is_aligned(workspace, 16)) {
    set_error("workspace must be 16-byte aligned");
    return B200_ERR_INVALID;
  }
  p.split_flag = reinterpret_cast<int*>(workspace);
  p.pieces = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  B200_CUDA_OK(cudaMemsetAsync(p.split_flag, 0, sizeof(int), s));
  const int want_stage = (p.stage >= 0) ? p.stage : options().stage_mode;
  const bool stage = want_stage == 1 && is_aligned(p.col, 16) && is_aligned(p.val, 16);
  return stage ? dispatch<true>(p, k, s) : dispatch<false>(p, k, s);
}

}  // namespace b200
