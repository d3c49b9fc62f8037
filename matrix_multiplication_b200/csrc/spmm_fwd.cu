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
//    atomics, C is written exactly once per row).  The boundaries are computed once
//    by partition_kernel (cached by a plan; recomputed per call otherwise).
//  * One warp executes one task.  Its 32 lanes are laid out as G groups of LPR
//    lanes: the LPR lanes span one slab of LPR*VEC dense columns (one 128-bit load
//    of a B row per lane), the G groups take consecutive non-zeros of the same
//    row; at a row end the G partial sums are combined with a warp-shuffle
//    reduction.  Control flow is warp-uniform: no divergence inside a task.
//  * The CTA's contiguous col/val segment is staged into shared memory with two
//    TMA 1-D bulk copies (cp.async.bulk + mbarrier; SASS UBLKCP) so that the index
//    fetch is off the critical path of the B gather; lanes read it back with
//    broadcast LDS.  `U` independent B-row gathers are issued before their FMAs.
#include "common.cuh"

namespace b200 {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS = 8;  // warps (= tasks) per CTA

__host__ __device__ inline int smem_capacity(int W, int heavy, int warps = WARPS) {
  return ((warps * W + heavy + 24) + 7) & ~7;  // non-zeros a CTA may have to stage
}

__host__ __device__ inline int mask_words(int cap) { return ((cap >> 3) + 4 + 3) & ~3; }  // keeps s_col 16-byte aligned

// Partition block in memory: [16 B pad][Boundary x (n_tasks + 1)].  It depends only on (row_ptr, W, heavy): a
// plan computes it once per side.
__host__ __device__ inline const Boundary* part_bnd(const void* part) {
  return reinterpret_cast<const Boundary*>(reinterpret_cast<const unsigned char*>(part) + 16);
}

__global__ void __launch_bounds__(256)
partition_kernel(const int* __restrict__ row_ptr, int M, int nnz, int W, int heavy, int n_tasks, int t1, int wf,
                 void* __restrict__ part) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tasks) return;
  const_cast<Boundary*>(part_bnd(part))[t] =
      merge_path_boundary(row_ptr, M, nnz, t == n_tasks ? (long long)M + nnz : task_first_item(t, W, t1, wf), heavy);
}

// A finished row slab of C: local store, plus the multi-GPU epilogue (peer stores / multicast store).
template <int VEC, bool PEERS>
__device__ __forceinline__ void store_c(float* C, const PeerArg<PEERS>& ps, long long off, const Vec<VEC>& o) {
  if constexpr (!PEERS) {
    o.store_stream(C + off);
  } else {
    float* mc = ps.s->mc;
    if (mc != nullptr) {
      o.store_mc(mc + off);  // the switch delivers it to every rank, this one included
      return;
    }
    o.store_stream(C + off);
    const int n_peers = ps.s->n_peers;
    for (int q = 0; q < n_peers; ++q) o.store_stream(ps.s->peer[q] + off);
  }
}

// Fused epilogue of a finished row: + bias[row], ReLU (both optional; the FC-layer caller,
// benchmarks/cusparse_fc_layer.py:41-46, adds the bias as a separate pass over the output).
template <int VEC>
__device__ __forceinline__ void row_epilogue(const FwdParams& p, int row, Vec<VEC>& o) {
  if (p.bias != nullptr) {
    const float b = __ldg(p.bias + row);
#pragma unroll
    for (int v = 0; v < VEC; ++v) o.v[v] += b;
  }
  if (p.relu) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) o.v[v] = fmaxf(o.v[v], 0.f);
  }
}

// A task has just stored one piece (partial sum) of split row `row`.  The pieces of a row are counted on a
// per-(row, column slab) counter; the warp whose increment completes the count adds them up in task order (a
// fixed order: the result does not depend on who does the adding) and writes the row of C, then re-arms the
// counter for the next launch.  This folds the former second launch (a fix-up kernel over the split rows) into
// the main kernel: no spinning, every warp only ever waits for its own loads.
template <int VEC, bool PEERS>
__device__ __forceinline__ void piece_done(const FwdParams& p, const PeerArg<PEERS>& ps, int row, int W, int col0,
                                           bool reader, int lane) {
  const long long first_item = (long long)__ldg(p.row_ptr + row) + row;
  const long long end_item = (long long)__ldg(p.row_ptr + row + 1) + row;  // the row-end item
  const int t_first = (int)task_of_item(first_item, W, p.t1, p.wf);
  int t_end = (int)task_of_item(end_item, W, p.t1, p.wf);
  if (end_item == task_first_item(t_end, W, p.t1, p.wf)) --t_end;  // a boundary on a row-end item is moved past it
  const int n_pieces = t_end - t_first + 1;
  int* cnt = p.counters + (size_t)t_end * gridDim.y + blockIdx.y;
  __threadfence();  // this lane's piece is visible device-wide before the count moves
  __syncwarp();
  int old = 0;
  if (lane == 0) old = atomicAdd(cnt, 1);
  old = __shfl_sync(FULL, old, 0);
  if (old != n_pieces - 1) return;
  __threadfence();
  if (reader) {
    Vec<VEC> sum;
#pragma unroll
    for (int v = 0; v < VEC; ++v) sum.v[v] = 0.f;
    // piece k of the row: k = 0 is the tail piece of task t_first, k >= 1 the head piece of task t_first + k
    const float* base = p.pieces + col0;
    int k = 0;
    for (; k + 4 <= n_pieces; k += 4) {
      Vec<VEC> x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        x[u].load_cg(base + (size_t)(2 * (size_t)(t_first + k + u) + (k + u == 0 ? 1 : 0)) * p.N);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) sum.v[v] += x[u].v[v];
    }
    for (; k < n_pieces; ++k) {
      Vec<VEC> x;
      x.load_cg(base + (size_t)(2 * (size_t)(t_first + k) + (k == 0 ? 1 : 0)) * p.N);
#pragma unroll
      for (int v = 0; v < VEC; ++v) sum.v[v] += x.v[v];
    }
    row_epilogue<VEC>(p, row, sum);
    store_c<VEC, PEERS>(p.C, ps, (long long)row * p.ldc + col0, sum);
  }
  if (lane == 0) *cnt = 0;
}

template <int VEC, int LPR, int U, bool STAGE>
__global__ void __launch_bounds__(WARPS * 32)
spmm_fwd_kernel(FwdParams p, const void* __restrict__ part, int W, int heavy, int n_tasks) {
  const Boundary* bnd = part_bnd(part);
  constexpr int G = 32 / LPR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [mbar 16B][Boundary (WARPS+1), padded to 80B][s_col CAP][s_val CAP]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  Boundary* bnd_s = reinterpret_cast<Boundary*>(smem_raw + 16);
  constexpr int BND_BYTES = ((WARPS + 1) * (int)sizeof(Boundary) + 15) & ~15;
  int* s_col = reinterpret_cast<int*>(smem_raw + 16 + BND_BYTES);
  float* s_val = reinterpret_cast<float*>(s_col + smem_capacity(W, heavy));

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int task0 = blockIdx.x * WARPS;

  if (tid <= WARPS) bnd_s[tid] = bnd[min(task0 + tid, n_tasks)];
  if (STAGE && tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int a0 = bnd_s[0].nz & ~3;  // staged window starts 16-byte aligned
  if (STAGE) {
    const int je = bnd_s[WARPS].nz;
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
  // the non-zero stream of this CTA, indexed by global position
  const int* cs = STAGE ? (s_col - a0) : p.col;
  const float* vs = STAGE ? (s_val - a0) : p.val;

  // ---- this warp's task ---------------------------------------------------------------
  const int t = task0 + warp;
  const Boundary b0 = bnd_s[warp], b1 = bnd_s[warp + 1];
  const int i1 = b1.row, j1 = b1.nz;
  int row = b0.row, pos = b0.nz;
  if (row >= p.M) return;  // padding task

  const int g = lane / LPR, l = lane % LPR;
  const int col0 = blockIdx.y * (LPR * VEC) + l * VEC;
  const bool active = col0 < p.N;
  const bool writer = active && g == 0;
  // inactive lanes gather from column 0 (always valid, never stored) so loads need no predicate
  const float* Bc = p.B + (active ? col0 : 0);

  bool head = pos > __ldg(p.row_ptr + row);  // the first row continues a split heavy row

  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;

  const unsigned ldb = (unsigned)p.ldb;  // < 2^31 (checked by the host): one IMAD.WIDE.U32 per gather
  auto brow = [&](int c) { return Bc + (unsigned long long)(unsigned)c * ldb; };

  // acc += sum over non-zeros [q0, q1) of the current row; group g takes q0 + g, q0 + g + G, ...
  auto accumulate = [&](int q0, int q1) {
    int q = q0;
    for (; q + U * G <= q1; q += U * G) {  // full batches: no predicates
      int c[U];
      float a[U];
      Vec<VEC> b[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        c[k] = cs[q + k * G + g];
        a[k] = vs[q + k * G + g];
      }
#pragma unroll
      for (int k = 0; k < U; ++k) b[k].load(brow(c[k]));
#pragma unroll
      for (int k = 0; k < U; ++k)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(a[k], b[k].v[v], acc[v]);
    }
    if (q < q1) {  // tail batch
      int c[U];
      float a[U];
      Vec<VEC> b[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int idx = q + k * G + g;
        const bool ok = idx < q1;
        c[k] = ok ? cs[idx] : -1;
        a[k] = ok ? vs[idx] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (c[k] >= 0) {
          b[k].load(brow(c[k]));
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) b[k].v[v] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(a[k], b[k].v[v], acc[v]);
    }
  };

  // combine the G partial sums of a row (warp-shuffle reduction); result valid in group 0
  auto reduce_groups = [&]() {
    if (G > 1) {
#pragma unroll
      for (int o = 16; o >= LPR; o >>= 1)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] += __shfl_xor_sync(FULL, acc[v], o);
    }
  };
  auto emit = [&](float* dst, int final_row = -1) {  // store acc (group 0 lanes) and clear it
    if (writer) {
      Vec<VEC> o;
#pragma unroll
      for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
      if (final_row >= 0) row_epilogue<VEC>(p, final_row, o);  // a complete row of C (pieces get it when summed)
      o.store_stream(dst);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  };

  // rows that END inside this task: [row, i1).  Row ends are fetched 32 at a time.
  int head_row = -1;
  int rbase = row;
  int rcache = (rbase + lane < p.M) ? __ldg(p.row_ptr + rbase + lane + 1) : 0;
  while (row < i1) {
    if (row - rbase >= 32) {
      rbase = row;
      rcache = (rbase + lane < p.M) ? __ldg(p.row_ptr + rbase + lane + 1) : 0;
    }
    const int re = __shfl_sync(FULL, rcache, row - rbase);
    accumulate(pos, re);
    pos = re;
    reduce_groups();
    if (head) {
      emit(p.pieces + (size_t)(2 * (size_t)t) * p.N + col0);
      head_row = row;  // counted after the loop (keeps the reduction code out of the hot path's live ranges)
      head = false;
    } else {
      emit(p.C + (long long)row * p.ldc + col0, row);
    }
    ++row;
  }
  // partial last row: a heavy row that goes on into the next task
  const bool tail = i1 < p.M && j1 > pos;
  if (tail) {
    accumulate(pos, j1);
    reduce_groups();
    emit(p.pieces + (size_t)(2 * (size_t)t + (head ? 0 : 1)) * p.N + col0);
  }
  if (head_row >= 0) piece_done<VEC, false>(p, PeerArg<false>{0}, head_row, W, col0, writer, lane);
  if (tail) piece_done<VEC, false>(p, PeerArg<false>{0}, i1, W, col0, writer, lane);
}

// ---- flat kernel (one lane group: N >= 32 or so) ------------------------------------------------
// The warp walks its task's non-zeros in aligned chunks of 8 regardless of row boundaries: per chunk
// four 16-byte broadcast reads of the staged (col, val) segment, 8 independent B-row gathers, 8*VEC
// FMAs per lane.  Row ends are a per-chunk bit mask (built once per CTA from row_ptr, together with
// the row id stored at each row-end position), so the common chunk without a row end runs
// straight-line code and a row end costs one store.  Empty rows are zero-filled by the CTA up front.
template <int VEC, bool TMA, int NB, int NW, int MINB, bool PEERS>
__global__ void __launch_bounds__(NW * 32, MINB)
spmm_flat_kernel(FwdParams p, PeerArg<PEERS> ps, const void* __restrict__ part, int W, int heavy, int n_tasks) {
  const Boundary* bnd = part_bnd(part);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [mbar 16B][Boundary x (WARPS+1), 80B][mask words CAP/8 + 4][s_col CAP][s_val CAP][s_rowend CAP]
  const int cap = smem_capacity(W, heavy, NW);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  Boundary* bnd_s = reinterpret_cast<Boundary*>(smem_raw + 16);
  constexpr int BND_BYTES = ((NW + 1) * (int)sizeof(Boundary) + 15) & ~15;
  unsigned* s_mask = reinterpret_cast<unsigned*>(smem_raw + 16 + BND_BYTES);
  int* s_col = reinterpret_cast<int*>(s_mask + mask_words(cap));
  float* s_val = reinterpret_cast<float*>(s_col + cap);
  int* s_rowend = reinterpret_cast<int*>(s_val + cap);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int task0 = blockIdx.x * NW;

  if (tid <= NW) bnd_s[tid] = bnd[min(task0 + tid, n_tasks)];
  if (TMA && tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int a0 = bnd_s[0].nz & ~7;  // staged window starts on a chunk (32-byte) boundary
  const int je = bnd_s[NW].nz;
  const int n_bulk = TMA ? ((je - a0) & ~3) : 0;
  if (TMA && tid == 0 && n_bulk > 0) {
    mbar_expect_tx(bar, (uint32_t)n_bulk * 8u);
    tma_bulk_g2s_stream(s_col, p.col + a0, (uint32_t)n_bulk * 4u, bar);
    tma_bulk_g2s_stream(s_val, p.val + a0, (uint32_t)n_bulk * 4u, bar);
  }
  for (int q = a0 + n_bulk + tid; q < je; q += blockDim.x) {  // fringe (or everything when !TMA)
    s_col[q - a0] = __ldg(p.col + q);
    s_val[q - a0] = __ldg(p.val + q);
  }
  for (int i = tid; i < ((je - a0 + 7) >> 3); i += blockDim.x) s_mask[i] = 0u;
  __syncthreads();

  const int col0 = blockIdx.y * (32 * VEC) + lane * VEC;
  const bool active = col0 < p.N;

  // rows that end inside this CTA: mark the row ends, zero-fill the empty rows
  {
    const int R0 = bnd_s[0].row, R1 = bnd_s[NW].row;
    for (int base = R0 + warp * 32; base < R1; base += NW * 32) {
      const int r = base + lane;
      int rs = 0, re = 0;
      if (r < R1) {
        rs = __ldg(p.row_ptr + r);
        re = __ldg(p.row_ptr + r + 1);
      }
      if (re > rs) {
        const int pos = re - 1 - a0;
        atomicOr(&s_mask[pos >> 3], 1u << (pos & 7));
        s_rowend[pos] = r;
      }
      unsigned em = __ballot_sync(FULL, r < R1 && re == rs);
      while (em) {
        const int b = __ffs(em) - 1;
        em &= em - 1;
        if (active) {
          Vec<VEC> z;
#pragma unroll
          for (int v = 0; v < VEC; ++v) z.v[v] = 0.f;
          row_epilogue<VEC>(p, base + b, z);
          store_c<VEC, PEERS>(p.C, ps, (long long)(base + b) * p.ldc + col0, z);
        }
      }
    }
  }
  __syncthreads();
  if (TMA && n_bulk > 0) mbar_wait(bar, 0);

  // ---- this warp's task ---------------------------------------------------------------
  const int t = task0 + warp;
  const Boundary b0 = bnd_s[warp], b1 = bnd_s[warp + 1];
  const int j0 = b0.nz, j1 = b1.nz;
  if (j1 <= j0) return;  // no non-zeros in this task
  bool head = j0 > __ldg(p.row_ptr + b0.row);  // the first row continues a split heavy row

  const float* Bc = p.B + (active ? col0 : 0);  // inactive lanes gather column 0 (valid, never stored)
  const unsigned ldb = (unsigned)p.ldb;
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;

  int head_row = -1;
  auto emit = [&](int off) {  // the non-zero at staged offset `off` is the last one of its row
    float* dst;
    if (head) {
      dst = p.pieces + (size_t)(2 * (size_t)t) * p.N + col0;
    } else {
      dst = p.C + (long long)s_rowend[off] * p.ldc + col0;
    }
    if (active) {
      Vec<VEC> o;
#pragma unroll
      for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
      if (!head) row_epilogue<VEC>(p, s_rowend[off], o);
      if (PEERS && !head) {
        store_c<VEC, PEERS>(p.C, ps, dst - p.C, o);
      } else {
        o.store_stream(dst);
      }
    }
    if (head) head_row = s_rowend[off];  // last piece of a split row: counted after the loop
    head = false;
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  };

  // A chunk in flight holds only its 8 gathered B slabs (the values are re-read from shared memory
  // when they are consumed), so nearly every live register is a load destination.
  struct Chunk {
    Vec<VEC> b[8];
    unsigned m;   // row-end bits of the chunk (restricted to this task)
    unsigned vm;  // which of the 8 positions belong to this task
    int off;
  };
  auto fetch = [&](Chunk& ch, int c0) {
    const int off = c0 - a0;
    ch.off = off;
    ch.m = s_mask[off >> 3];
    const int4 ca = *reinterpret_cast<const int4*>(s_col + off);
    const int4 cb = *reinterpret_cast<const int4*>(s_col + off + 4);
    const int c[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
    if (c0 >= j0 && c0 + 8 <= j1) {  // interior chunk: no predicates
      ch.vm = 0xffu;
#pragma unroll
      for (int k = 0; k < 8; ++k) ch.b[k].load(Bc + (unsigned long long)(unsigned)c[k] * ldb);
    } else {  // first / last chunk of the task: elements outside [j0, j1) belong to other tasks
      const int lo = max(j0 - c0, 0), hi = min(j1 - c0, 8);
      ch.vm = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
      ch.m &= ch.vm;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if ((ch.vm >> k) & 1u) {
          ch.b[k].load(Bc + (unsigned long long)(unsigned)c[k] * ldb);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) ch.b[k].v[v] = 0.f;
        }
      }
    }
  };
  auto consume = [&](const Chunk& ch) {
    const float4 va = *reinterpret_cast<const float4*>(s_val + ch.off);
    const float4 vb = *reinterpret_cast<const float4*>(s_val + ch.off + 4);
    float a[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
    if (ch.vm != 0xffu) {  // values of other tasks' positions may be anything (Inf, NaN): mask them
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (!((ch.vm >> k) & 1u)) a[k] = 0.f;
    }
    if (ch.m == 0u) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(a[k], ch.b[k].v[v], acc[v]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(a[k], ch.b[k].v[v], acc[v]);
        if ((ch.m >> k) & 1u) emit(ch.off + k);
      }
    }
  };
  // NB chunks (8 * NB gathers) are issued back to back before their FMAs: bytes in flight per warp,
  // not occupancy, is what feeds the L2 -> SM path ("binding roof").
  for (int c0 = j0 & ~7; c0 < j1; c0 += 8 * NB) {
    Chunk ch[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (i == 0 || c0 + 8 * i < j1) fetch(ch[i], c0 + 8 * i);
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (i == 0 || c0 + 8 * i < j1) consume(ch[i]);
  }
  // the task ends inside a (heavy) row: its partial sum is a piece
  const int last = j1 - 1 - a0;
  const bool tail = ((s_mask[last >> 3] >> (last & 7)) & 1u) == 0u;
  if (tail && active) {
    Vec<VEC> o;
#pragma unroll
    for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
    o.store_stream(p.pieces + (size_t)(2 * (size_t)t + (head ? 0 : 1)) * p.N + col0);
  }
  if (head_row >= 0) piece_done<VEC, PEERS>(p, ps, head_row, W, col0, active, lane);
  if (tail) piece_done<VEC, PEERS>(p, ps, b1.row, W, col0, active, lane);
}

__global__ void zero_rows_kernel(float* C, long long ldc, int M, int N) {
  long long n = (long long)M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    C[(i / N) * ldc + (i % N)] = 0.f;
  }
}

// all-empty shard of a multi-GPU forward: the zero rows must also reach the other ranks' copies of C
__global__ void zero_rows_peers_kernel(float* C, long long ldc, int M, int N, const PeerSet* ps) {
  const long long n = (long long)M * N;
  float* mc = ps->mc;
  const int n_peers = ps->n_peers;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long off = (i / N) * ldc + (i % N);
    if (mc != nullptr) {
      multimem_st(mc + off, 0.f);
    } else {
      C[off] = 0.f;
      for (int q = 0; q < n_peers; ++q) ps->peer[q][off] = 0.f;
    }
  }
}

inline bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <int VEC, int LPR, bool STAGE>
int launch_one(const FwdParams& p, const FwdPick& k, const void* bnd, cudaStream_t s, const PeerSet* peers) {
  constexpr int U = (32 / LPR >= 4) ? 4 : 8;
  const int cap = smem_capacity(k.W, k.heavy);
  constexpr int BND_BYTES = ((WARPS + 1) * (int)sizeof(Boundary) + 15) & ~15;
  const size_t smem = 16 + BND_BYTES + (STAGE ? (size_t)cap * 8 : 0);
  auto kern = spmm_fwd_kernel<VEC, LPR, U, STAGE>;
  if (smem > 48 * 1024) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((unsigned)((k.n_tasks + WARPS - 1) / WARPS), (unsigned)k.slabs);
  kern<<<grid, WARPS * 32, smem, s>>>(p, bnd, k.W, k.heavy, k.n_tasks);
  B200_LAUNCH_OK("spmm_fwd_kernel");
  return B200_OK;
}

template <int VEC, bool TMA, int NB, int NW, int MINB, bool PEERS>
int launch_flat(const FwdParams& p, const FwdPick& k, const void* bnd, cudaStream_t s, const PeerSet* peers) {
  const int cap = smem_capacity(k.W, k.heavy, NW);
  constexpr int BND_BYTES = ((NW + 1) * (int)sizeof(Boundary) + 15) & ~15;
  const size_t smem = 16 + BND_BYTES + (size_t)mask_words(cap) * 4 + (size_t)cap * 12;
  auto kern = spmm_flat_kernel<VEC, TMA, NB, NW, MINB, PEERS>;
  if (smem > 48 * 1024) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((unsigned)((k.n_tasks + NW - 1) / NW), (unsigned)k.slabs);
  PeerArg<PEERS> pa{};
  if constexpr (PEERS) pa.s = peers;
  kern<<<grid, NW * 32, smem, s>>>(p, pa, bnd, k.W, k.heavy, k.n_tasks);
  B200_LAUNCH_OK("spmm_flat_kernel");
  return B200_OK;
}

template <int VEC, bool TMA>
int launch_flat_nb(const FwdParams& p, const FwdPick& k, const void* bnd, cudaStream_t s, const PeerSet* peers) {
  // <gathers in flight per warp / 8, warps per CTA, min CTAs per SM>.  Measured on the R-MAT config
  // (profiles/r1_variant_sweep.md): 4-warp CTAs at 64 registers (8 CTAs/SM) win; deeper per-warp
  // batches lose more to occupancy and spills than they gain in bytes in flight.
  if (peers != nullptr) return launch_flat<VEC, TMA, 1, 4, 8, true>(p, k, bnd, s, peers);
  switch (options().pipeline) {
    case 1: return launch_flat<VEC, TMA, 1, 8, 4, false>(p, k, bnd, s, nullptr);
    case 2: return launch_flat<VEC, TMA, 2, 4, 5, false>(p, k, bnd, s, nullptr);
    case 3: return launch_flat<VEC, TMA, 3, 4, 3, false>(p, k, bnd, s, nullptr);
    default: return launch_flat<VEC, TMA, 1, 4, 8, false>(p, k, bnd, s, nullptr);
  }
}

template <bool STAGE>
int dispatch(const FwdParams& p, const FwdPick& k, const void* bnd, cudaStream_t s, const PeerSet* peers) {
  if (k.lpr != 32 && peers != nullptr) {
    set_error("the multi-GPU epilogue needs N >= 24 dense columns");
    return B200_ERR_INVALID;
  }
  if (k.lpr == 32) {
    if (k.vec == 4) return launch_flat_nb<4, STAGE>(p, k, bnd, s, peers);
    if (k.vec == 2) return launch_flat_nb<2, STAGE>(p, k, bnd, s, peers);
    return launch_flat_nb<1, STAGE>(p, k, bnd, s, peers);
  }
#define B200_CASE(V, L) \
  if (k.vec == V && k.lpr == L) return launch_one<V, L, STAGE>(p, k, bnd, s, peers)
  B200_CASE(4, 16); B200_CASE(4, 8); B200_CASE(4, 4);
  B200_CASE(2, 16); B200_CASE(2, 8); B200_CASE(2, 4);
  B200_CASE(1, 16); B200_CASE(1, 8); B200_CASE(1, 4);
#undef B200_CASE
  set_error("internal: no kernel for vec %d lpr %d", k.vec, k.lpr);
  return B200_ERR_INVALID;
}

}  // namespace

// Lane layout and task size for a problem; depends only on (M, nnz, N, alignment), so a plan
// can cache the boundaries computed for it.
FwdPick fwd_pick(long long M, long long nnz, long long N, bool aligned4, bool aligned2, int max_row) {
  FwdPick k;
  const int vmax = (N % 4 == 0 && aligned4) ? 4 : ((N % 2 == 0 && aligned2) ? 2 : 1);
  if (N >= options().flat_min_n) {
    // one lane group of 32 lanes (flat kernel): the narrowest vector that covers N in one slab
    k.vec = 1;
    while (k.vec < vmax && 32LL * k.vec < N) k.vec <<= 1;
    k.lpr = 32;
  } else {
    k.vec = vmax;
    k.lpr = 4;
    while (k.lpr < 16 && (long long)k.lpr * k.vec < N) k.lpr <<= 1;
  }
  const long long slabs = (N + (long long)k.lpr * k.vec - 1) / ((long long)k.lpr * k.vec);
  int W = options().items_per_task;
  if (W <= 0) {
    // 256 items per warp task; smaller problems get smaller tasks (down to 32 items) so that they still fill
    // the 148 SMs several times over and drain quickly: a task is a serial chain of 8-wide gather batches,
    // ~1 us per batch (measured on 1/8 shards of the R-MAT config: 160-192 items beat 256 by 4-5 %).
    // Every slab of dense columns is its own set of warp tasks, so wide N needs no help.
    const long long spread = (M + nnz) * slabs / (148 * 96);
    W = (int)std::min<long long>(256, std::max<long long>(32, spread));
  }
  W = (W + 31) & ~31;
  if (W > 2048) W = 2048;
  while (W > 128 && (size_t)smem_capacity(W, W) * 12 > 160 * 1024) W >>= 1;
  k.W = W;
  k.heavy = W;
  const long long items = M + nnz;
  // The launch ends when its slowest warp does, and a W-item task is a serial chain of W/8 gather batches (~20 us at
  // W = 256): the last items are cut into short tasks (W/4) so that the tail drains in a quarter of that.  Fine
  // region: about two waves of resident warps, at most a quarter of the items.
  k.wf = std::max(32, (W / 4 + 31) & ~31);
  long long fine_items = std::min<long long>(items / 4, (long long)options().tail_tasks * k.wf);
  if (options().tail_tasks <= 0 || k.wf >= W) fine_items = 0;
  k.t1 = (int)((items - fine_items + W - 1) / W);
  const long long rest = std::max<long long>(0, items - (long long)k.t1 * W);
  k.n_tasks = k.t1 + (int)((rest + k.wf - 1) / k.wf);
  if (k.n_tasks < 1) k.n_tasks = 1;
  k.slabs = (int)((N + (long long)k.lpr * k.vec - 1) / ((long long)k.lpr * k.vec));
  k.may_split = (max_row < 0) || (max_row > k.heavy);
  return k;
}

size_t fwd_bounds_bytes(int n_tasks) {
  return (16 + ((size_t)n_tasks + 2) * sizeof(Boundary) + 255) & ~(size_t)255;
}

// one int per (task, column slab): arrival counters of the split rows' pieces (zero between launches)
size_t fwd_counter_bytes(const FwdPick& k) {
  return k.may_split ? (((size_t)k.n_tasks * (size_t)k.slabs * sizeof(int) + 255) & ~(size_t)255) : 0;
}

size_t fwd_workspace_bytes(long long M, long long nnz, long long N) {
  // upper bound over the lane layouts fwd_pick may choose for this N (alignment decides the vector width,
  // hence the number of column slabs): counters are sized for one slab per 4 columns
  const FwdPick k = fwd_pick(M, nnz, N > 0 ? N : 1, true, true, -1);
  const size_t slabs_max = (size_t)(N + 3) / 4 + 1;
  return fwd_bounds_bytes(k.n_tasks) + (size_t)2 * (size_t)k.n_tasks * (size_t)N * sizeof(float) +
         (((size_t)k.n_tasks * slabs_max * sizeof(int) + 255) & ~(size_t)255) + 256;
}

int launch_partition(const int* row_ptr, int M, int nnz, const FwdPick& k, void* part, cudaStream_t s) {
  const int n = k.n_tasks + 1;
  partition_kernel<<<(n + 255) / 256, 256, 0, s>>>(row_ptr, M, nnz, k.W, k.heavy, k.n_tasks, k.t1, k.wf, part);
  B200_LAUNCH_OK("partition_kernel");
  return B200_OK;
}

// workspace layout: [partition block (unless cached)][pieces 2 * n_tasks * N][piece counters (unless the caller
// keeps a persistent, zeroed set: FwdParams::counters)]
int launch_spmm_forward(const FwdParams& pin, void* workspace, size_t workspace_bytes,
                        const void* cached_part, const FwdPick* cached_pick, cudaStream_t s,
                        const PeerSet* peers) {
  FwdParams p = pin;
  if (p.M == 0 || p.N == 0) return B200_OK;
  if (p.nnz == 0 && (p.bias != nullptr || p.relu) && peers == nullptr) {  // all rows empty: C = epilogue(0)
    return launch_spmm_strided(p.row_ptr, p.col, p.val, p.M, p.N, p.B, p.ldb, 1, p.C, p.ldc, 1, p.bias, p.relu != 0, s);
  }
  if (p.nnz == 0) {  // all rows empty: C = 0
    int blocks = (int)std::min<long long>(((long long)p.M * p.N + 255) / 256, 148 * 8);
    if (peers != nullptr) zero_rows_peers_kernel<<<blocks, 256, 0, s>>>(p.C, p.ldc, p.M, p.N, peers);
    else zero_rows_kernel<<<blocks, 256, 0, s>>>(p.C, p.ldc, p.M, p.N);
    B200_LAUNCH_OK("zero_rows_kernel");
    return B200_OK;
  }
  const bool a4 = is_aligned(p.B, 16) && is_aligned(p.C, 16) && (p.ldb % 4 == 0) && (p.ldc % 4 == 0);
  const bool a2 = is_aligned(p.B, 8) && is_aligned(p.C, 8) && (p.ldb % 2 == 0) && (p.ldc % 2 == 0);
  FwdPick k = fwd_pick(p.M, p.nnz, p.N, a4, a2, -1);
  const bool use_cache = cached_part != nullptr && cached_pick != nullptr && cached_pick->W == k.W &&
                         cached_pick->heavy == k.heavy && cached_pick->n_tasks == k.n_tasks &&
                         cached_pick->t1 == k.t1 && cached_pick->wf == k.wf;
  p.t1 = k.t1;
  p.wf = k.wf;
  if (use_cache) k.may_split = cached_pick->may_split;
  const size_t part_bytes = use_cache ? 0 : fwd_bounds_bytes(k.n_tasks);
  const size_t pieces_bytes =
      k.may_split ? (((size_t)2 * (size_t)k.n_tasks * (size_t)p.N * sizeof(float) + 255) & ~(size_t)255) : 0;
  const size_t counter_bytes = (p.counters == nullptr) ? fwd_counter_bytes(k) : 0;
  const size_t need = part_bytes + pieces_bytes + counter_bytes;
  if (workspace == nullptr && need > 0) {
    void* w = nullptr;
    int rc = get_scratch(need, &w, s);
    if (rc != B200_OK) return rc;
    workspace = w;
    workspace_bytes = need;
  }
  if (workspace_bytes < need) {
    set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
    return B200_ERR_INVALID;
  }
  if (need > 0 && !is_aligned(workspace, 16)) {
    set_error("workspace must be 16-byte aligned");
    return B200_ERR_INVALID;
  }
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  const void* part = cached_part;
  if (!use_cache) {
    int rc = launch_partition(p.row_ptr, p.M, p.nnz, k, ws, s);
    if (rc != B200_OK) return rc;
    part = ws;
  }
  p.pieces = reinterpret_cast<float*>(ws + part_bytes);
  if (counter_bytes > 0) {  // no persistent counters: zero a set inside the workspace for this call
    p.counters = reinterpret_cast<int*>(ws + part_bytes + pieces_bytes);
    B200_CUDA_OK(cudaMemsetAsync(p.counters, 0, counter_bytes, s));
  }
  const int want_stage = (p.stage >= 0) ? p.stage : options().stage_mode;
  const bool stage = want_stage == 1 && is_aligned(p.col, 16) && is_aligned(p.val, 16);
  return stage ? dispatch<true>(p, k, part, s, peers) : dispatch<false>(p, k, part, s, peers);
}

}  // namespace b200
