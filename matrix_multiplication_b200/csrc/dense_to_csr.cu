// dense_to_csr.cu — dense row-major fp32 [M,K] -> CSR (row_ptr int32, col_ind int32, val fp32).
//
// Replaces dense_to_csr (reference src/baseline_mm.cu:218-264, cuSPARSE
// DenseToSparse) and the per-call `a.to_sparse_csr()` of matmuls.py:215,228.
// Bit-exact contract: keep x != 0 (so NaN is kept and -0.0 is dropped), row-major
// order, values copied bit for bit (the test is done on the bit pattern so that
// flush-to-zero settings cannot turn a denormal into a dropped element).
//
//   one pass (onepass_kernel): count, cross-CTA offsets (round totals, no chain) and fill while the chunks are held
//          in registers — the dense matrix is read once.  The two-phase kernels below remain for the
//          analysis / convert split of the reference API and for unaligned or very narrow inputs:
//   count: one warp per row, 128-bit loads, ballot/popc -> per-row counts
//   scan : one CTA, shuffle-based block scan with a running carry -> row_ptr
//   fill : one warp per row, ballot + popc(mask below lane) gives every kept
//          element its output slot; writes are contiguous per row.
#include "common.cuh"

namespace b200 {
namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ bool keep(float x) { return (__float_as_uint(x) & 0x7fffffffu) != 0u; }

template <bool VEC4>
__global__ void __launch_bounds__(256)
count_kernel(const float* __restrict__ dense, int M, int K, long long ld, int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M; row += warps_per_grid) {
    const float* r = dense + (long long)row * ld;
    int cnt = 0;
    if (VEC4) {
      const int K4 = K >> 2;
      const float4* r4 = reinterpret_cast<const float4*>(r);
      for (int c = lane; c < K4; c += 32) {
        float4 x = __ldcs(r4 + c);
        cnt += keep(x.x) + keep(x.y) + keep(x.z) + keep(x.w);
      }
      for (int c = (K4 << 2) + lane; c < K; c += 32) cnt += keep(__ldcs(r + c));
    } else {
      for (int c = lane; c < K; c += 32) cnt += keep(__ldcs(r + c));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    if (lane == 0) counts[row] = cnt;
  }
}

// In-place exclusive scan of counts[0..M) into row_ptr[0..M]; row_ptr[M] = total.
__global__ void __launch_bounds__(1024) scan_kernel(int* __restrict__ data, int M) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  constexpr int PER = 4;
  for (int base = 0; base < M; base += 1024 * PER) {
    int v[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int idx = base + tid * PER + i;
      v[i] = (idx < M) ? data[idx] : 0;
      sum += v[i];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(FULL, wi, o);
        if (lane >= o) wi += n;
      }
      warp_sums[lane] = wi - w;  // exclusive prefix of warp sums
    }
    __syncthreads();
    const int carry = carry_s;
    int excl = carry + warp_sums[warp] + (incl - sum);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int idx = base + tid * PER + i;
      if (idx < M) data[idx] = excl;
      excl += v[i];
    }
    __syncthreads();
    if (tid == 1023) carry_s = excl;  // total including this chunk
    __syncthreads();
  }
  if (tid == 0) data[M] = carry_s;
}

__global__ void __launch_bounds__(256)
fill_kernel(const float* __restrict__ dense, int M, int K, long long ld,
            const int* __restrict__ row_ptr, int* __restrict__ col, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M; row += warps_per_grid) {
    const float* r = dense + (long long)row * ld;
    int base = __ldg(row_ptr + row);
    const int end = __ldg(row_ptr + row + 1);
    if (base == end) continue;  // nothing to write for this row
    for (int c0 = 0; c0 < K; c0 += 128) {
      float x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 32 + lane;
        x[u] = (c < K) ? __ldcs(r + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool k = keep(x[u]);
        const unsigned ball = __ballot_sync(FULL, k);
        if (k) {
          const int pos = base + __popc(ball & lt);
          col[pos] = c0 + u * 32 + lane;
          val[pos] = x[u];
        }
        base += __popc(ball);
      }
    }
  }
}

// 128-bit variant of the fill: lane l owns columns 4l .. 4l+3 of every 128-column chunk.  Output order is
// by column, i.e. lane-major: slot = base + (kept in lanes < l) + (kept before j within the lane).  The
// cross-lane prefix comes from four ballots ("lane keeps >= k elements") instead of a shuffle scan.
// Rows that keep more than a quarter of their elements take the one-element-per-lane path instead (their
// time goes into the stores, which that path coalesces perfectly).
__global__ void __launch_bounds__(256)
fill_kernel_v4(const float* __restrict__ dense, int M, int K, long long ld,
               const int* __restrict__ row_ptr, int* __restrict__ col, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const int K4 = K >> 2;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M; row += warps_per_grid) {
    const float* r = dense + (long long)row * ld;
    int base = __ldg(row_ptr + row);
    const int end = __ldg(row_ptr + row + 1);
    if (base == end) continue;
    if ((long long)(end - base) * 4 > K) {
      // dense row (> 25 % kept): the stores dominate; one element per lane keeps them perfectly coalesced
      for (int c0 = 0; c0 < K; c0 += 128) {
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * 32 + lane;
          x[u] = (c < K) ? __ldcs(r + c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool k = keep(x[u]);
          const unsigned ball = __ballot_sync(FULL, k);
          if (k) {
            const int pos = base + __popc(ball & lt);
            col[pos] = c0 + u * 32 + lane;
            val[pos] = x[u];
          }
          base += __popc(ball);
        }
      }
      continue;
    }
    const float4* r4 = reinterpret_cast<const float4*>(r);
    for (int c4 = 0; c4 < K4; c4 += 64) {  // two 128-column chunks per iteration: two loads in flight per lane
      float4 x[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = c4 + u * 32 + lane;
        x[u] = (i < K4) ? __ldcs(r4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float e[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
        const bool k0 = keep(e[0]), k1 = keep(e[1]), k2 = keep(e[2]), k3 = keep(e[3]);
        const int cnt = (int)k0 + (int)k1 + (int)k2 + (int)k3;
        const unsigned b1 = __ballot_sync(FULL, cnt >= 1), b2 = __ballot_sync(FULL, cnt >= 2);
        const unsigned b3 = __ballot_sync(FULL, cnt >= 3), b4 = __ballot_sync(FULL, cnt >= 4);
        int pos = base + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt) + __popc(b4 & lt);
        const int c0 = (c4 + u * 32 + lane) << 2;
        if (k0) { col[pos] = c0;     val[pos] = e[0]; ++pos; }
        if (k1) { col[pos] = c0 + 1; val[pos] = e[1]; ++pos; }
        if (k2) { col[pos] = c0 + 2; val[pos] = e[2]; ++pos; }
        if (k3) { col[pos] = c0 + 3; val[pos] = e[3]; }
        base += __popc(b1) + __popc(b2) + __popc(b3) + __popc(b4);
      }
    }
    // K not a multiple of 4: the last 1-3 columns
    for (int c = (K4 << 2) + lane; c < K; c += 32) {  // at most one iteration, lanes 0..2
      const float xv = __ldcs(r + c);
      const bool k = keep(xv);
      const unsigned ball = __ballot_sync(__activemask(), k);
      if (k) {
        const int pos = base + __popc(ball & lt);
        col[pos] = c;
        val[pos] = xv;
      }
    }
  }
}

__global__ void row_stats_kernel(const int* __restrict__ row_ptr, int M, int* __restrict__ out2) {
  int mx = 0, empty = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M; r += gridDim.x * blockDim.x) {
    int len = __ldg(row_ptr + r + 1) - __ldg(row_ptr + r);
    mx = max(mx, len);
    empty += (len == 0);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(FULL, mx, o));
    empty += __shfl_xor_sync(FULL, empty, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out2, mx);
    atomicAdd(out2 + 1, empty);
  }
}

// col[p] += (row / M) * K for every non-zero of every row: warp per row
__global__ void __launch_bounds__(256)
block_diag_kernel(const int* __restrict__ row_ptr, int* __restrict__ col, int rows, int M, int K) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps_per_grid) {
    const int shift = (row / M) * K;
    if (shift == 0) continue;
    const int e = __ldg(row_ptr + row + 1);
    for (int p = __ldg(row_ptr + row) + lane; p < e; p += 32) col[p] += shift;
  }
}

// ---- one pass: count, global offsets and fill in ONE kernel; the dense matrix comes out of HBM once ---------------
// The matrix is cut into chunks of 1024 consecutive elements of one row (a row owns ceil(K / 1024) chunks).
// Persistent CTAs, one per SM, all co-resident (cooperative launch).  Round i hands the NC chunks
// (i * G + b) * NC ... to CTA b, so a round is one contiguous stretch of the matrix (148 x NC x 4 KB).
// A CTA is split into two kinds of warps that only meet through counters:
//
//   NC count warps  stream their chunk of every round (eight independent 128-bit loads per lane, compare + popc,
//                   shuffle sum).  A chunk that keeps at most OP_SMALL elements is compacted right away — ballot/popc
//                   slots, column order preserved — into the chunk's (column, value) slot of a shared-memory ring;
//                   the chunk's total (tagged with the round) goes to shared memory and is added to the CTA's
//                   word of the round in global memory (total | arrivals << 16).  Count warps never wait for another
//                   CTA, only for their own CTA's fill warps when those fall more than LAG rounds behind.
//   NF fill warps   read ALL CTAs' totals of a round (one word per thread; complete LAG rounds ago in the steady
//                   state), which gives the CTA's global offset (every CTA keeps the running total itself); the
//                   chunk totals in shared memory give every chunk its offset inside the CTA; compacted chunks are
//                   copied from the ring to their place (coalesced stores), denser chunks are re-read — the LAG + 1
//                   rounds in flight (<= 60 MB) are L2 hits — and written with ballot/popc slots.
//
// Nobody needs anybody's prefix — only totals, which depend on nothing — so there is no chain between CTAs (the
// look-back chain is what made the earlier one-pass variants slower than reading the matrix twice), and the
// streaming side is as free-running as a plain copy.  row_ptr[r] is the global offset of the chunk that starts
// row r.  Writes beyond `capacity` are dropped (the totals still complete, so row_ptr and nnz stay exact and the
// caller can re-run the fill with enough room).
constexpr int OP_CHUNK = 1024;
constexpr int OP_G = OP_CHUNK / 128;  // 128-bit loads per lane and chunk
constexpr int OP_SMALL = 192;         // chunks keeping at most this many elements are compacted by the count warp

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fill_barrier(int threads) {
  asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory");
}

// One 128-column group of a chunk: lane l holds columns c .. c+3.  Calls put(slot, column, value) for every kept
// element, slots counted from `base` in column order; returns the kept elements of the group.
template <class Put>
__device__ __forceinline__ int compact_group(const float4& x, int c, unsigned lt, int base, Put put) {
  const float e[4] = {x.x, x.y, x.z, x.w};
  const bool k0 = keep(e[0]), k1 = keep(e[1]), k2 = keep(e[2]), k3 = keep(e[3]);
  const int cnt = (int)k0 + (int)k1 + (int)k2 + (int)k3;
  const unsigned a1 = __ballot_sync(FULL, cnt >= 1);
  if (a1 == 0u) return 0;  // nothing kept in these 128 columns
  const unsigned a2 = __ballot_sync(FULL, cnt >= 2), a3 = __ballot_sync(FULL, cnt >= 3);
  const unsigned a4 = __ballot_sync(FULL, cnt >= 4);
  int q = base + __popc(a1 & lt) + __popc(a2 & lt) + __popc(a3 & lt) + __popc(a4 & lt);
  if (k0) put(q++, c, e[0]);
  if (k1) put(q++, c + 1, e[1]);
  if (k2) put(q++, c + 2, e[2]);
  if (k3) put(q, c + 3, e[3]);
  return __popc(a1) + __popc(a2) + __popc(a3) + __popc(a4);
}

template <bool STREAM>
__device__ __forceinline__ float4 load_group(const float* r, int c, int K) {
  float4 x;
  if (c + 3 < K) {
    x = STREAM ? __ldcs(reinterpret_cast<const float4*>(r + c)) : __ldcg(reinterpret_cast<const float4*>(r + c));
  } else {
    x.x = (c < K) ? __ldcg(r + c) : 0.f;
    x.y = (c + 1 < K) ? __ldcg(r + c + 1) : 0.f;
    x.z = (c + 2 < K) ? __ldcg(r + c + 2) : 0.f;
    x.w = 0.f;  // c + 3 >= K here
  }
  return x;
}

template <int NC, int NF, int D, int P>
__global__ void __launch_bounds__((NC + NF + 1) * 32, 1)
onepass_kernel(const float* __restrict__ dense, int M, int K, long long ld, int cpr, long long n_chunks, int rounds,
               int* __restrict__ row_ptr, int* __restrict__ col, float* __restrict__ val, long long capacity,
               unsigned* __restrict__ status, int dbg) {
  static_assert(NC * OP_CHUNK < 65536 && NC < 65536, "a CTA's round total and its arrivals share one 32-bit word");
  static_assert(OP_CHUNK < 2048, "a chunk total and its round tag share one 32-bit word");
  static_assert(NC <= 32 && NF <= NC, "one lane of the offset warp per chunk; every fill warp owns at least one chunk");
  static_assert((P & (P - 1)) == 0 && P >= OP_SMALL, "pair buffers are power-of-two rings");
  constexpr int R = 16;          // rounds a CTA may have between "counted" and "written" (bookkeeping rings)
  constexpr int DENSE_LAG = 3;   // ... when the chunk will be re-read by the fill warps (it has to stay in the L2)
  extern __shared__ __align__(16) unsigned char op_smem[];
  // per count warp a ring of P compacted (column, value) pairs: [NC][P] x 8 bytes; then the
  // count warps' landing slots for raw chunks: [NC][D][OP_CHUNK] floats, filled by TMA bulk copies
  uint2* s_pair = reinterpret_cast<uint2*>(op_smem);  // (column, value bits)
  float* s_raw = reinterpret_cast<float*>(op_smem + (size_t)NC * P * 8);
  __shared__ __align__(8) uint64_t s_bar[NC][D];
  __shared__ int s_tot[R][NC];         // chunk totals, tagged with the round: ((round + 1) << 11) | total
  __shared__ unsigned s_start[R][NC];  // where the chunk's pairs start in its count warp's ring (running position)
  __shared__ long long s_off[R][NC];   // global offset of every chunk of a round (written by the offset warp)
  __shared__ int s_off_tag[R];         // round + 1 once s_off[round % R] is complete
  __shared__ int s_fill_prog[NF];      // rounds completed by every fill warp
  __shared__ unsigned s_tail[NC];      // ring position up to which a count warp's pairs have been written out
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int G = gridDim.x, b = blockIdx.x;
  for (int t = tid; t < R * NC; t += blockDim.x) (&s_tot[0][0])[t] = 0;
  if (tid < R) s_off_tag[tid] = 0;
  if (tid < NF) s_fill_prog[tid] = 0;
  if (tid < NC) s_tail[tid] = 0u;
  __syncthreads();

  if (warp < NC) {
    // ---- count warps ----
    // Every count warp owns D landing slots: the bulk copy of its chunk of round i + D is in flight while the chunk
    // of round i (read into registers) is counted and compacted.  The part of a chunk that is a whole number of
    // 16-byte groups inside the row comes through the slot; a ragged row end (K % 4 != 0) is read directly.
    float* raw = s_raw + (size_t)warp * D * OP_CHUNK;
    auto chunk_geom = [&](int round, int& row, int& c0) {
      const long long chunk = ((long long)round * G + b) * NC + warp;
      if (round >= rounds || chunk >= n_chunks) return false;
      row = (int)(chunk / cpr);
      c0 = (int)(chunk % cpr) * OP_CHUNK;
      return true;
    };
    auto issue = [&](int round) {  // lane 0 only
      int row, c0;
      uint64_t* bar = &s_bar[warp][round % D];
      uint32_t bytes = 0;
      if (chunk_geom(round, row, c0)) bytes = (uint32_t)(min(OP_CHUNK, K - c0) & ~3) * 4u;
      mbar_expect_tx(bar, bytes);
      if (bytes) tma_bulk_g2s(raw + (size_t)(round % D) * OP_CHUNK, dense + (long long)row * ld + c0, bytes, bar);
    };
    if (lane == 0) {
      for (int d = 0; d < D; ++d) mbar_init(&s_bar[warp][d], 1);
      mbar_fence_init();
      for (int d = 0; d < D; ++d) issue(d);
    }
    __syncwarp();
    unsigned head = 0u;  // running position in this warp's pair ring
    for (int i = 0; i < rounds; ++i) {
      const int slot = i % R;
      int row = 0, c0 = 0;
      const bool valid = chunk_geom(i, row, c0);
      mbar_wait(&s_bar[warp][i % D], (uint32_t)(i / D) & 1u);
      float4 x[OP_G];
      // kept elements of this lane in each of the 8 groups of 128 columns, one byte per group (a group keeps at most
      // 128): a warp scan of the two packed words gives every lane its slot in all 8 groups at once
      unsigned w0 = 0u, w1 = 0u;
      if (valid) {
        const float* r = dense + (long long)row * ld;
        const float* sl = raw + (size_t)(i % D) * OP_CHUNK;
#pragma unroll
        for (int u = 0; u < OP_G; ++u) {
          const int o = u * 128 + lane * 4;
          if (c0 + o + 3 < K) x[u] = *reinterpret_cast<const float4*>(sl + o);
          else x[u] = load_group<false>(r, c0 + o, K);
        }
#pragma unroll
        for (int u = 0; u < OP_G; ++u) {
          const unsigned c = (unsigned)keep(x[u].x) + (unsigned)keep(x[u].y) + (unsigned)keep(x[u].z) + (unsigned)keep(x[u].w);
          if (u < 4) w0 |= c << (8 * u);
          else w1 |= c << (8 * (u - 4));
        }
      }
      const unsigned own0 = w0, own1 = w1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned n0 = __shfl_up_sync(FULL, w0, o), n1 = __shfl_up_sync(FULL, w1, o);
        if (lane >= o) {
          w0 += n0;
          w1 += n1;
        }
      }
      const unsigned t0 = __shfl_sync(FULL, w0, 31), t1 = __shfl_sync(FULL, w1, 31);  // group totals, one byte each
      const unsigned h0 = __vsadu4(t0, 0u), h1 = __vsadu4(t1, 0u);                    // byte sums
      const int cnt = (int)(h0 + h1);
      // the sum depends on every lane's reads of the slot: it is free again, start the copy of round i + D
      if (lane == 0) issue(i + D);
      const bool small = cnt != 0 && cnt <= OP_SMALL;
      if (!(dbg & 1)) {
        // bookkeeping slot of round i - R and (dense chunks) the L2 window: wait for fill warp (warp % NF)
        const int need = (cnt > OP_SMALL ? i - DENSE_LAG : i - R) + 1;
        while (*(volatile int*)&s_fill_prog[warp % NF] < need) __nanosleep(32);
        // room for the pairs in this warp's ring
        if (small)
          while (head + (unsigned)cnt - *(volatile unsigned*)&s_tail[warp] > (unsigned)P) __nanosleep(32);
      }
      if (small) {
        uint2* pp = s_pair + (size_t)warp * P;
        unsigned gbase = head;  // ring position of the group's first kept element
#pragma unroll
        for (int u = 0; u < OP_G; ++u) {
          const unsigned incl = ((u < 4 ? w0 : w1) >> (8 * (u & 3))) & 0xffu;
          const unsigned mine = ((u < 4 ? own0 : own1) >> (8 * (u & 3))) & 0xffu;
          if (mine) {
            unsigned q = gbase + incl - mine;
            const int c = c0 + u * 128 + lane * 4;
            const float e[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (keep(e[j])) {
                pp[q & (unsigned)(P - 1)] = make_uint2((unsigned)(c + j), __float_as_uint(e[j]));
                ++q;
              }
          }
          gbase += ((u < 4 ? t0 : t1) >> (8 * (u & 3))) & 0xffu;
        }
        if (lane == 0) s_start[slot][warp] = head;
        head += (unsigned)cnt;
        __threadfence_block();  // pairs and start are in shared memory before the tagged total is
        __syncwarp();
      }
      if (lane == 0) {
        // total tagged with the round: one word is both the datum and the "ready" flag
        *(volatile int*)&s_tot[slot][warp] = ((i + 1) << 11) | cnt;
        atomicAdd(status + (size_t)i * G + b, (unsigned)cnt + 0x10000u);
      }
    }
    return;
  }
  if (dbg & 1) return;

  if (warp == NC) {
    // ---- offset warp: the only one that talks to the other CTAs ----
    // Per round: all CTAs' totals (G words, one or more per lane, requested two rounds ahead) -> kept elements below
    // this CTA and in the whole round; this CTA's chunk totals -> exclusive scan; s_off[round][k] = global offset of
    // chunk k.  Runs as far ahead of the fill warps as the count warps are.
    constexpr int PF = 5;  // status words per lane held in registers (G <= 160); CTAs beyond that are polled directly
    // Three rounds of status words are kept in registers and every poll refreshes the incomplete words of all three
    // with independent loads, so the L2 round trips of consecutive rounds overlap instead of adding up.
    unsigned pa[PF], pb[PF], pc3[PF];
    auto refresh = [&](int round, unsigned (&v)[PF]) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int t = lane + 32 * j;
        if (round < rounds && t < G && (v[j] >> 16) != (unsigned)NC) v[j] = ld_relaxed_u32(status + (size_t)round * G + t);
      }
    };
    auto complete = [&](int round, const unsigned (&v)[PF]) {
      bool ok = true;
#pragma unroll
      for (int j = 0; j < PF; ++j) ok = ok && (lane + 32 * j >= G || (v[j] >> 16) == (unsigned)NC);
      return __all_sync(FULL, ok);
    };
#pragma unroll
    for (int j = 0; j < PF; ++j) pa[j] = pb[j] = pc3[j] = 0u;
    refresh(0, pa);
    refresh(1, pb);
    refresh(2, pc3);
    long long carry = 0;  // kept elements of all earlier rounds (every CTA keeps its own copy)
    auto one_round = [&](int i, unsigned (&v)[PF], unsigned (&v1)[PF], unsigned (&v2)[PF]) {
      while (!complete(i, v)) {
        __nanosleep(100);
        refresh(i, v);
        refresh(i + 1, v1);
        refresh(i + 2, v2);
      }
      int below = 0, all = 0;
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int t = lane + 32 * j;
        const int n = t < G ? (int)(v[j] & 0xffffu) : 0;
        all += n;
        if (t < b) below += n;
      }
      for (int t = lane + 32 * PF; t < G; t += 32) {  // more than 160 CTAs
        const unsigned* p = status + (size_t)i * G + t;
        unsigned w = ld_relaxed_u32(p);
        while ((w >> 16) != (unsigned)NC) {
          __nanosleep(64);
          w = ld_relaxed_u32(p);
        }
        all += (int)(w & 0xffffu);
        if (t < b) below += (int)(w & 0xffffu);
      }
#pragma unroll
      for (int j = 0; j < PF; ++j) v[j] = 0u;
      refresh(i + 3, v);  // this register set now follows round i + 3
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        below += __shfl_xor_sync(FULL, below, o);
        all += __shfl_xor_sync(FULL, all, o);
      }
      const int slot = i % R;
      int mine = 0;
      if (lane < NC) {
        int t = *(volatile int*)&s_tot[slot][lane];
        while ((t >> 11) != i + 1) {
          __nanosleep(60);
          t = *(volatile int*)&s_tot[slot][lane];
        }
        mine = t & 0x7ff;
      }
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += n;
      }
      if (lane < NC) s_off[slot][lane] = carry + below + (incl - mine);
      __threadfence_block();
      __syncwarp();
      if (lane == 0) *(volatile int*)&s_off_tag[slot] = i + 1;
      carry += all;
    };
    for (int i = 0; i < rounds; i += 3) {
      one_round(i, pa, pb, pc3);
      if (i + 1 < rounds) one_round(i + 1, pb, pc3, pa);
      if (i + 2 < rounds) one_round(i + 2, pc3, pa, pb);
    }
    if (b == 0 && lane == 0) row_ptr[M] = (int)carry;
    return;
  }

  // ---- fill warps: chunk k of every round belongs to fill warp k % NF ----
  const int wf = warp - NC - 1;
  for (int i = 0; i < rounds; ++i) {
    const int slot = i % R;
    while (*(volatile int*)&s_off_tag[slot] != i + 1) __nanosleep(200);
    __threadfence_block();
    for (int k = wf; k < NC; k += NF) {
      const long long chunk = ((long long)i * G + b) * NC + k;
      if (chunk >= n_chunks) break;
      const int total = *(volatile int*)&s_tot[slot][k] & 0x7ff;
      const long long base = s_off[slot][k];
      const int row = (int)(chunk / cpr);
      const int c0 = (int)(chunk % cpr) * OP_CHUNK;
      if (c0 == 0 && lane == 0) row_ptr[row] = (int)base;
      if (total == 0) continue;
      if (total <= OP_SMALL) {
        const uint2* pp = s_pair + (size_t)k * P;
        const unsigned start = s_start[slot][k];
        for (int idx = lane; idx < total; idx += 32) {
          const long long q = base + idx;
          if (q < capacity) {
            const uint2 pr = pp[(start + (unsigned)idx) & (unsigned)(P - 1)];
            col[q] = (int)pr.x;
            val[q] = __uint_as_float(pr.y);
          }
        }
        __syncwarp();
        if (lane == 0) *(volatile unsigned*)&s_tail[k] = start + (unsigned)total;  // the count warp may reuse the room
        continue;
      }
      const float* r = dense + (long long)row * ld;
      int done = 0;
#pragma unroll
      for (int h = 0; h < OP_G; h += 4) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = load_group<true>(r, c0 + (h + u) * 128 + lane * 4, K);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          done += compact_group(x[u], c0 + (h + u) * 128 + lane * 4, lt, done, [&](int q, int c, float v) {
            const long long qq = base + q;
            if (qq < capacity) {
              col[qq] = c;
              val[qq] = v;
            }
          });
      }
    }
    __syncwarp();  // every lane is done with the ring slot and the totals of this round
    if (lane == 0) *(volatile int*)&s_fill_prog[wf] = i + 1;
  }
}

inline int warp_grid(int M) {
  long long ctas = ((long long)M + 7) / 8;  // 8 warps per CTA, one row per warp
  long long cap = 148LL * 8;
  return (int)(ctas < 1 ? 1 : (ctas > cap ? cap : ctas));
}

}  // namespace

int launch_dense_count(const float* dense, int M, int K, long long ld, int* row_ptr, cudaStream_t s) {
  if (M == 0) {
    B200_CUDA_OK(cudaMemsetAsync(row_ptr, 0, sizeof(int), s));
    return B200_OK;
  }
  const bool v4 = (reinterpret_cast<uintptr_t>(dense) % 16 == 0) && (ld % 4 == 0);
  if (v4) count_kernel<true><<<warp_grid(M), 256, 0, s>>>(dense, M, K, ld, row_ptr);
  else count_kernel<false><<<warp_grid(M), 256, 0, s>>>(dense, M, K, ld, row_ptr);
  B200_LAUNCH_OK("dense_count_kernel");
  scan_kernel<<<1, 1024, 0, s>>>(row_ptr, M);
  B200_LAUNCH_OK("dense_scan_kernel");
  return B200_OK;
}

// One-pass conversion.  scratch: one 32-bit word per (round, CTA), zeroed here.  Needs 16-byte aligned rows.
// d2c_warps: count warps per CTA (16, 20 or 24 = default; the rest of the 32 warps: one offset warp + fill warps).
// Measured on 16384^2 (profiles/r2_dense2csr.md): 24 count warps are the fastest at every sparsity >= 90 %.
static int onepass_count_warps() {
  const int w = options().d2c_warps;
  return (w == 16 || w == 20) ? w : 24;
}
size_t dense_onepass_scratch_bytes(int M, int K) {
  // upper bound for every grid the launcher may choose: a CTA's round has at least 16 chunks
  const long long cpr = ((long long)K + OP_CHUNK - 1) / OP_CHUNK;
  const long long n_chunks = (long long)M * cpr;
  return (size_t)((n_chunks + 15) / 16 + 2 * 1024) * 4 + 16;
}
bool dense_onepass_supported(const float* dense, int M, int K, long long ld) {
  return M > 0 && K >= 64 && (reinterpret_cast<uintptr_t>(dense) % 16 == 0) && (ld % 4 == 0);
}
int launch_dense_onepass(const float* dense, int M, int K, long long ld, int* row_ptr, int* col, float* val,
                         long long capacity, void* scratch, cudaStream_t s) {
  int cpr = (K + OP_CHUNK - 1) / OP_CHUNK;
  long long n_chunks = (long long)M * cpr;
  const int NC = onepass_count_warps();
  const int lag = options().d2c_lag & 0xff;
  int dbg = options().d2c_lag >> 8;  // timing experiments only (results are wrong with any bit set)
  // one CTA per SM, all co-resident (fewer CTAs when the matrix is small)
  const long long ctas_needed = (n_chunks + NC - 1) / NC;
  int grid = (int)std::min<long long>(ctas_needed, (long long)device_sm_count());
  const long long per_round = (long long)grid * NC;
  const long long rounds_ll = (n_chunks + per_round - 1) / per_round;
  if (rounds_ll >= (1LL << 20) || (size_t)rounds_ll * grid * 4 + 16 > dense_onepass_scratch_bytes(M, K)) {
    set_error("dense_to_csr one-pass: matrix too large");
    return B200_ERR_INVALID;
  }
  int rounds = (int)rounds_ll;
  unsigned* status = reinterpret_cast<unsigned*>(scratch);
  B200_CUDA_OK(cudaMemsetAsync(status, 0, (size_t)rounds * grid * 4, s));
  void* args[] = {(void*)&dense, (void*)&M, (void*)&K, (void*)&ld, (void*)&cpr, (void*)&n_chunks, (void*)&rounds,
                  (void*)&row_ptr, (void*)&col, (void*)&val, (void*)&capacity, (void*)&status, (void*)&dbg};
  // cooperative launch: the driver guarantees that every CTA is resident (the CTAs read each other's round totals)
  const void* fn;
  int D, P;
  const bool small_ring = lag == 1;  // (experiments) 256 pairs per count warp instead of 512
  if (NC == 16) { D = 2; P = small_ring ? 256 : 512; fn = small_ring ? (const void*)onepass_kernel<16, 15, 2, 256> : (const void*)onepass_kernel<16, 15, 2, 512>; }
  else if (NC == 20) { D = small_ring ? 2 : 1; P = small_ring ? 256 : 512; fn = small_ring ? (const void*)onepass_kernel<20, 11, 2, 256> : (const void*)onepass_kernel<20, 11, 1, 512>; }
  else { D = 1; P = small_ring ? 256 : 512; fn = small_ring ? (const void*)onepass_kernel<24, 7, 1, 256> : (const void*)onepass_kernel<24, 7, 1, 512>; }
  const size_t smem = (size_t)NC * P * 8 + (size_t)NC * D * OP_CHUNK * 4;
  B200_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B200_CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(1024), args, smem, s));
  B200_LAUNCH_OK("dense_onepass_kernel");
  return B200_OK;
}

int launch_dense_fill(const float* dense, int M, int K, long long ld, const int* row_ptr, int* col,
                      float* val, cudaStream_t s) {
  if (M == 0 || K == 0) return B200_OK;
  const bool v4 = (reinterpret_cast<uintptr_t>(dense) % 16 == 0) && (ld % 4 == 0) && K >= 4;
  if (v4) fill_kernel_v4<<<warp_grid(M), 256, 0, s>>>(dense, M, K, ld, row_ptr, col, val);
  else fill_kernel<<<warp_grid(M), 256, 0, s>>>(dense, M, K, ld, row_ptr, col, val);
  B200_LAUNCH_OK("dense_fill_kernel");
  return B200_OK;
}

int launch_block_diag(const int* row_ptr, int* col, int rows, int M, int K, cudaStream_t s) {
  if (rows == 0) return B200_OK;
  block_diag_kernel<<<warp_grid(rows), 256, 0, s>>>(row_ptr, col, rows, M, K);
  B200_LAUNCH_OK("block_diag_kernel");
  return B200_OK;
}

int launch_row_stats(const int* row_ptr, int M, int* d_out2, cudaStream_t s) {
  B200_CUDA_OK(cudaMemsetAsync(d_out2, 0, 2 * sizeof(int), s));
  if (M == 0) return B200_OK;
  int blocks = (int)std::min<long long>(((long long)M + 255) / 256, 148 * 8);
  row_stats_kernel<<<blocks, 256, 0, s>>>(row_ptr, M, d_out2);
  B200_LAUNCH_OK("row_stats_kernel");
  return B200_OK;
}

}  // namespace b200
