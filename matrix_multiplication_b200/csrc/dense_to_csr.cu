// dense_to_csr.cu — dense row-major fp32 [M,K] -> CSR (row_ptr int32, col_ind int32, val fp32).
//
// Replaces dense_to_csr (reference src/baseline_mm.cu:218-264, cuSPARSE
// DenseToSparse) and the per-call `a.to_sparse_csr()` of matmuls.py:215,228.
// Bit-exact contract: keep x != 0 (so NaN is kept and -0.0 is dropped), row-major
// order, values copied bit for bit (the test is done on the bit pattern so that
// flush-to-zero settings cannot turn a denormal into a dropped element).
//
//   one pass (onepass_kernel): count, cross-CTA prefix (decoupled look-back) and fill while the tile is held in
//          registers — the dense matrix is read once.  The two-phase kernels below remain for the
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

// ---- one pass: count, global prefix and fill in ONE kernel, the dense matrix is read exactly once -----------------
// The matrix is cut into chunks of 512 consecutive elements of one row (a row owns ceil(K / 512) chunks); a warp
// holds one chunk in registers (four 128-bit loads per lane), 16 warps form a tile.  Per chunk the keep bits are
// counted with warp ballots + popc (every kept element gets its slot inside the chunk, column order preserved);
// the tile's total goes through a decoupled look-back across CTAs (tile ids are handed out by an atomic ticket, so a
// tile only ever waits for tiles that are already running), which gives every chunk its global offset; the kept
// (column, value) pairs are then written straight from the registers.  row_ptr[r] is the global offset of the
// chunk that starts row r.  Writes beyond `capacity` are dropped (the scan still completes, so row_ptr and the
// total stay exact and the caller can re-run the fill with enough room).
constexpr int OP_CHUNK = 512;
constexpr unsigned long long OP_AGG = 1ull << 62, OP_PREFIX = 2ull << 62, OP_VAL = (1ull << 62) - 1;

template <int OP_WARPS>
__global__ void __launch_bounds__(OP_WARPS * 32)
onepass_kernel(const float* __restrict__ dense, int M, int K, long long ld, int cpr, long long n_chunks,
               int* __restrict__ row_ptr, int* __restrict__ col, float* __restrict__ val, long long capacity,
               unsigned long long* __restrict__ status, int* __restrict__ ticket) {
  // Persistent CTAs, all co-resident (the launcher sizes the grid from the occupancy), tiles dealt round-robin:
  // tile t belongs to CTA t % gridDim.x, so every predecessor of a tile is either finished or in flight on a running
  // CTA: the look-back cannot starve.  The loads of a CTA's next tile are issued before the current tile goes through
  // ballots, look-back and stores (two register sets), so the memory system never idles; the whole CTA looks back
  // (one predecessor per thread and step), so a walk across all concurrently running tiles takes 2-3 round trips.
  constexpr int NT = OP_WARPS * 32;
  __shared__ int s_wsum[OP_WARPS];
  __shared__ long long s_part[OP_WARPS];
  __shared__ int s_stop[OP_WARPS];
  __shared__ long long s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const long long n_tiles = (n_chunks + OP_WARPS - 1) / OP_WARPS;
  volatile unsigned long long* st = status;

  auto load_tile = [&](long long tile, float (&e)[4][4]) {
    const long long chunk = tile * OP_WARPS + warp;
    if (tile < n_tiles && chunk < n_chunks) {
      const int row = (int)(chunk / cpr);
      const int c0 = (int)(chunk % cpr) * OP_CHUNK;
      const float* r = dense + (long long)row * ld;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 128 + lane * 4;
        if (c + 3 < K) {
          const float4 x = __ldcs(reinterpret_cast<const float4*>(r + c));
          e[u][0] = x.x; e[u][1] = x.y; e[u][2] = x.z; e[u][3] = x.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) e[u][j] = (c + j < K) ? __ldcs(r + c + j) : 0.f;
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) e[u][j] = 0.f;
    }
  };

  float cur[4][4], nxt[4][4];
  long long tile = blockIdx.x;
  load_tile(tile, cur);
  for (; tile < n_tiles; tile += gridDim.x) {
    load_tile(tile + gridDim.x, nxt);  // in flight during everything below

    const long long chunk = tile * OP_WARPS + warp;
    const bool valid = chunk < n_chunks;
    const int row = valid ? (int)(chunk / cpr) : 0;
    const int c0 = valid ? (int)(chunk % cpr) * OP_CHUNK : 0;
    int pos[4];  // slot of this lane's first kept element of group u, relative to the chunk
    int total = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int cnt = (int)keep(cur[u][0]) + (int)keep(cur[u][1]) + (int)keep(cur[u][2]) + (int)keep(cur[u][3]);
      const unsigned b1 = __ballot_sync(FULL, cnt >= 1), b2 = __ballot_sync(FULL, cnt >= 2);
      const unsigned b3 = __ballot_sync(FULL, cnt >= 3), b4 = __ballot_sync(FULL, cnt >= 4);
      pos[u] = total + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt) + __popc(b4 & lt);
      total += __popc(b1) + __popc(b2) + __popc(b3) + __popc(b4);
    }
    if (lane == 0) s_wsum[warp] = total;
    __syncthreads();
    // every warp: exclusive prefix of the warp totals (tiny), tile aggregate
    int w_excl = 0, aggregate = 0;
#pragma unroll
    for (int i = 0; i < OP_WARPS; ++i) {
      const int v = s_wsum[i];
      if (i < warp) w_excl += v;
      aggregate += v;
    }
    if (tid == 0) {
      st[tile] = (tile == 0 ? OP_PREFIX : OP_AGG) | (unsigned long long)aggregate;
      __threadfence();
    }
    // look-back by the whole CTA: thread i inspects predecessor (base - i); warp w holds the w-th nearest 32
    long long exclusive = 0;
    for (long long base = tile - 1; base >= 0; base -= NT) {
      const long long idx = base - tid;
      unsigned long long v = OP_PREFIX;  // before the first tile: a prefix of 0
      if (idx >= 0) {
        do { v = st[idx]; } while ((v >> 62) == 0ull);
      }
      const unsigned has_prefix = __ballot_sync(FULL, (v >> 62) == 2ull);
      const int stop = has_prefix ? __ffs(has_prefix) - 1 : 32;
      long long contrib = (lane <= stop) ? (long long)(v & OP_VAL) : 0;
#pragma unroll
      for (int o = 16; o; o >>= 1) contrib += __shfl_xor_sync(FULL, contrib, o);
      if (lane == 0) {
        s_part[warp] = contrib;
        s_stop[warp] = has_prefix ? 1 : 0;
      }
      __syncthreads();
      bool done = false;
#pragma unroll
      for (int i = 0; i < OP_WARPS; ++i) {  // nearest warp first; stop at the first one that met a prefix
        if (!done) {
          exclusive += s_part[i];
          done = s_stop[i] != 0;
        }
      }
      __syncthreads();
      if (done) break;
    }
    if (tid == 0 && tile != 0) {
      __threadfence();
      st[tile] = OP_PREFIX | (unsigned long long)(exclusive + aggregate);
    }
    if (tid == 0 && tile == n_tiles - 1) {  // the last tile: total = nnz
      row_ptr[M] = (int)(exclusive + aggregate);
      ticket[1] = (int)(exclusive + aggregate);
    }
    if (valid) {
      const long long base = exclusive + w_excl;
      if (c0 == 0 && lane == 0) row_ptr[row] = (int)base;
      if (total != 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          long long q = base + pos[u];
          const int c = c0 + u * 128 + lane * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (keep(cur[u][j])) {
              if (q < capacity) {
                col[q] = c + j;
                val[q] = cur[u][j];
              }
              ++q;
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) cur[u][j] = nxt[u][j];
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

// One-pass conversion.  scratch: (n_tiles + 1) * 8 + 16 bytes, zeroed here.  Needs 16-byte aligned rows.
// Small tiles (4 warps = 2048 elements): many CTAs per SM, so that the loads of some overlap the look-back and the
// stores of others.
static int onepass_warps() { return options().d2c_warps > 0 ? options().d2c_warps : 8; }
size_t dense_onepass_scratch_bytes(int M, int K) {
  const long long cpr = ((long long)K + OP_CHUNK - 1) / OP_CHUNK;
  const long long n_tiles = ((long long)M * cpr + 3) / 4;  // sized for the smallest tile
  return (size_t)(n_tiles + 1) * 8 + 16;
}
bool dense_onepass_supported(const float* dense, int M, int K, long long ld) {
  return M > 0 && K >= 64 && (reinterpret_cast<uintptr_t>(dense) % 16 == 0) && (ld % 4 == 0);
}
int launch_dense_onepass(const float* dense, int M, int K, long long ld, int* row_ptr, int* col, float* val,
                         long long capacity, void* scratch, cudaStream_t s) {
  const int cpr = (K + OP_CHUNK - 1) / OP_CHUNK;
  const long long n_chunks = (long long)M * cpr;
  const int OP_WARPS = onepass_warps();
  const long long n_tiles = (n_chunks + OP_WARPS - 1) / OP_WARPS;
  if (n_tiles >= 0x7fffffffLL) {
    set_error("dense_to_csr one-pass: matrix too large");
    return B200_ERR_INVALID;
  }
  B200_CUDA_OK(cudaMemsetAsync(scratch, 0, dense_onepass_scratch_bytes(M, K), s));
  int* ticket = reinterpret_cast<int*>(scratch);
  unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(scratch) + 16);
  // all CTAs must be co-resident (see the kernel): size the grid from the occupancy the driver reports
  int per_sm = 1;
  auto occ = [&](auto kern, int threads) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, 0);
    return std::max(n, 1);
  };
  if (OP_WARPS == 16) per_sm = occ(onepass_kernel<16>, 512);
  else if (OP_WARPS == 8) per_sm = occ(onepass_kernel<8>, 256);
  else if (OP_WARPS == 2) per_sm = occ(onepass_kernel<2>, 64);
  else per_sm = occ(onepass_kernel<4>, 128);
  const unsigned grid = (unsigned)std::min<long long>(n_tiles, (long long)device_sm_count() * per_sm);
#define B200_ONEPASS(W)                                                                                        \
  onepass_kernel<W><<<grid, W * 32, 0, s>>>(dense, M, K, ld, cpr, n_chunks, row_ptr, col, val, capacity, status, ticket)
  if (OP_WARPS == 16) B200_ONEPASS(16);
  else if (OP_WARPS == 8) B200_ONEPASS(8);
  else if (OP_WARPS == 2) B200_ONEPASS(2);
  else B200_ONEPASS(4);
#undef B200_ONEPASS
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
