// capi.cu — the extern "C" surface declared in include/spmm_b200.h: argument checks,
// plans (inspect / execute / clean), host-buffer entry points, options and errors.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200 {

std::atomic<uint64_t> g_launches{0};
static thread_local char tl_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}

Options& options() {
  static Options o;
  return o;
}

int device_sm_count() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

// scratch for the plan-less entry points when the caller passes no workspace: one buffer per (device, stream), grown on
// demand.  Launches on one stream are ordered, so a buffer is never shared by concurrent work; growing it waits for
// that stream only.  (Not usable under CUDA-graph capture: pass a workspace there.)
static std::mutex g_scratch_mu;
struct ScratchBuf { void* p = nullptr; size_t bytes = 0; };
static std::map<std::pair<int, cudaStream_t>, ScratchBuf> g_scratch;

int get_scratch(size_t bytes, void** ptr, cudaStream_t stream) {
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  ScratchBuf& sb = g_scratch[{dev, stream}];
  if (sb.bytes < bytes) {
    if (sb.p) {
      B200_CUDA_OK(cudaStreamSynchronize(stream));
      B200_CUDA_OK(cudaFree(sb.p));
      sb.p = nullptr;
      sb.bytes = 0;
    }
    size_t want = std::max(bytes + bytes / 4, (size_t)1 << 20);
    if (cudaMalloc(&sb.p, want) != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaMalloc of %zu scratch bytes failed", want);
      return B200_ERR_ALLOC;
    }
    sb.bytes = want;
  }
  *ptr = sb.p;
  return B200_OK;
}

static int check_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available: the SpMM engine has no CPU fallback");
    return B200_ERR_NO_DEVICE;
  }
  return B200_OK;
}

static int check_csr_dims(int64_t M, int64_t K, int64_t nnz) {
  B200_REQUIRE(M >= 0 && K >= 0 && nnz >= 0, "negative dimension");
  B200_REQUIRE(M < 0x7fffffffLL && K < 0x7fffffffLL && nnz < 0x7fffffffLL, "int32 index range exceeded");
  return B200_OK;
}

}  // namespace b200

using namespace b200;

struct b200_plan {
  const int* row_ptr;
  const int* col;
  const float* val;
  int M, K, nnz;
  int max_row, empty_rows;
  int device = 0;  // the device that was current at creation: every later call must run on it
  // cached transposed CSR (owned)
  int* t_row_ptr = nullptr;
  int* t_col = nullptr;
  float* t_val = nullptr;
  int* t_perm = nullptr;
  bool t_built = false;
  bool t_val_stale = false;
  // owned scratch for forward / backward_B
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // cached merge-path boundaries of A (fwd) and A^T (bwd), valid for the FwdPick stored beside them: the task size
  // depends on the number of column slabs (hence on N), so a plan keeps a few sets per side and products that
  // alternate between widths do not re-partition
  struct BndSet { void* part; FwdPick pick; };
  std::vector<BndSet> bnd_sets[2];
  void* bnd[2] = {nullptr, nullptr};   // the set selected for the current call
  FwdPick pick[2];
  bool bnd_valid[2] = {false, false};
  int* cnt[2] = {nullptr, nullptr};  // persistent arrival counters of split-row pieces (self-resetting)
  size_t cnt_bytes[2] = {0, 0};
  int t_max_row = -1;
  Boundary* sd_bnd = nullptr;  // SDDMM task boundaries (heavy threshold 0)
  // host-buffer pipeline (b200_plan_forward_host): device B / C, three streams, per-panel events
  float* hB = nullptr; float* hC = nullptr; size_t hB_bytes = 0, hC_bytes = 0;
  cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t hev[2][16] = {};
  // extra streams / events when a panel copy is split by rows over several copy engines (host_copy_split > 1)
  std::vector<cudaStream_t> hx[2];
  std::vector<cudaEvent_t> hxe[2];
  PeerSet* d_peers = nullptr;  // device copy of the peer pointers of the multi-GPU epilogue
  PeerSet h_peers{};           // what d_peers currently holds
  // block-dense tensor-core path: densified tiles of A (side 0) and A^T (side 1), built on first use
  TcStore tc[2];
  bool tc_built[2] = {false, false};
  bool tc_failed[2] = {false, false};  // the build was refused (memory): do not retry on every call
  bool tc_stale[2] = {false, false};   // values changed since the tiles were packed
  void* tc_ws = nullptr;               // [non-finite flag | debug word | packed dense operand]
  size_t tc_ws_bytes = 0;
  int last_path = 0;                   // 0 gather kernels, 1 tensor-core tiles, 2 strided fallback
  SlabStore slab[2];                   // narrow-N slab kernel state per side (built for one N at a time)
  bool slab_built[2] = {false, false};
  bool slab_stale[2] = {false, false};  // values changed since the sliced stream was written
  TcSample samp;                       // A's pattern as bit masks, for the tensor-core SDDMM's sampling epilogue
  bool samp_built = false, samp_failed = false;
  void* sd_ws = nullptr;               // packed G and B of the tensor-core SDDMM
  size_t sd_ws_bytes = 0;
};

// Boundaries depend on (row_ptr, W, heavy, tail split) only: computed once per distinct task shape, kept per side.
static int plan_ensure_bounds(b200_plan* p, int side, const int* row_ptr, int M, int max_row, int64_t N,
                              cudaStream_t s) {
  FwdPick k = fwd_pick(M, p->nnz, N, true, true, max_row);
  p->bnd_valid[side] = false;
  if (p->nnz == 0 || M == 0) return B200_OK;
  for (auto& e : p->bnd_sets[side]) {
    if (e.pick.W == k.W && e.pick.heavy == k.heavy && e.pick.n_tasks == k.n_tasks && e.pick.t1 == k.t1 &&
        e.pick.wf == k.wf) {
      p->bnd[side] = e.part;
      p->pick[side] = e.pick;
      p->bnd_valid[side] = true;
      return B200_OK;
    }
  }
  if (p->bnd_sets[side].size() >= 8) {  // a plan used with many task shapes: drop the oldest set
    B200_CUDA_OK(cudaStreamSynchronize(s));
    B200_CUDA_OK(cudaFree(p->bnd_sets[side].front().part));
    p->bnd_sets[side].erase(p->bnd_sets[side].begin());
  }
  void* part = nullptr;
  B200_CUDA_OK(cudaMalloc(&part, fwd_bounds_bytes(k.n_tasks)));
  int rc = launch_partition(row_ptr, M, p->nnz, k, part, s);
  if (rc != B200_OK) { cudaFree(part); return rc; }
  p->bnd_sets[side].push_back({part, k});
  p->bnd[side] = part;
  p->pick[side] = k;
  p->bnd_valid[side] = true;
  return B200_OK;
}

static int plan_check_device(const b200_plan* p) {
  int dev = -1;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (dev != p->device) {
    set_error("plan was created on device %d but the current device is %d", p->device, dev);
    return B200_ERR_INVALID;
  }
  return B200_OK;
}

static int plan_ensure_ws(b200_plan* p, size_t need) {
  if (p->ws_bytes >= need) return B200_OK;
  if (p->ws) {
    B200_CUDA_OK(cudaDeviceSynchronize());
    B200_CUDA_OK(cudaFree(p->ws));
    p->ws = nullptr;
    p->ws_bytes = 0;
  }
  if (cudaMalloc(&p->ws, need) != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMalloc of %zu plan workspace bytes failed", need);
    return B200_ERR_ALLOC;
  }
  p->ws_bytes = need;
  return B200_OK;
}

static int plan_ensure_transpose(b200_plan* p, cudaStream_t s) {
  if (!p->t_built) {
    B200_CUDA_OK(cudaMalloc(&p->t_row_ptr, ((size_t)p->K + 1) * sizeof(int)));
    const size_t n = std::max<size_t>((size_t)p->nnz, 1);
    B200_CUDA_OK(cudaMalloc(&p->t_col, n * sizeof(int)));
    B200_CUDA_OK(cudaMalloc(&p->t_val, n * sizeof(float)));
    B200_CUDA_OK(cudaMalloc(&p->t_perm, n * sizeof(int)));
    int rc = launch_transpose(p->row_ptr, p->col, p->val, p->M, p->K, p->nnz, p->t_row_ptr, p->t_col,
                              p->t_val, p->t_perm, s);
    if (rc != B200_OK) return rc;
    // longest row of A^T: lets the backward side skip the piece machinery when no row can be split
    int* d_stats = nullptr;
    int h_stats[2] = {0, 0};
    B200_CUDA_OK(cudaMalloc(&d_stats, 2 * sizeof(int)));
    rc = launch_row_stats(p->t_row_ptr, p->K, d_stats, s);
    if (rc == B200_OK) {
      B200_CUDA_OK(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, s));
      B200_CUDA_OK(cudaStreamSynchronize(s));
      p->t_max_row = h_stats[0];
    }
    cudaFree(d_stats);
    if (rc != B200_OK) return rc;
    p->t_built = true;
    p->t_val_stale = false;
  } else if (p->t_val_stale) {
    int rc = launch_gather(p->val, p->t_perm, p->nnz, p->t_val, s);
    if (rc != B200_OK) return rc;
    p->t_val_stale = false;
  }
  return B200_OK;
}

namespace {
__global__ void smoke_kernel() {}

// Column panel of a row-major matrix copied by SMs instead of the copy engine: `rows` rows of `w4` 16-byte vectors,
// both sides with the same pitch.  One side is pinned host memory (mapped into the device address space): the SMs
// issue 16-byte loads / stores straight over PCIe, 128-512 bytes per row segment per warp.  A 2-D DMA with 256-byte
// rows runs well below the link rate; this does not care about the row length.
__global__ void __launch_bounds__(256)
panel_copy_kernel(float4* __restrict__ dst, const float4* __restrict__ src, long long rows, int w4, long long pitch4) {
  const long long total = rows * w4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / w4;
    const int c = (int)(i - r * w4);
    dst[r * pitch4 + c] = src[r * pitch4 + c];
  }
}

// device-visible alias of a pinned host buffer (nullptr for pageable memory)
const void* device_alias(const void* host) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return (a.type == cudaMemoryTypeHost) ? a.devicePointer : nullptr;
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaMalloc of %zu bytes failed", bytes);
      return B200_ERR_ALLOC;
    }
    return B200_OK;
  }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
#define H2D(dst, src, bytes, s) \
  do { if ((bytes) > 0) B200_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s)); } while (0)
#define D2H(dst, src, bytes, s) \
  do { if ((bytes) > 0) B200_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s)); } while (0)
}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* b200_last_error(void) { return tl_error; }
int b200_abi_version(void) { return 1; }
const char* b200_build_info(void) {
  return "spmm_b200 sm_100a fp32 CSR SpMM engine (merge-path tasks, TMA-staged col/val), built " __DATE__;
}
uint64_t b200_kernel_launch_count(void) { return g_launches.load(); }

int b200_smoke_kernel(void* stream, int* sm_count) {
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  smoke_kernel<<<1, 32, 0, s>>>();
  B200_LAUNCH_OK("smoke_kernel");
  B200_CUDA_OK(cudaStreamSynchronize(s));
  if (sm_count) *sm_count = device_sm_count();
  return B200_OK;
}

static int g_trace_idx = 0;  // debug timeline cursor (tc_trace_idx / tc_trace_word)
int b200_set_option(const char* name, int64_t value) {
  B200_REQUIRE(name != nullptr, "option name");
  if (!strcmp(name, "stage_mode")) { options().stage_mode = (int)value; return B200_OK; }
  if (!strcmp(name, "items_per_task")) { options().items_per_task = (int)value; return B200_OK; }
  if (!strcmp(name, "pipeline")) { options().pipeline = (int)value; return B200_OK; }
  if (!strcmp(name, "flat_min_n")) { options().flat_min_n = (int)value; return B200_OK; }
  if (!strcmp(name, "slab_mode")) { options().slab_mode = (int)value; return B200_OK; }
  if (!strcmp(name, "host_copy_mode")) { options().host_copy_mode = (int)value; return B200_OK; }
  if (!strcmp(name, "host_copy_ctas")) { options().host_copy_ctas = (int)value; return B200_OK; }
  if (!strcmp(name, "host_copy_split")) { options().host_copy_split = (int)value; return B200_OK; }
  if (!strcmp(name, "tail_tasks")) { options().tail_tasks = (int)std::max<int64_t>(0, value); return B200_OK; }
  if (!strcmp(name, "d2c_lag")) { options().d2c_lag = (int)value; return B200_OK; }
  if (!strcmp(name, "d2c_warps")) {
    B200_REQUIRE(value == 0 || value == 16 || value == 20 || value == 24, "d2c_warps must be 0, 16, 20 or 24");
    options().d2c_warps = (int)value; return B200_OK;
  }
  if (!strcmp(name, "tc_mode")) { options().tc_mode = (int)value; return B200_OK; }
  if (!strcmp(name, "tc_nt")) {
    B200_REQUIRE(value == 0 || (value % 64 == 0 && value >= 64 && value <= 256), "tc_nt must be 0 or a multiple of 64 up to 256");
    options().tc_nt = (int)value; return B200_OK;
  }
  if (!strcmp(name, "tc_terms")) { options().tc_terms = (int)value & 1023;  /* bits 3-5: timing studies (skip store / B copy / A copy); bits 6-8: SDDMM epilogue studies; bit 9: timeline of CTA 0 */ return B200_OK; }
  if (!strcmp(name, "tc_trace_sync")) { return (value ? b200::tc_trace_sync() : b200::tc_trace_clear()) == 0 ? B200_OK : B200_ERR_CUDA; }
  if (!strcmp(name, "tc_trace_idx")) { g_trace_idx = (int)value; return B200_OK; }
  if (!strcmp(name, "tc_chunk")) { options().tc_chunk = (int)std::max<int64_t>(1, value); return B200_OK; }
  if (!strcmp(name, "tc_kind")) { options().tc_kind = value ? 1 : 0; return B200_OK; }
  if (!strcmp(name, "tc_pair")) { options().tc_pair = (int)std::min<int64_t>(2, std::max<int64_t>(0, value)); return B200_OK; }
  if (!strcmp(name, "tc_sddmm_short")) { options().tc_sddmm_short = (int)(value & 3); return B200_OK; }
  if (!strcmp(name, "tc_min_density_pct")) { options().tc_min_density_pct = (int)value; return B200_OK; }
  set_error("unknown option '%s'", name);
  return B200_ERR_INVALID;
}
int64_t b200_get_option(const char* name) {
  if (!name) return -1;
  if (!strcmp(name, "tc_last_kernel")) return b200::g_tc_last_kernel;
  if (!strcmp(name, "tc_trace_word")) return b200::tc_trace_word(g_trace_idx++);
  if (!strcmp(name, "tc_pair_capacity")) return b200::g_tc_pair_capacity;
  if (!strcmp(name, "stage_mode")) return options().stage_mode;
  if (!strcmp(name, "items_per_task")) return options().items_per_task;
  if (!strcmp(name, "pipeline")) return options().pipeline;
  if (!strcmp(name, "flat_min_n")) return options().flat_min_n;
  if (!strcmp(name, "slab_mode")) return options().slab_mode;
  if (!strcmp(name, "host_copy_mode")) return options().host_copy_mode;
  if (!strcmp(name, "host_copy_ctas")) return options().host_copy_ctas;
  if (!strcmp(name, "host_copy_split")) return options().host_copy_split;
  if (!strcmp(name, "tail_tasks")) return options().tail_tasks;
  if (!strcmp(name, "d2c_warps")) return options().d2c_warps;
  if (!strcmp(name, "d2c_lag")) return options().d2c_lag;
  if (!strcmp(name, "tc_mode")) return options().tc_mode;
  if (!strcmp(name, "tc_nt")) return options().tc_nt;
  if (!strcmp(name, "tc_terms")) return options().tc_terms;
  if (!strcmp(name, "tc_chunk")) return options().tc_chunk;
  if (!strcmp(name, "tc_kind")) return options().tc_kind;
  if (!strcmp(name, "tc_pair")) return options().tc_pair;
  if (!strcmp(name, "tc_sddmm_short")) return options().tc_sddmm_short;
  if (!strcmp(name, "tc_min_density_pct")) return options().tc_min_density_pct;
  return -1;
}

size_t b200_spmm_workspace_bytes(int64_t M, int64_t nnz, int64_t N) {
  if (M < 0 || nnz < 0 || N < 0) return 0;
  return fwd_workspace_bytes(M, nnz, N);
}

int b200_spmm_csr_forward(const int32_t* d_row_ptr, const int32_t* d_col_ind, const float* d_val,
                          int64_t M, int64_t K, int64_t nnz, const float* d_B, int64_t ldb, int64_t N,
                          float* d_C, int64_t ldc, void* d_workspace, size_t workspace_bytes,
                          void* stream) {
  return b200_spmm_csr_forward_flags(d_row_ptr, d_col_ind, d_val, M, K, nnz, d_B, ldb, N, d_C, ldc,
                                     d_workspace, workspace_bytes, stream, 0u);
}

int b200_spmm_csr_forward_flags(const int32_t* d_row_ptr, const int32_t* d_col_ind, const float* d_val,
                                int64_t M, int64_t K, int64_t nnz, const float* d_B, int64_t ldb,
                                int64_t N, float* d_C, int64_t ldc, void* d_workspace,
                                size_t workspace_bytes, void* stream, uint32_t flags) {
  int rc = check_csr_dims(M, K, nnz);
  if (rc) return rc;
  B200_REQUIRE(N >= 0 && N < 0x7fffffffLL, "N out of range");
  B200_REQUIRE(ldb >= N && ldc >= N, "leading dimension smaller than N");
  B200_REQUIRE(ldb < 0x7fffffffLL && ldc < 0x7fffffffLL, "leading dimension out of range");
  if (M == 0 || N == 0) return B200_OK;
  B200_REQUIRE(d_row_ptr != nullptr && d_C != nullptr, "null row_ptr / C");
  B200_REQUIRE(nnz == 0 || (d_col_ind != nullptr && d_val != nullptr && d_B != nullptr), "null col/val/B");
  FwdParams p{};
  p.row_ptr = d_row_ptr; p.col = d_col_ind; p.val = d_val;
  p.M = (int)M; p.nnz = (int)nnz;
  p.B = d_B; p.ldb = ldb; p.N = (int)N; p.C = d_C; p.ldc = ldc;
  p.stage = (flags & B200_FLAG_NO_STAGE) ? 0 : -1;
  return launch_spmm_forward(p, d_workspace, workspace_bytes, nullptr, nullptr, (cudaStream_t)stream);
}

int b200_sddmm_csr(const int32_t* d_row_ptr, const int32_t* d_col_ind, int64_t M, int64_t K, int64_t nnz,
                   const float* d_G, int64_t ldg, const float* d_B, int64_t ldb, int64_t N,
                   float* d_out_val, void* stream) {
  int rc = check_csr_dims(M, K, nnz);
  if (rc) return rc;
  B200_REQUIRE(N >= 0 && N < 0x7fffffffLL, "N out of range");
  B200_REQUIRE(ldg >= N && ldb >= N, "leading dimension smaller than N");
  if (nnz == 0) return B200_OK;
  B200_REQUIRE(d_row_ptr && d_col_ind && d_out_val, "null CSR / output pointer");
  B200_REQUIRE(N == 0 || (d_G && d_B), "null G / B");
  B200_REQUIRE(ldg < 0x7fffffffLL && ldb < 0x7fffffffLL, "leading dimension out of range");
  return launch_sddmm(d_row_ptr, d_col_ind, (int)M, (int)nnz, d_G, ldg, d_B, ldb, (int)N, d_out_val, nullptr,
                      (cudaStream_t)stream);
}

int b200_csr_transpose(const int32_t* d_row_ptr, const int32_t* d_col_ind, const float* d_val, int64_t M,
                       int64_t K, int64_t nnz, int32_t* d_t_row_ptr, int32_t* d_t_col_ind,
                       float* d_t_val, int32_t* d_t_perm, void* stream) {
  int rc = check_csr_dims(M, K, nnz);
  if (rc) return rc;
  B200_REQUIRE(d_t_row_ptr != nullptr, "null output row_ptr");
  B200_REQUIRE(nnz == 0 || (d_row_ptr && d_col_ind && d_t_col_ind), "null CSR pointer");
  B200_REQUIRE(d_val == nullptr || nnz == 0 || d_t_val != nullptr, "null output values");
  return launch_transpose(d_row_ptr, d_col_ind, d_val, (int)M, (int)K, (int)nnz, d_t_row_ptr,
                          d_t_col_ind, d_t_val, d_t_perm, (cudaStream_t)stream);
}

int b200_gather_f32(const float* d_src, const int32_t* d_perm, int64_t n, float* d_dst, void* stream) {
  B200_REQUIRE(n >= 0, "negative length");
  if (n == 0) return B200_OK;
  B200_REQUIRE(d_src && d_perm && d_dst, "null pointer");
  return launch_gather(d_src, d_perm, n, d_dst, (cudaStream_t)stream);
}

int b200_dense_to_csr_count(const float* d_dense, int64_t M, int64_t K, int64_t ld, int32_t* d_row_ptr,
                            int64_t* h_nnz, void* stream) {
  B200_REQUIRE(M >= 0 && K >= 0 && M < 0x7fffffffLL && K < 0x7fffffffLL, "dimension out of range");
  B200_REQUIRE(ld >= K, "leading dimension smaller than K");
  B200_REQUIRE(d_row_ptr != nullptr, "null row_ptr");
  B200_REQUIRE(M == 0 || K == 0 || d_dense != nullptr, "null dense");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_dense_count(d_dense, (int)M, (int)K, ld, d_row_ptr, s);
  if (rc) return rc;
  if (h_nnz) {
    int n = 0;
    B200_CUDA_OK(cudaMemcpyAsync(&n, d_row_ptr + M, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    *h_nnz = n;
  }
  return B200_OK;
}

int b200_dense_to_csr_fill(const float* d_dense, int64_t M, int64_t K, int64_t ld,
                           const int32_t* d_row_ptr, int32_t* d_col_ind, float* d_val, void* stream) {
  B200_REQUIRE(M >= 0 && K >= 0 && M < 0x7fffffffLL && K < 0x7fffffffLL, "dimension out of range");
  B200_REQUIRE(ld >= K, "leading dimension smaller than K");
  if (M == 0 || K == 0) return B200_OK;
  B200_REQUIRE(d_dense && d_row_ptr, "null dense / row_ptr");
  return launch_dense_fill(d_dense, (int)M, (int)K, ld, d_row_ptr, d_col_ind, d_val, (cudaStream_t)stream);
}

size_t b200_dense_to_csr_onepass_scratch_bytes(int64_t M, int64_t K) {
  if (M <= 0 || K <= 0 || M >= 0x7fffffffLL || K >= 0x7fffffffLL) return 16;
  return dense_onepass_scratch_bytes((int)M, (int)K);
}

int b200_dense_to_csr_onepass(const float* d_dense, int64_t M, int64_t K, int64_t ld, int32_t* d_row_ptr,
                              int32_t* d_col_ind, float* d_val, int64_t capacity, void* d_scratch,
                              size_t scratch_bytes, int64_t* h_nnz, void* stream) {
  B200_REQUIRE(M >= 0 && K >= 0 && M < 0x7fffffffLL && K < 0x7fffffffLL, "dimension out of range");
  B200_REQUIRE(ld >= K, "leading dimension smaller than K");
  B200_REQUIRE(d_row_ptr != nullptr && capacity >= 0, "null row_ptr / negative capacity");
  B200_REQUIRE(capacity == 0 || (d_col_ind != nullptr && d_val != nullptr), "null col / val");
  B200_REQUIRE(M == 0 || K == 0 || d_dense != nullptr, "null dense");
  cudaStream_t s = (cudaStream_t)stream;
  if (!dense_onepass_supported(d_dense, (int)M, (int)K, ld)) {
    // unaligned rows or very narrow matrices: the two-pass kernels (the matrix is small or oddly shaped anyway)
    int64_t n = 0;
    int rc = b200_dense_to_csr_count(d_dense, M, K, ld, d_row_ptr, &n, stream);
    if (rc) return rc;
    if (h_nnz) *h_nnz = n;
    if (n > capacity) {
      set_error("capacity %lld too small for %lld non-zeros (row_ptr is complete)", (long long)capacity, (long long)n);
      return B200_ERR_INVALID;
    }
    return b200_dense_to_csr_fill(d_dense, M, K, ld, d_row_ptr, d_col_ind, d_val, stream);
  }
  const size_t need = dense_onepass_scratch_bytes((int)M, (int)K);
  void* scratch = d_scratch;
  DevBuf own;
  if (scratch == nullptr) {  // convenience: callers on a hot path pass their own scratch (no allocation here)
    int rc = own.alloc(need);
    if (rc) return rc;
    scratch = own.p;
  } else {
    B200_REQUIRE(scratch_bytes >= need, "one-pass scratch too small (b200_dense_to_csr_onepass_scratch_bytes)");
    B200_REQUIRE(reinterpret_cast<uintptr_t>(scratch) % 16 == 0, "one-pass scratch must be 16-byte aligned");
  }
  int rc = launch_dense_onepass(d_dense, (int)M, (int)K, ld, d_row_ptr, d_col_ind, d_val, capacity, scratch, s);
  int n = 0;
  if (rc == B200_OK && (h_nnz || own.p)) {
    B200_CUDA_OK(cudaMemcpyAsync(&n, d_row_ptr + M, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));  // (also keeps a library-owned scratch alive until the kernel is done)
  }
  if (rc) return rc;
  if (h_nnz) {
    *h_nnz = n;
    if (n > capacity) {
      set_error("capacity %lld too small for %lld non-zeros (row_ptr is complete: allocate and call b200_dense_to_csr_fill)",
                (long long)capacity, (long long)n);
      return B200_ERR_INVALID;
    }
  }
  return B200_OK;
}

int b200_dense_to_csr_alloc(const float* d_dense, int64_t M, int64_t K, int64_t ld, float** d_val,
                            int32_t** d_col_ind, int32_t** d_row_ptr, int64_t* nnz, void* stream) {
  B200_REQUIRE(d_val && d_col_ind && d_row_ptr && nnz, "null output pointer");
  B200_REQUIRE(M >= 0 && K >= 0 && M < 0x7fffffffLL && K < 0x7fffffffLL, "dimension out of range");
  *d_val = nullptr; *d_col_ind = nullptr; *d_row_ptr = nullptr; *nnz = 0;
  B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(d_row_ptr), ((size_t)M + 1) * sizeof(int)));
  int rc = b200_dense_to_csr_count(d_dense, M, K, ld, *d_row_ptr, nnz, stream);
  if (rc) return rc;
  const size_t n = std::max<size_t>((size_t)*nnz, 1);
  B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(d_col_ind), n * sizeof(int)));
  B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(d_val), n * sizeof(float)));
  rc = b200_dense_to_csr_fill(d_dense, M, K, ld, *d_row_ptr, *d_col_ind, *d_val, stream);
  if (rc) return rc;
  B200_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  return B200_OK;
}

int b200_free_csr(float* d_val, int32_t* d_col_ind, int32_t* d_row_ptr) {
  B200_CUDA_OK(cudaFree(d_val));
  B200_CUDA_OK(cudaFree(d_col_ind));
  B200_CUDA_OK(cudaFree(d_row_ptr));
  return B200_OK;
}

int b200_csr_block_diag_cols(const int32_t* d_row_ptr, int32_t* d_col_ind, int64_t rows, int64_t M, int64_t K,
                             void* stream) {
  B200_REQUIRE(rows >= 0 && M > 0 && K >= 0 && rows < 0x7fffffffLL, "bad dimensions");
  B200_REQUIRE(rows % M == 0, "rows must be a multiple of M");
  B200_REQUIRE((rows / M) * K < 0x7fffffffLL, "L * K exceeds the int32 column range");
  if (rows == 0) return B200_OK;
  B200_REQUIRE(d_row_ptr != nullptr && d_col_ind != nullptr, "null CSR pointer");
  return launch_block_diag(d_row_ptr, d_col_ind, (int)rows, (int)M, (int)K, (cudaStream_t)stream);
}

// ---- plans ---------------------------------------------------------------------------
int b200_plan_create(const int32_t* d_row_ptr, const int32_t* d_col_ind, const float* d_val, int64_t M,
                     int64_t K, int64_t nnz, b200_plan** out_plan, void* stream) {
  B200_REQUIRE(out_plan != nullptr, "null out_plan");
  *out_plan = nullptr;
  int rc = check_device();
  if (rc) return rc;
  rc = check_csr_dims(M, K, nnz);
  if (rc) return rc;
  B200_REQUIRE(d_row_ptr != nullptr, "null row_ptr");
  B200_REQUIRE(nnz == 0 || (d_col_ind && d_val), "null col/val");
  b200_plan* p = new (std::nothrow) b200_plan();
  if (!p) { set_error("out of host memory"); return B200_ERR_ALLOC; }
  p->row_ptr = d_row_ptr; p->col = d_col_ind; p->val = d_val;
  p->M = (int)M; p->K = (int)K; p->nnz = (int)nnz;
  if (cudaGetDevice(&p->device) != cudaSuccess) { cudaGetLastError(); p->device = 0; }
  cudaStream_t s = (cudaStream_t)stream;
  int* d_stats = nullptr;
  int h_stats[2] = {0, 0};
  if (cudaMalloc(&d_stats, 2 * sizeof(int)) != cudaSuccess) {
    cudaGetLastError(); delete p; set_error("cudaMalloc failed"); return B200_ERR_ALLOC;
  }
  rc = launch_row_stats(d_row_ptr, (int)M, d_stats, s);
  if (rc == B200_OK) {
    cudaError_t e = cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error("row stats copy failed: %s", cudaGetErrorString(e)); rc = B200_ERR_CUDA; }
  }
  cudaFree(d_stats);
  if (rc) { delete p; return rc; }
  p->max_row = h_stats[0];
  p->empty_rows = h_stats[1];
  *out_plan = p;
  return B200_OK;
}

int b200_plan_destroy(b200_plan* p) {
  if (!p) return B200_OK;
  cudaFree(p->t_row_ptr); cudaFree(p->t_col); cudaFree(p->t_val); cudaFree(p->t_perm); cudaFree(p->ws);
  for (int side = 0; side < 2; ++side)
    for (auto& e : p->bnd_sets[side]) cudaFree(e.part);
  cudaFree(p->sd_bnd); cudaFree(p->d_peers);
  cudaFree(p->cnt[0]); cudaFree(p->cnt[1]);
  cudaFree(p->hB); cudaFree(p->hC); cudaFree(p->tc_ws); cudaFree(p->sd_ws);
  tc_sample_free(&p->samp);
  slab_store_free(&p->slab[0]); slab_store_free(&p->slab[1]);
  tc_store_free(&p->tc[0]); tc_store_free(&p->tc[1]);
  for (int i = 0; i < 3; ++i) if (p->hs[i]) cudaStreamDestroy(p->hs[i]);
  for (int d = 0; d < 2; ++d) {
    for (auto st : p->hx[d]) cudaStreamDestroy(st);
    for (auto ev : p->hxe[d]) cudaEventDestroy(ev);
  }
  for (int j = 0; j < 2; ++j) for (int i = 0; i < 16; ++i) if (p->hev[j][i]) cudaEventDestroy(p->hev[j][i]);
  cudaGetLastError();
  delete p;
  return B200_OK;
}

int b200_plan_values_updated(b200_plan* p) {
  B200_REQUIRE(p != nullptr, "null plan");
  p->t_val_stale = true;
  p->tc_stale[0] = p->tc_stale[1] = true;
  p->slab_stale[0] = p->slab_stale[1] = true;
  return B200_OK;
}

int b200_plan_stats(const b200_plan* p, int64_t* max_row_nnz, int64_t* empty_rows) {
  B200_REQUIRE(p != nullptr, "null plan");
  if (max_row_nnz) *max_row_nnz = p->max_row;
  if (empty_rows) *empty_rows = p->empty_rows;
  return B200_OK;
}

static int plan_spmm(b200_plan* p, int side, const int* row_ptr, const int* col, const float* val, int M,
                     int K, int max_row, const float* d_B, int64_t ldb, int64_t N, float* d_C, int64_t ldc,
                     cudaStream_t s, float* const* peer_C = nullptr, int n_peers = 0, float* mc_C = nullptr,
                     const float* d_bias = nullptr, int relu = 0) {
  B200_REQUIRE(N >= 0 && N < 0x7fffffffLL, "N out of range");
  B200_REQUIRE(ldb >= N && ldc >= N, "leading dimension smaller than N");
  B200_REQUIRE(ldb < 0x7fffffffLL && ldc < 0x7fffffffLL, "leading dimension out of range");
  if (M == 0 || N == 0) return B200_OK;
  B200_REQUIRE(d_C != nullptr && (p->nnz == 0 || d_B != nullptr), "null B / C");
  int rc = plan_check_device(p);
  if (rc) return rc;
  rc = plan_ensure_bounds(p, side, row_ptr, M, max_row, N, s);
  if (rc) return rc;
  rc = plan_ensure_ws(p, fwd_workspace_bytes(M, p->nnz, N));  // exact for this side (the task size depends on M + nnz)
  if (rc) return rc;
  FwdParams q{};
  q.row_ptr = row_ptr; q.col = col; q.val = val; q.M = M; q.nnz = p->nnz;
  q.B = d_B; q.ldb = ldb; q.N = (int)N; q.C = d_C; q.ldc = ldc; q.stage = -1;
  q.bias = d_bias; q.relu = relu;
  if (p->bnd_valid[side] && p->pick[side].may_split) {
    // the counters return to zero at the end of every launch, so one zeroed set serves every call; it only has
    // to be large enough for the widest slab count seen (vector width 1: one slab per 32 columns at least 4 wide)
    FwdPick k = p->pick[side];
    k.slabs = (int)((N + 3) / 4 + 1);
    const size_t need = fwd_counter_bytes(k);
    if (p->cnt_bytes[side] < need) {
      if (p->cnt[side]) { B200_CUDA_OK(cudaStreamSynchronize(s)); B200_CUDA_OK(cudaFree(p->cnt[side])); p->cnt[side] = nullptr; p->cnt_bytes[side] = 0; }
      B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&p->cnt[side]), need));
      B200_CUDA_OK(cudaMemsetAsync(p->cnt[side], 0, need, s));
      p->cnt_bytes[side] = need;
    }
    q.counters = p->cnt[side];
  }
  const PeerSet* d_peers = nullptr;
  if (n_peers > 0 || mc_C != nullptr) {
    PeerSet ps{};
    for (int i = 0; i < n_peers; ++i) ps.peer[i] = peer_C[i];
    ps.n_peers = n_peers;
    ps.mc = mc_C;
    if (p->d_peers == nullptr) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&p->d_peers), sizeof(PeerSet)));
    if (memcmp(&ps, &p->h_peers, sizeof(PeerSet)) != 0) {  // uploaded once per set of peers, not per call
      B200_CUDA_OK(cudaMemcpyAsync(p->d_peers, &ps, sizeof(PeerSet), cudaMemcpyHostToDevice, s));
      B200_CUDA_OK(cudaStreamSynchronize(s));
      p->h_peers = ps;
    }
    d_peers = p->d_peers;
  }
  return launch_spmm_forward(q, p->ws, p->ws_bytes, p->bnd_valid[side] ? p->bnd[side] : nullptr,
                             p->bnd_valid[side] ? &p->pick[side] : nullptr, s, d_peers);
}

// Should this product run on the block-dense tensor-core path?  ("crossover rule")
// Auto mode compares two measured cost models (B200, profiles/r2_tc_crossover.md):
//   gather kernels : 12 us + nnz * N / 5.0e6 us   (one gathered B element per FMA; ~2e6 per us below 24 columns)
//   tensor tiles   : 28 us + max(streaming the densified tiles, the MMAs, a CTA's serial walk over its k-tiles):
//                      tile bytes (x column blocks when they exceed L2) / 5.5e6 B per us,
//                      6 * M_pad * K * N_pad / R, R = 1.3e9 (fp16 split) or 0.65e9 (tf32 split) flop per us,
//                      waves * k-tiles * (0.4 + 0.8 Nt / 256) us
// so small problems stay on the gather kernels (launch-bound: config 1-2), and large ones switch at ~4-5 % density.
static double g_last_tc_estimate_us = 0.0;  // the tiles' estimated time of the last plan_wants_tc() call (same thread)
static bool plan_wants_tc(const b200_plan* p, int M, int K, int64_t N, uint32_t flags, bool sddmm = false) {
  g_last_tc_estimate_us = 1e30;
  if (flags & B200_OP_NO_TC) return false;
  const int mode = options().tc_mode;
  if (mode == 0 && !(flags & B200_OP_FORCE_TC)) return false;
  if (M <= 0 || K <= 0 || p->nnz <= 0 || N <= 0) return false;
  const int64_t MB = (M + 127) / 128, KT = (K + 31) / 32;
  if (N < 8) return false;
  if (MB * KT > ((int64_t)1 << 26)) return false;
  if (std::min<int64_t>(MB * KT, p->nnz) * 32768 > ((int64_t)24 << 30)) return false;  // densified tiles
  if ((flags & B200_OP_FORCE_TC) || mode == 2) return true;
  const double density = (double)p->nnz / ((double)M * (double)K);
  if (density * 100.0 < (double)options().tc_min_density_pct) return false;
  const double n_pad = (double)((N + 63) / 64 * 64);
  const double t_gather = 12.0 + (double)p->nnz * (double)N / (N < 24 ? 2.0e6 : 5.0e6);
  const bool f16 = options().tc_kind == 1;
  const double tile_bytes = (double)(MB * 128) * (double)K * (f16 ? 4.0 : 8.0);
  const double passes = tile_bytes > 64e6 ? (double)((N + 255) / 256) : 1.0;
  const double t_stream = tile_bytes * passes / 5.5e6;
  const double t_mma = 6.0 * (double)(MB * 128) * (double)K * n_pad / (f16 ? 1.3e9 : 0.65e9);
  if (sddmm) {
    // grad of the values: every output tile of the M x K product is computed and sampled, so the tile count (not the
    // contraction length N) dominates.  Measured on 16384^2 with the transposed-product kernel (profiles/
    // r2_sddmm_short_probe.txt): 0.20 / 0.20 / 0.23 / 0.31 ms at N = 16 / 64 / 128 / 256 and 90 % sparsity (0.23 / 0.24 /
    // 0.25 / 0.37 at 50 %) -> 30 us + M_pad * K / 1.7e6 for the sampled stores + 0.45 x the MMAs (they overlap the
    // stores of the previous tiles); the gather SDDMM runs at ~3.3e6 nnz*N per us (1.25e6 below 24 columns)
    const double t_g = 12.0 + (double)p->nnz * (double)N / (N < 24 ? 1.25e6 : 3.3e6);
    const double t_t = 30.0 + (double)(MB * 128) * (double)K / 1.7e6 + 0.45 * t_mma;
    return t_t < t_g;
  }
  // a CTA walks the k-tiles of its tiles one after the other: ~0.4 us of stage round trip per k-tile (after the
  // stages that fit have been divided in) + the 12 MMAs (0.8 us at Nt = 256), times the tiles a CTA gets
  // (tools/tc_stream_probe.py; config 3 fc1: 3 waves x 12 x 1.2 = 43 us, config 5 x 256: 256 x 1.2 = 307 us)
  const int sms = device_sm_count();
  const int nt = options().tc_nt > 0 ? options().tc_nt : tc_pick_nt((int)MB, (int)std::min<int64_t>(N, 1 << 20), sms);
  const double waves = std::ceil((double)MB * (double)((N + nt - 1) / nt) / (double)std::max(1, sms));
  const double k_tiles = (double)((K + (f16 ? 63 : 31)) / (f16 ? 64 : 32));
  const double t_chain = waves * k_tiles * (0.4 + 0.8 * (double)nt / 256.0);
  const double t_tc = 28.0 + std::max(std::max(t_stream, t_mma), t_chain);
  g_last_tc_estimate_us = t_tc;
  return t_tc < t_gather;
}

static int plan_ensure_tc(b200_plan* p, int side, const int* row_ptr, const int* col, const float* val, int M,
                          int K, cudaStream_t s) {
  if (p->tc_failed[side]) return B200_ERR_ALLOC;
  if (!p->tc_built[side] || p->tc[side].kind != options().tc_kind) {
    if (p->tc_built[side]) tc_store_free(&p->tc[side]);
    p->tc_built[side] = false;
    int rc = tc_store_build(row_ptr, col, val, M, K, p->nnz, options().tc_kind, &p->tc[side], s);
    if (rc != B200_OK) {
      p->tc_failed[side] = true;
      return rc;
    }
    p->tc_built[side] = true;
    p->tc_stale[side] = false;
  } else if (p->tc_stale[side]) {
    int rc = tc_store_refresh(row_ptr, col, val, p->nnz, &p->tc[side], s);
    if (rc != B200_OK) return rc;
    p->tc_stale[side] = false;
  }
  return B200_OK;
}

// Extended product over one side of the plan (A or the cached A^T): optional transposed dense operand /
// transposed result, bias per output row and ReLU fused into the epilogue, tensor-core tiles when the
// operand is dense enough.
static int plan_spmm_ex(b200_plan* p, int side, const int* row_ptr, const int* col, const float* val, int M, int K,
                        int max_row, const float* d_B, int64_t ldb, int64_t N, float* d_C, int64_t ldc,
                        const float* d_bias, uint32_t flags, cudaStream_t s) {
  const bool bt = (flags & B200_OP_B_TRANSPOSED) != 0, ct = (flags & B200_OP_C_TRANSPOSED) != 0;
  const bool relu = (flags & B200_OP_RELU) != 0;
  B200_REQUIRE(N >= 0 && N < 0x7fffffffLL, "N out of range");
  B200_REQUIRE(ldb >= (bt ? K : N) && ldc >= (ct ? M : N), "leading dimension too small");
  B200_REQUIRE(ldb < 0x7fffffffLL && ldc < 0x7fffffffLL, "leading dimension out of range");
  if (M == 0 || N == 0) return B200_OK;
  B200_REQUIRE(d_C != nullptr && (p->nnz == 0 || d_B != nullptr), "null B / C");
  int rc = plan_check_device(p);
  if (rc) return rc;
  bool want_tc = plan_wants_tc(p, M, K, N, flags);
  // narrow dense operand whose rows fit a few shared-memory slabs: ~15 us + nnz / 4.0e5 us on the slab kernel
  const bool slab_ok = !bt && !ct && options().slab_mode >= 1 && !(flags & B200_OP_FORCE_TC) &&
                       slab_wanted(M, K, p->nnz, (int)N, d_B, ldb, d_C, ldc);
  if (slab_ok && want_tc && 15.0 + (double)p->nnz / 4.0e5 < g_last_tc_estimate_us) want_tc = false;
  if (want_tc) {
    rc = plan_ensure_tc(p, side, row_ptr, col, val, M, K, s);
    if (rc == B200_OK) {
      const int nt = options().tc_nt > 0 ? options().tc_nt : tc_pick_nt(p->tc[side].MB, (int)N, device_sm_count());
      const size_t need = tc_pack_b_bytes(K, (int)N, nt, p->tc[side].kind);
      if (p->tc_ws_bytes < need) {
        if (p->tc_ws) { B200_CUDA_OK(cudaStreamSynchronize(s)); B200_CUDA_OK(cudaFree(p->tc_ws)); p->tc_ws = nullptr; p->tc_ws_bytes = 0; }
        if (cudaMalloc(&p->tc_ws, need) != cudaSuccess) {
          cudaGetLastError();
          set_error("cudaMalloc of %zu bytes for the packed dense operand failed", need);
          return B200_ERR_ALLOC;
        }
        p->tc_ws_bytes = need;
        B200_CUDA_OK(cudaMemsetAsync(p->tc_ws, 0, 256, s));  // control words: the kernels keep them zero afterwards
      }
      // (Inf/NaN in the dense operand are found by the packer: that launch then runs in CSR order, no host sync)
      rc = launch_tc_forward(p->tc[side], row_ptr, col, val, d_B, ldb, (int)N, d_C, ldc, d_bias, flags, p->tc_ws,
                             p->tc_ws_bytes, nt, s);
      p->last_path = 1;
      return rc;
    }
    if (flags & B200_OP_FORCE_TC) return rc;  // otherwise fall through to the gather kernels
  }
  if (bt || ct) {  // transposed operands on a matrix too sparse for the tiles: plain CSR-order kernel
    p->last_path = 2;
    return launch_spmm_strided(row_ptr, col, val, M, (int)N, d_B, bt ? 1 : ldb, bt ? ldb : 1, d_C, ct ? 1 : ldc,
                               ct ? ldc : 1, d_bias, relu, s);
  }
  if (slab_ok) {
    // gather from shared-memory slabs of B instead of L1 / L2
    if (!p->slab_built[side] || p->slab[side].N != (int)N || p->slab[side].v2 != (options().slab_mode != 2)) {
      if (p->slab_built[side]) { B200_CUDA_OK(cudaStreamSynchronize(s)); slab_store_free(&p->slab[side]); p->slab_built[side] = false; }
      rc = slab_store_build(row_ptr, col, val, M, K, p->nnz, (int)N, &p->slab[side], s);
      if (rc) return rc;
      p->slab_built[side] = true;
      p->slab_stale[side] = false;
    } else if (p->slab_stale[side]) {  // values changed in place: rewrite the stream
      rc = slab_store_refresh(col, val, &p->slab[side], s);
      if (rc) return rc;
      p->slab_stale[side] = false;
    }
    p->last_path = 3;
    return launch_spmm_slab(p->slab[side], col, val, d_B, d_C, ldc, d_bias, relu ? 1 : 0, s);
  }
  p->last_path = 0;
  return plan_spmm(p, side, row_ptr, col, val, M, K, max_row, d_B, ldb, N, d_C, ldc, s, nullptr, 0, nullptr, d_bias,
                   relu ? 1 : 0);
}

int b200_plan_forward(b200_plan* p, const float* d_B, int64_t ldb, int64_t N, float* d_C, int64_t ldc,
                      void* stream) {
  B200_REQUIRE(p != nullptr, "null plan");
  return plan_spmm_ex(p, 0, p->row_ptr, p->col, p->val, p->M, p->K, p->max_row, d_B, ldb, N, d_C, ldc, nullptr, 0u,
                      (cudaStream_t)stream);
}

int b200_plan_forward_ex(b200_plan* p, const float* d_B, int64_t ldb, int64_t N, float* d_C, int64_t ldc,
                         const float* d_bias, uint32_t flags, void* stream) {
  B200_REQUIRE(p != nullptr, "null plan");
  return plan_spmm_ex(p, 0, p->row_ptr, p->col, p->val, p->M, p->K, p->max_row, d_B, ldb, N, d_C, ldc, d_bias,
                      flags, (cudaStream_t)stream);
}

int b200_plan_backward_B_ex(b200_plan* p, const float* d_G, int64_t ldg, int64_t N, float* d_dB, int64_t lddb,
                            uint32_t flags, void* stream) {
  B200_REQUIRE(p != nullptr, "null plan");
  B200_REQUIRE((flags & B200_OP_RELU) == 0, "ReLU is a forward epilogue");
  int rc = plan_ensure_transpose(p, (cudaStream_t)stream);
  if (rc) return rc;
  return plan_spmm_ex(p, 1, p->t_row_ptr, p->t_col, p->t_val, p->K, p->M, p->t_max_row, d_G, ldg, N, d_dB, lddb,
                      nullptr, flags, (cudaStream_t)stream);
}

int b200_plan_would_use_tc(const b200_plan* p, int side, int64_t N, uint32_t flags) {
  if (p == nullptr || side < 0 || side > 2) return 0;
  if (side == 2) return (!p->samp_failed && plan_wants_tc(p, p->M, p->K, N, flags, true)) ? 1 : 0;
  if (p->tc_failed[side]) return 0;
  return plan_wants_tc(p, side == 0 ? p->M : p->K, side == 0 ? p->K : p->M, N, flags) ? 1 : 0;
}

int b200_plan_tc_info(const b200_plan* p, int side, int64_t* n_tiles, int64_t* tile_bytes, int* last_path) {
  B200_REQUIRE(p != nullptr && (side == 0 || side == 1), "null plan / bad side");
  if (n_tiles) *n_tiles = p->tc_built[side] ? p->tc[side].n_tiles : 0;
  if (tile_bytes) *tile_bytes = p->tc_built[side] ? (int64_t)p->tc[side].tile_bytes : 0;
  if (last_path) *last_path = p->last_path;
  return B200_OK;
}

int b200_tc_last_timeout(const b200_plan* p, int* code) {
  B200_REQUIRE(p != nullptr && code != nullptr, "null plan / code");
  *code = 0;
  if (p->tc_ws != nullptr) {
    int h[2] = {0, 0};
    B200_CUDA_OK(cudaMemcpy(h, p->tc_ws, sizeof(h), cudaMemcpyDeviceToHost));
    *code = h[1];
  }
  return B200_OK;
}

int b200_plan_forward_host(b200_plan* p, const float* h_B, int64_t N, float* h_C, int panels) {
  B200_REQUIRE(p != nullptr, "null plan");
  B200_REQUIRE(N >= 0 && N < 0x7fffffffLL, "N out of range");
  if (p->M == 0 || N == 0) return B200_OK;
  B200_REQUIRE(h_C != nullptr && (p->nnz == 0 || h_B != nullptr), "null host pointer");
  // panels: multiples of 32 columns (one 128-byte line per B row and panel at least), at most 16
  // auto: panels of >= 64 columns (narrower 2-D PCIe copies lose more than the extra overlap gains; measured
  // with tools/host_pipeline_check.py: 1 / 2 / 4 panels of a 128-column operand = 4.97 / 3.96 / 4.45 ms)
  if (panels <= 0) panels = (int)std::min<int64_t>(4, std::max<int64_t>(1, N / 64));
  while (panels > 1 && (N % panels != 0 || (N / panels) % 32 != 0)) --panels;
  if (panels > 16) panels = 16;
  const int64_t w = N / panels;
  const size_t bB = (size_t)p->K * N * sizeof(float), bC = (size_t)p->M * N * sizeof(float);
  if (p->hB_bytes < bB) {
    if (p->hB) { B200_CUDA_OK(cudaDeviceSynchronize()); B200_CUDA_OK(cudaFree(p->hB)); p->hB = nullptr; p->hB_bytes = 0; }
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&p->hB), std::max<size_t>(bB, 16)));
    p->hB_bytes = bB;
  }
  if (p->hC_bytes < bC) {
    if (p->hC) { B200_CUDA_OK(cudaDeviceSynchronize()); B200_CUDA_OK(cudaFree(p->hC)); p->hC = nullptr; p->hC_bytes = 0; }
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&p->hC), std::max<size_t>(bC, 16)));
    p->hC_bytes = bC;
  }
  for (int i = 0; i < 3; ++i)
    if (!p->hs[i]) B200_CUDA_OK(cudaStreamCreateWithFlags(&p->hs[i], cudaStreamNonBlocking));
  for (int j = 0; j < 2; ++j)
    for (int i = 0; i < panels; ++i)
      if (!p->hev[j][i]) B200_CUDA_OK(cudaEventCreateWithFlags(&p->hev[j][i], cudaEventDisableTiming));
  const size_t pitch = (size_t)N * sizeof(float);
  // SM-driven panel copies when both host buffers are pinned (mapped) and 16-byte friendly; the copy engine otherwise
  const float* dB = reinterpret_cast<const float*>(device_alias(h_B));
  float* dC = reinterpret_cast<float*>(const_cast<void*>(device_alias(h_C)));
  const bool sm_copy = options().host_copy_mode == 1 && panels > 1 && dB != nullptr && dC != nullptr && N % 4 == 0 &&
                       w % 4 == 0 && reinterpret_cast<uintptr_t>(dB) % 16 == 0 && reinterpret_cast<uintptr_t>(dC) % 16 == 0;
  const int copy_ctas = std::max(16, options().host_copy_ctas);
  // a 2-D copy with short rows is limited by one copy engine, not by PCIe: split every panel copy by rows over
  // `split` streams so that several engines work on it
  const int split = (!sm_copy && panels > 1) ? std::max(1, std::min(8, options().host_copy_split)) : 1;
  if (split > 1) {
    for (int d = 0; d < 2; ++d) {
      while ((int)p->hx[d].size() < split - 1) {
        cudaStream_t st;
        B200_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        p->hx[d].push_back(st);
      }
      while ((int)p->hxe[d].size() < (split - 1) * panels + panels) {
        cudaEvent_t ev;
        B200_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        p->hxe[d].push_back(ev);
      }
    }
  }
  for (int i = 0; i < panels; ++i) {
    const size_t off = (size_t)i * w;
    if (p->nnz > 0 && split > 1) {
      // rows [r0, r1) of the panel on stream c (c = 0: the main upload stream)
      for (int c = 0; c < split; ++c) {
        const size_t r0 = (size_t)p->K * c / split, r1 = (size_t)p->K * (c + 1) / split;
        cudaStream_t st = c == 0 ? p->hs[0] : p->hx[0][c - 1];
        if (r1 > r0)
          B200_CUDA_OK(cudaMemcpy2DAsync(p->hB + r0 * N + off, pitch, h_B + r0 * N + off, pitch, (size_t)w * sizeof(float),
                                         r1 - r0, cudaMemcpyHostToDevice, st));
        if (c > 0) {
          cudaEvent_t ev = p->hxe[0][(size_t)i * (split - 1) + (c - 1)];
          B200_CUDA_OK(cudaEventRecord(ev, st));
          B200_CUDA_OK(cudaStreamWaitEvent(p->hs[0], ev, 0));
        }
      }
    } else if (p->nnz > 0) {
      if (sm_copy) {
        panel_copy_kernel<<<copy_ctas, 256, 0, p->hs[0]>>>(reinterpret_cast<float4*>(p->hB + off),
                                                          reinterpret_cast<const float4*>(dB + off), p->K, (int)(w / 4),
                                                          N / 4);
        B200_LAUNCH_OK("panel_copy_kernel");
      } else {
        B200_CUDA_OK(cudaMemcpy2DAsync(p->hB + off, pitch, h_B + off, pitch, (size_t)w * sizeof(float), (size_t)p->K,
                                       cudaMemcpyHostToDevice, p->hs[0]));
      }
    }
    B200_CUDA_OK(cudaEventRecord(p->hev[0][i], p->hs[0]));
    B200_CUDA_OK(cudaStreamWaitEvent(p->hs[1], p->hev[0][i], 0));
    int rc = plan_spmm(p, 0, p->row_ptr, p->col, p->val, p->M, p->K, p->max_row, p->hB + off, N, w, p->hC + off, N,
                       p->hs[1]);
    if (rc) return rc;
    B200_CUDA_OK(cudaEventRecord(p->hev[1][i], p->hs[1]));
    B200_CUDA_OK(cudaStreamWaitEvent(p->hs[2], p->hev[1][i], 0));
    if (split > 1) {
      for (int c = 0; c < split; ++c) {
        const size_t r0 = (size_t)p->M * c / split, r1 = (size_t)p->M * (c + 1) / split;
        cudaStream_t st = c == 0 ? p->hs[2] : p->hx[1][c - 1];
        if (c > 0) B200_CUDA_OK(cudaStreamWaitEvent(st, p->hev[1][i], 0));
        if (r1 > r0)
          B200_CUDA_OK(cudaMemcpy2DAsync(h_C + r0 * N + off, pitch, p->hC + r0 * N + off, pitch, (size_t)w * sizeof(float),
                                         r1 - r0, cudaMemcpyDeviceToHost, st));
        if (c > 0) {
          cudaEvent_t ev = p->hxe[1][(size_t)i * (split - 1) + (c - 1)];
          B200_CUDA_OK(cudaEventRecord(ev, st));
          B200_CUDA_OK(cudaStreamWaitEvent(p->hs[2], ev, 0));
        }
      }
    } else if (sm_copy) {
      panel_copy_kernel<<<copy_ctas, 256, 0, p->hs[2]>>>(reinterpret_cast<float4*>(dC + off),
                                                        reinterpret_cast<const float4*>(p->hC + off), p->M, (int)(w / 4),
                                                        N / 4);
      B200_LAUNCH_OK("panel_copy_kernel");
    } else {
      B200_CUDA_OK(cudaMemcpy2DAsync(h_C + off, pitch, p->hC + off, pitch, (size_t)w * sizeof(float), (size_t)p->M,
                                     cudaMemcpyDeviceToHost, p->hs[2]));
    }
  }
  B200_CUDA_OK(cudaStreamSynchronize(p->hs[2]));
  return B200_OK;
}

int b200_plan_forward_peers(b200_plan* p, const float* d_B, int64_t ldb, int64_t N, float* d_C, int64_t ldc,
                            float* const* peer_C, int n_peers, float* mc_C, void* stream) {
  B200_REQUIRE(p != nullptr, "null plan");
  B200_REQUIRE(n_peers >= 0 && n_peers <= 7, "n_peers must be in [0, 7]");
  B200_REQUIRE(n_peers == 0 || peer_C != nullptr, "null peer pointer array");
  for (int i = 0; i < n_peers; ++i) B200_REQUIRE(peer_C[i] != nullptr, "null peer pointer");
  return plan_spmm(p, 0, p->row_ptr, p->col, p->val, p->M, p->K, p->max_row, d_B, ldb, N, d_C, ldc,
                   (cudaStream_t)stream, peer_C, n_peers, mc_C);
}

int b200_plan_backward_B(b200_plan* p, const float* d_G, int64_t ldg, int64_t N, float* d_dB,
                         int64_t lddb, void* stream) {
  B200_REQUIRE(p != nullptr, "null plan");
  int rc = plan_ensure_transpose(p, (cudaStream_t)stream);
  if (rc) return rc;
  return plan_spmm_ex(p, 1, p->t_row_ptr, p->t_col, p->t_val, p->K, p->M, p->t_max_row, d_G, ldg, N, d_dB, lddb,
                      nullptr, 0u, (cudaStream_t)stream);
}

static int plan_backward_A(b200_plan* p, const float* d_G, int64_t ldg, const float* d_B, int64_t ldb, int64_t N,
                           float* d_dval, uint32_t flags, cudaStream_t s) {
  const bool tr = (flags & B200_OP_B_TRANSPOSED) != 0;  // G given as [N, M], B as [N, K]
  B200_REQUIRE(N >= 0 && N < 0x7fffffffLL, "N out of range");
  B200_REQUIRE(ldg >= (tr ? p->M : N) && ldb >= (tr ? p->K : N) && ldg < 0x7fffffffLL && ldb < 0x7fffffffLL,
               "bad leading dimension");
  if (p->nnz == 0 || p->M == 0) return B200_OK;
  B200_REQUIRE(d_dval != nullptr && (N == 0 || (d_G && d_B)), "null G / B / output");
  int rc_dev = plan_check_device(p);
  if (rc_dev) return rc_dev;
  if (N > 0 && !p->samp_failed && plan_wants_tc(p, p->M, p->K, N, flags, /*sddmm=*/true)) {
    // D = G * B^T on the tensor cores, sampled on A's pattern
    int rc = B200_OK;
    if (!p->samp_built) {
      rc = tc_sample_build(p->row_ptr, p->col, p->M, p->K, p->nnz, &p->samp, s);
      if (rc == B200_OK) p->samp_built = true;
      else p->samp_failed = true;
    }
    if (rc == B200_OK) {
      const int kind = options().tc_kind;
      const int nt = options().tc_nt > 0 ? options().tc_nt
                                         : tc_pick_nt_sddmm((p->M + 127) / 128, p->K, (int)N, kind, device_sm_count());
      const size_t need = tc_sddmm_ws_bytes(p->M, p->K, (int)N, nt, kind);
      if (p->sd_ws_bytes < need) {
        if (p->sd_ws) { B200_CUDA_OK(cudaStreamSynchronize(s)); B200_CUDA_OK(cudaFree(p->sd_ws)); p->sd_ws = nullptr; p->sd_ws_bytes = 0; }
        if (cudaMalloc(&p->sd_ws, need) != cudaSuccess) {
          cudaGetLastError();
          set_error("cudaMalloc of %zu bytes for the packed SDDMM operands failed", need);
          return B200_ERR_ALLOC;
        }
        p->sd_ws_bytes = need;
        B200_CUDA_OK(cudaMemsetAsync(p->sd_ws, 0, 256, s));
      }
      p->last_path = 1;
      return launch_tc_sddmm(p->samp, p->row_ptr, p->col, d_G, ldg, d_B, ldb, (int)N, tr, d_dval, p->sd_ws,
                             p->sd_ws_bytes, nt, kind, s);
    }
    if (flags & B200_OP_FORCE_TC) return rc;
  }
  if (tr) {  // operands held transposed on a matrix that is too sparse / small for the tiles: plain CSR-order kernel
    p->last_path = 2;
    return launch_sddmm_strided(p->row_ptr, p->col, p->M, (int)N, d_G, 1, ldg, d_B, 1, ldb, d_dval, s);
  }
  if (p->sd_bnd == nullptr) {
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&p->sd_bnd), sddmm_bounds_bytes(p->M, p->nnz)));
    int rc = launch_sddmm_partition(p->row_ptr, p->M, p->nnz, p->sd_bnd, s);
    if (rc) return rc;
  }
  p->last_path = 0;
  return launch_sddmm(p->row_ptr, p->col, p->M, p->nnz, d_G, ldg, d_B, ldb, (int)N, d_dval, p->sd_bnd, s);
}

int b200_plan_backward_A(b200_plan* p, const float* d_G, int64_t ldg, const float* d_B, int64_t ldb,
                         int64_t N, float* d_dval, void* stream) {
  B200_REQUIRE(p != nullptr, "null plan");
  return plan_backward_A(p, d_G, ldg, d_B, ldb, N, d_dval, 0u, (cudaStream_t)stream);
}

int b200_plan_backward_A_ex(b200_plan* p, const float* d_G, int64_t ldg, const float* d_B, int64_t ldb,
                            int64_t N, float* d_dval, uint32_t flags, void* stream) {
  B200_REQUIRE(p != nullptr, "null plan");
  B200_REQUIRE((flags & (B200_OP_RELU | B200_OP_C_TRANSPOSED)) == 0, "flags that do not apply to the SDDMM");
  return plan_backward_A(p, d_G, ldg, d_B, ldb, N, d_dval, flags, (cudaStream_t)stream);
}

// ---- host-buffer entry points ------------------------------------------------------------

int b200_spmm_csr_forward_host(const int32_t* h_row_ptr, const int32_t* h_col_ind, const float* h_val,
                               int64_t M, int64_t K, int64_t nnz, const float* h_B, int64_t N,
                               float* h_C) {
  int rc = check_device();
  if (rc) return rc;
  rc = check_csr_dims(M, K, nnz);
  if (rc) return rc;
  B200_REQUIRE(N >= 0, "negative N");
  if (M == 0 || N == 0) return B200_OK;
  B200_REQUIRE(h_row_ptr && h_C && (nnz == 0 || (h_col_ind && h_val && h_B)), "null host pointer");
  cudaStream_t s = nullptr;
  DevBuf rp, ci, va, b, c;
  if ((rc = rp.alloc((M + 1) * 4)) || (rc = ci.alloc(nnz * 4)) || (rc = va.alloc(nnz * 4)) ||
      (rc = b.alloc((size_t)K * N * 4)) || (rc = c.alloc((size_t)M * N * 4)))
    return rc;
  H2D(rp.p, h_row_ptr, (size_t)(M + 1) * 4, s);
  H2D(ci.p, h_col_ind, (size_t)nnz * 4, s);
  H2D(va.p, h_val, (size_t)nnz * 4, s);
  if (nnz > 0) H2D(b.p, h_B, (size_t)K * N * 4, s);
  rc = b200_spmm_csr_forward(rp.as<int>(), ci.as<int>(), va.as<float>(), M, K, nnz, b.as<float>(), N, N,
                             c.as<float>(), N, nullptr, 0, s);
  if (rc) return rc;
  D2H(h_C, c.p, (size_t)M * N * 4, s);
  B200_CUDA_OK(cudaStreamSynchronize(s));
  return B200_OK;
}

int b200_sddmm_csr_host(const int32_t* h_row_ptr, const int32_t* h_col_ind, int64_t M, int64_t K,
                        int64_t nnz, const float* h_G, const float* h_B, int64_t N, float* h_out_val) {
  int rc = check_device();
  if (rc) return rc;
  rc = check_csr_dims(M, K, nnz);
  if (rc) return rc;
  B200_REQUIRE(N >= 0, "negative N");
  if (nnz == 0) return B200_OK;
  B200_REQUIRE(h_row_ptr && h_col_ind && h_out_val && (N == 0 || (h_G && h_B)), "null host pointer");
  cudaStream_t s = nullptr;
  DevBuf rp, ci, g, b, o;
  if ((rc = rp.alloc((M + 1) * 4)) || (rc = ci.alloc(nnz * 4)) || (rc = g.alloc((size_t)M * N * 4)) ||
      (rc = b.alloc((size_t)K * N * 4)) || (rc = o.alloc(nnz * 4)))
    return rc;
  H2D(rp.p, h_row_ptr, (size_t)(M + 1) * 4, s);
  H2D(ci.p, h_col_ind, (size_t)nnz * 4, s);
  H2D(g.p, h_G, (size_t)M * N * 4, s);
  H2D(b.p, h_B, (size_t)K * N * 4, s);
  rc = b200_sddmm_csr(rp.as<int>(), ci.as<int>(), M, K, nnz, g.as<float>(), N, b.as<float>(), N, N,
                      o.as<float>(), s);
  if (rc) return rc;
  D2H(h_out_val, o.p, (size_t)nnz * 4, s);
  B200_CUDA_OK(cudaStreamSynchronize(s));
  return B200_OK;
}

int b200_csr_transpose_host(const int32_t* h_row_ptr, const int32_t* h_col_ind, const float* h_val,
                            int64_t M, int64_t K, int64_t nnz, int32_t* h_t_row_ptr,
                            int32_t* h_t_col_ind, float* h_t_val) {
  int rc = check_device();
  if (rc) return rc;
  rc = check_csr_dims(M, K, nnz);
  if (rc) return rc;
  B200_REQUIRE(h_row_ptr && h_t_row_ptr, "null host pointer");
  B200_REQUIRE(nnz == 0 || (h_col_ind && h_val && h_t_col_ind && h_t_val), "null host pointer");
  cudaStream_t s = nullptr;
  DevBuf rp, ci, va, trp, tci, tva;
  if ((rc = rp.alloc((M + 1) * 4)) || (rc = ci.alloc(nnz * 4)) || (rc = va.alloc(nnz * 4)) ||
      (rc = trp.alloc((K + 1) * 4)) || (rc = tci.alloc(nnz * 4)) || (rc = tva.alloc(nnz * 4)))
    return rc;
  H2D(rp.p, h_row_ptr, (size_t)(M + 1) * 4, s);
  H2D(ci.p, h_col_ind, (size_t)nnz * 4, s);
  H2D(va.p, h_val, (size_t)nnz * 4, s);
  rc = b200_csr_transpose(rp.as<int>(), ci.as<int>(), va.as<float>(), M, K, nnz, trp.as<int>(),
                          tci.as<int>(), tva.as<float>(), nullptr, s);
  if (rc) return rc;
  D2H(h_t_row_ptr, trp.p, (size_t)(K + 1) * 4, s);
  D2H(h_t_col_ind, tci.p, (size_t)nnz * 4, s);
  D2H(h_t_val, tva.p, (size_t)nnz * 4, s);
  B200_CUDA_OK(cudaStreamSynchronize(s));
  return B200_OK;
}

int b200_spmm_csr_backward_B_host(const int32_t* h_row_ptr, const int32_t* h_col_ind, const float* h_val,
                                  int64_t M, int64_t K, int64_t nnz, const float* h_G, int64_t N,
                                  float* h_dB) {
  int rc = check_device();
  if (rc) return rc;
  rc = check_csr_dims(M, K, nnz);
  if (rc) return rc;
  B200_REQUIRE(N >= 0, "negative N");
  if (K == 0 || N == 0) return B200_OK;
  B200_REQUIRE(h_row_ptr && h_dB && (nnz == 0 || (h_col_ind && h_val && h_G)), "null host pointer");
  cudaStream_t s = nullptr;
  DevBuf rp, ci, va, g, db;
  if ((rc = rp.alloc((M + 1) * 4)) || (rc = ci.alloc(nnz * 4)) || (rc = va.alloc(nnz * 4)) ||
      (rc = g.alloc((size_t)M * N * 4)) || (rc = db.alloc((size_t)K * N * 4)))
    return rc;
  H2D(rp.p, h_row_ptr, (size_t)(M + 1) * 4, s);
  H2D(ci.p, h_col_ind, (size_t)nnz * 4, s);
  H2D(va.p, h_val, (size_t)nnz * 4, s);
  if (nnz > 0) H2D(g.p, h_G, (size_t)M * N * 4, s);
  b200_plan* plan = nullptr;
  rc = b200_plan_create(rp.as<int>(), ci.as<int>(), va.as<float>(), M, K, nnz, &plan, s);
  if (rc) return rc;
  rc = b200_plan_backward_B(plan, g.as<float>(), N, N, db.as<float>(), N, s);
  if (rc == B200_OK) {
    cudaError_t e = cudaMemcpyAsync(h_dB, db.p, (size_t)K * N * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = B200_ERR_CUDA; }
  }
  b200_plan_destroy(plan);
  return rc;
}

int b200_dense_to_csr_host(const float* h_dense, int64_t M, int64_t K, int32_t* h_row_ptr,
                           int32_t* h_col_ind, float* h_val, int64_t capacity, int64_t* nnz) {
  int rc = check_device();
  if (rc) return rc;
  B200_REQUIRE(M >= 0 && K >= 0 && M < 0x7fffffffLL && K < 0x7fffffffLL, "dimension out of range");
  B200_REQUIRE(nnz != nullptr && h_row_ptr != nullptr, "null nnz / row_ptr");
  B200_REQUIRE(M == 0 || K == 0 || h_dense != nullptr, "null dense");
  cudaStream_t s = nullptr;
  DevBuf d, rp, ci, va;
  if ((rc = d.alloc((size_t)M * K * 4)) || (rc = rp.alloc((M + 1) * 4))) return rc;
  H2D(d.p, h_dense, (size_t)M * K * 4, s);
  rc = b200_dense_to_csr_count(d.as<float>(), M, K, K, rp.as<int>(), nnz, s);
  if (rc) return rc;
  D2H(h_row_ptr, rp.p, (size_t)(M + 1) * 4, s);
  B200_CUDA_OK(cudaStreamSynchronize(s));
  if (*nnz > capacity) {
    set_error("capacity %lld too small for %lld non-zeros", (long long)capacity, (long long)*nnz);
    return B200_ERR_INVALID;
  }
  if (*nnz == 0) return B200_OK;
  B200_REQUIRE(h_col_ind && h_val, "null col/val");
  if ((rc = ci.alloc(*nnz * 4)) || (rc = va.alloc(*nnz * 4))) return rc;
  rc = b200_dense_to_csr_fill(d.as<float>(), M, K, K, rp.as<int>(), ci.as<int>(), va.as<float>(), s);
  if (rc) return rc;
  D2H(h_col_ind, ci.p, (size_t)*nnz * 4, s);
  D2H(h_val, va.p, (size_t)*nnz * 4, s);
  B200_CUDA_OK(cudaStreamSynchronize(s));
  return B200_OK;
}

// ---- host-side partition helpers ---------------------------------------------------------
int b200_merge_path_boundary(const int32_t* h_row_ptr, int64_t M, int64_t nnz, int64_t d,
                             int64_t heavy_threshold, int64_t* row, int64_t* nz) {
  B200_REQUIRE(h_row_ptr && row && nz, "null pointer");
  B200_REQUIRE(M >= 0 && nnz >= 0 && M < 0x7fffffffLL && nnz < 0x7fffffffLL, "dimension out of range");
  B200_REQUIRE(heavy_threshold >= 0 && heavy_threshold < 0x7fffffffLL, "heavy threshold out of range");
  Boundary b = merge_path_boundary(h_row_ptr, (int)M, (int)nnz, d, (int)heavy_threshold);
  *row = b.row;
  *nz = b.nz;
  return B200_OK;
}

int b200_shard_rows(const int32_t* h_row_ptr, int64_t M, int parts, double nnz_weight,
                    double row_weight, int64_t* bounds) {
  B200_REQUIRE(h_row_ptr && bounds, "null pointer");
  B200_REQUIRE(M >= 0 && parts >= 1, "bad M / parts");
  B200_REQUIRE(nnz_weight >= 0 && row_weight >= 0 && (nnz_weight + row_weight) > 0, "bad weights");
  // cost(r) = nnz_weight * row_ptr[r] + row_weight * r is non-decreasing; shard g starts at the
  // first row whose prefix cost reaches g/parts of the total.
  const double total = nnz_weight * (double)h_row_ptr[M] + row_weight * (double)M;
  bounds[0] = 0;
  for (int g = 1; g < parts; ++g) {
    const double target = total * (double)g / (double)parts;
    int64_t lo = 0, hi = M;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      double c = nnz_weight * (double)h_row_ptr[mid] + row_weight * (double)mid;
      if (c >= target) hi = mid; else lo = mid + 1;
    }
    bounds[g] = std::max(lo, bounds[g - 1]);
  }
  bounds[parts] = M;
  return B200_OK;
}

}  // extern "C"
#pragma GCC visibility pop
