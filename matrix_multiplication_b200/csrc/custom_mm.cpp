// custom_mm.cpp — pybind11 module `custom_mm`: the reference's Python-visible boundary
// (reference src/custom_mm.cpp:393-416, same 16 names and argument meaning) on top of
// the C ABI in include/spmm_b200.h.  Host code only; every device kernel lives in
// libspmm_b200.so.  torch is used for what the reference uses it for: tensors as
// argument carriers, the caching allocator, the current stream and the cuBLAS handle.
//
// Differences from the reference that are deliberate:
//   * strides are checked instead of ignored (src/custom_mm.cpp:173,210);
//   * kernels run on torch's current stream, not the legacy default stream;
//   * every failure raises (cuSPARSE failures were printed and swallowed,
//     src/Utilities.cuh:145-154);
//   * the layer registries own references to their tensors (the reference kept raw
//     pointers and cudaFree'd memory it did not own, src/custom_mm.cpp:236-281).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <cublas_v2.h>
#include <torch/extension.h>

#include <climits>
#include <stdexcept>
#include <string>
#include <unordered_map>

#include "spmm_b200.h"

namespace {

void check(int rc) {
  if (rc == B200_OK) return;
  std::string msg = std::string("spmm_b200: ") + b200_last_error();
  if (rc == B200_ERR_INVALID) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

void* cur_stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }

void check_csr(const torch::Tensor& values, const torch::Tensor& columns, const torch::Tensor& offsets,
               int64_t nnz, int64_t rows) {
  TORCH_CHECK(values.is_cuda() && columns.is_cuda() && offsets.is_cuda(), "CSR arrays must be CUDA tensors");
  TORCH_CHECK(values.scalar_type() == torch::kFloat32, "A_values must be float32");
  TORCH_CHECK(columns.scalar_type() == torch::kInt32, "A_columns must be int32");
  TORCH_CHECK(offsets.scalar_type() == torch::kInt32, "A_offsets must be int32");
  TORCH_CHECK(values.is_contiguous() && columns.is_contiguous() && offsets.is_contiguous(),
              "CSR arrays must be contiguous");
  TORCH_CHECK(values.numel() >= nnz && columns.numel() >= nnz, "CSR arrays shorter than nnz");
  TORCH_CHECK(offsets.numel() >= rows + 1, "A_offsets shorter than rows + 1");
}

void check_dense(const torch::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
  TORCH_CHECK(t.dim() == 2, name, " must be 2-D");
  TORCH_CHECK(t.size(1) <= 1 || t.stride(1) == 1, name,
              " must have unit stride in its last dimension (strides are honoured, not ignored)");
  TORCH_CHECK(t.size(0) <= 1 || t.stride(0) >= t.size(1), name, " has overlapping rows");
}

int64_t ld_of(const torch::Tensor& t) { return t.size(0) > 1 ? t.stride(0) : std::max<int64_t>(t.size(1), 1); }

// C[M,N] = A_csr[M,K] * B[K,N]; writes C in place and returns it (reference convention).
torch::Tensor spmm_into(const torch::Tensor& values, const torch::Tensor& columns,
                        const torch::Tensor& offsets, int64_t nnz, int64_t rows, int64_t cols,
                        const torch::Tensor& B, torch::Tensor C, uint32_t flags) {
  check_csr(values, columns, offsets, nnz, rows);
  check_dense(B, "B");
  check_dense(C, "C");
  TORCH_CHECK(B.size(0) == cols, "B has ", B.size(0), " rows, A has ", cols, " columns");
  TORCH_CHECK(C.size(0) == rows && C.size(1) == B.size(1), "C must be [A_rows, B_cols]");
  c10::cuda::CUDAGuard guard(B.device());
  const int64_t N = B.size(1);
  const size_t ws_bytes = b200_spmm_workspace_bytes(rows, nnz, N);
  auto ws = torch::empty({(int64_t)ws_bytes}, torch::dtype(torch::kUInt8).device(B.device()));
  check(b200_spmm_csr_forward_flags(offsets.data_ptr<int>(), columns.data_ptr<int>(),
                                    values.data_ptr<float>(), rows, cols, nnz, B.data_ptr<float>(),
                                    ld_of(B), N, C.data_ptr<float>(), ld_of(C), ws.data_ptr(), ws_bytes,
                                    cur_stream(), flags));
  return C;
}

// ---- the reference's entry points ---------------------------------------------------------
bool g_cublas_init = false, g_engine_init = false;

void init_cublas_handle() {
  at::globalContext().lazyInitDevice(c10::DeviceType::CUDA);  // the first CUDA use of the process may be this call
  (void)at::cuda::getCurrentCUDABlasHandle();
  g_cublas_init = true;
}
void destroy_cublas_handle() { g_cublas_init = false; }
void init_engine() {
  at::globalContext().lazyInitDevice(c10::DeviceType::CUDA);
  int sms = 0;
  check(b200_smoke_kernel(cur_stream(), &sms));
  g_engine_init = true;
}
void cusparse_clean();
void tiledspmm_clean();
void destroy_engine() {
  cusparse_clean();
  tiledspmm_clean();
  g_engine_init = false;
}

torch::Tensor cusparse_mmul(torch::Tensor A_values, torch::Tensor A_columns, torch::Tensor A_offsets,
                            int nnzA, int A_rows, int A_cols, torch::Tensor B, torch::Tensor C) {
  return spmm_into(A_values, A_columns, A_offsets, nnzA, A_rows, A_cols, B, C, 0);
}

// Same math as cusparse_mmul; kept as a separate entry point like the reference
// (src/custom_mm.cpp:166-179) and routed to the variant that reads col/val straight from
// global memory, so both staging variants stay exercised.  A 3-D B ([batch,K,N], the
// reference's `batch_size = B.numel() / (n*k)`, src/naive_sparse_mm.cu:114) reuses A per batch.
torch::Tensor naive_spmm(torch::Tensor A_values, torch::Tensor A_columns, torch::Tensor A_offsets,
                         int nnzA, int A_rows, int A_cols, torch::Tensor B, torch::Tensor C) {
  if (B.dim() == 3) {
    TORCH_CHECK(C.dim() == 3 && C.size(0) == B.size(0), "batched naive_spmm needs C [batch, M, N]");
    for (int64_t i = 0; i < B.size(0); ++i)
      spmm_into(A_values, A_columns, A_offsets, nnzA, A_rows, A_cols, B[i], C[i], B200_FLAG_NO_STAGE);
    return C;
  }
  return spmm_into(A_values, A_columns, A_offsets, nnzA, A_rows, A_cols, B, C, B200_FLAG_NO_STAGE);
}

// Dense comparison path: row-major C = op(A) * op(B) through cuBLAS (reference
// src/baseline_mm.cu:52-155).  Row-major C is column-major C^T = op(B)^T * op(A)^T.
void gemm_rowmajor(const float* A, const float* B, float* C, int64_t a_rows, int64_t a_cols,
                   int64_t b_rows, int64_t b_cols, bool transa, bool transb, int64_t batch,
                   int64_t strideA, int64_t strideB, int64_t strideC) {
  const int64_t c_rows = transa ? a_cols : a_rows;
  const int64_t c_cols = transb ? b_rows : b_cols;
  const int64_t k = transa ? a_rows : a_cols;
  const int64_t kb = transb ? b_cols : b_rows;
  TORCH_CHECK(k == kb, "inner dimensions differ: ", k, " vs ", kb);
  if (c_rows == 0 || c_cols == 0) return;
  TORCH_CHECK(c_rows <= INT_MAX && c_cols <= INT_MAX && k <= INT_MAX && a_cols <= INT_MAX && b_cols <= INT_MAX &&
              batch <= INT_MAX, "cuBLAS takes 32-bit dimensions");
  auto handle = at::cuda::getCurrentCUDABlasHandle();
  const float alpha = 1.f, beta = 0.f;
  cublasStatus_t st;
  if (batch <= 1) {
    st = cublasSgemm(handle, transb ? CUBLAS_OP_T : CUBLAS_OP_N, transa ? CUBLAS_OP_T : CUBLAS_OP_N,
                     (int)c_cols, (int)c_rows, (int)k, &alpha, B, (int)b_cols, A, (int)a_cols, &beta, C,
                     (int)c_cols);
  } else {
    st = cublasSgemmStridedBatched(handle, transb ? CUBLAS_OP_T : CUBLAS_OP_N,
                                   transa ? CUBLAS_OP_T : CUBLAS_OP_N, (int)c_cols, (int)c_rows, (int)k,
                                   &alpha, B, (int)b_cols, strideB, A, (int)a_cols, strideA, &beta, C,
                                   (int)c_cols, strideC, (int)batch);
  }
  if (st != CUBLAS_STATUS_SUCCESS)
    throw std::logic_error("cuBLAS API failed with status " + std::to_string((int)st));
}

void check_gemm_tensor(const torch::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kFloat32, name, " must be a float32 CUDA tensor");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous (pass transa/transb instead of a transposed view)");
}

torch::Tensor cublas_mmul(torch::Tensor A, torch::Tensor B, torch::Tensor C, bool transa, bool transb) {
  check_gemm_tensor(A, "A"); check_gemm_tensor(B, "B"); check_gemm_tensor(C, "C");
  TORCH_CHECK(A.dim() == 2 && B.dim() == 2, "cublas_mmul takes 2-D tensors");
  c10::cuda::CUDAGuard guard(A.device());
  gemm_rowmajor(A.data_ptr<float>(), B.data_ptr<float>(), C.data_ptr<float>(), A.size(0), A.size(1),
                B.size(0), B.size(1), transa, transb, 1, 0, 0, 0);
  return C;
}

torch::Tensor cublas_bmm(torch::Tensor A, torch::Tensor B, torch::Tensor C, int dim, bool transa,
                         bool transb) {
  if (dim == 2) return cublas_mmul(A, B, C, transa, transb);
  if (dim != 3 && dim != 4) throw std::invalid_argument("Invalid dim argument.");
  check_gemm_tensor(A, "A"); check_gemm_tensor(B, "B"); check_gemm_tensor(C, "C");
  TORCH_CHECK(A.dim() == dim && B.dim() == dim, "tensor rank does not match dim");
  int64_t batch = 1;
  for (int d = 0; d < dim - 2; ++d) {
    TORCH_CHECK(A.size(d) == B.size(d), "batch dimensions differ");
    batch *= A.size(d);
  }
  const int64_t ar = A.size(dim - 2), ac = A.size(dim - 1), br = B.size(dim - 2), bc = B.size(dim - 1);
  const int64_t c_rows = transa ? ac : ar, c_cols = transb ? br : bc;
  c10::cuda::CUDAGuard guard(A.device());
  gemm_rowmajor(A.data_ptr<float>(), B.data_ptr<float>(), C.data_ptr<float>(), ar, ac, br, bc, transa,
                transb, batch, ar * ac, br * bc, c_rows * c_cols);
  return C;
}

void dummy_kernel_launch() {
  int sms = 0;
  check(b200_smoke_kernel(cur_stream(), &sms));
}

// ---- layer-keyed inspect / execute / clean registries --------------------------------------
struct Layer {
  torch::Tensor offsets, columns, values;  // device CSR kept alive by the registry
  int64_t M = 0, K = 0, N = 0, nnz = 0;    // A is M x K, dense operand has N columns
  b200_plan* plan = nullptr;
};
std::unordered_map<std::string, Layer> g_cusparse_layers, g_tiled_layers;

void drop(std::unordered_map<std::string, Layer>& reg, const std::string& name) {
  auto it = reg.find(name);
  if (it != reg.end()) {
    b200_plan_destroy(it->second.plan);
    reg.erase(it);
  }
}

void add_layer(std::unordered_map<std::string, Layer>& reg, const std::string& name, Layer L) {
  drop(reg, name);
  c10::cuda::CUDAGuard guard(L.offsets.device());
  check(b200_plan_create(L.offsets.data_ptr<int>(), L.columns.data_ptr<int>(), L.values.data_ptr<float>(),
                         L.M, L.K, L.nnz, &L.plan, cur_stream()));
  reg.emplace(name, std::move(L));
}

Layer& find_layer(std::unordered_map<std::string, Layer>& reg, const std::string& name) {
  auto it = reg.find(name);
  if (it == reg.end()) throw std::runtime_error("Invalid handle_id! (unknown layer '" + name + "')");
  return it->second;
}

// B and C are column-major by default, as the reference compiles them
// (src/baseline_mm.cu:297,300 CUSPARSE_ORDER_COL; src/sparse_mm.cu:16 ORDER COLMAJOR):
// B's storage is [N][K] and C's is [N][M].  row_major=true takes [K,N] / [M,N] directly.
void layer_mm(Layer& L, const torch::Tensor& B, torch::Tensor& C, bool row_major) {
  TORCH_CHECK(B.is_cuda() && C.is_cuda() && B.scalar_type() == torch::kFloat32 &&
              C.scalar_type() == torch::kFloat32, "B and C must be float32 CUDA tensors");
  TORCH_CHECK(B.is_contiguous() && C.is_contiguous(), "B and C must be contiguous");
  TORCH_CHECK(B.numel() == L.K * L.N, "B must hold K*N = ", L.K * L.N, " elements");
  TORCH_CHECK(C.numel() == L.M * L.N, "C must hold M*N = ", L.M * L.N, " elements");
  c10::cuda::CUDAGuard guard(B.device());
  if (row_major) {
    check(b200_plan_forward(L.plan, B.data_ptr<float>(), L.N, L.N, C.data_ptr<float>(), L.N, cur_stream()));
    return;
  }
  // column-major B [K,N] is B^T [N,K] row-major, column-major C is C^T [N,M] row-major: the engine reads / writes
  // them in place (no transposed copies)
  check(b200_plan_forward_ex(L.plan, B.data_ptr<float>(), L.K, L.N, C.data_ptr<float>(), L.M, nullptr,
                             B200_OP_B_TRANSPOSED | B200_OP_C_TRANSPOSED, cur_stream()));
}

void cusparse_inspect(torch::Tensor displ, torch::Tensor colindex, torch::Tensor value, int nnz, int M,
                      int N, int K, std::string layer) {
  check_csr(value, colindex, displ, nnz, M);
  Layer L;
  L.offsets = displ; L.columns = colindex; L.values = value;
  L.M = M; L.K = K; L.N = N; L.nnz = nnz;
  add_layer(g_cusparse_layers, layer, std::move(L));
}
torch::Tensor cusparse_mmul_opt(torch::Tensor B, torch::Tensor C, std::string layer, bool row_major) {
  layer_mm(find_layer(g_cusparse_layers, layer), B, C, row_major);
  return C;
}
void cusparse_clean() {
  for (auto& kv : g_cusparse_layers) b200_plan_destroy(kv.second.plan);
  g_cusparse_layers.clear();
}

// TiledSpMM convention (src/custom_mm.cpp:283-290): C = M x K, A = M x N, B = N x K;
// the CSR / COO arrays arrive on the CPU (int64 displ/colindex for CSR, int32 for COO).
void tiled_register(int M, int N, int K, torch::Tensor crow64, torch::Tensor col64, torch::Tensor val,
                    const std::string& layer) {
  TORCH_CHECK(at::cuda::is_available(), "tiledspmm needs a CUDA device: no CPU fallback");
  Layer L;
  auto dev = torch::Device(torch::kCUDA, at::cuda::current_device());
  L.offsets = crow64.to(torch::kInt32).to(dev).contiguous();
  L.columns = col64.to(torch::kInt32).to(dev).contiguous();
  L.values = val.to(torch::kFloat32).to(dev).contiguous();
  L.M = M; L.K = N; L.N = K; L.nnz = L.values.numel();
  TORCH_CHECK(L.offsets.numel() == M + 1, "displ must have M + 1 entries");
  add_layer(g_tiled_layers, layer, std::move(L));
}
void tiledspmm_inspect_csr(int M, int N, int K, torch::Tensor displ, torch::Tensor colindex,
                           torch::Tensor value, std::string layer) {
  tiled_register(M, N, K, displ, colindex, value, layer);
}
void tiledspmm_inspect_coo(int M, int N, int K, long nnz, torch::Tensor rowidx, torch::Tensor colidx,
                           torch::Tensor value, std::string layer) {
  TORCH_CHECK(rowidx.numel() >= nnz && colidx.numel() >= nnz && value.numel() >= nnz, "COO arrays shorter than nnz");
  auto r = rowidx.flatten().slice(0, 0, nnz).to(torch::kCPU).to(torch::kInt64);
  auto c = colidx.flatten().slice(0, 0, nnz).to(torch::kCPU).to(torch::kInt64);
  auto v = value.flatten().slice(0, 0, nnz).to(torch::kCPU);
  // the reference's coo2csr (src/sparse_mm.cu:110-134) assumes row-sorted input; a stable
  // sort by (row, col) accepts any order and leaves sorted input untouched.
  auto order = (r * (int64_t)N + c).argsort(/*stable=*/true);
  r = r.index_select(0, order); c = c.index_select(0, order); v = v.index_select(0, order);
  auto crow = torch::zeros({M + 1}, torch::kInt64);
  if (nnz > 0) crow.slice(0, 1, M + 1).copy_(torch::bincount(r, {}, M).cumsum(0));
  tiled_register(M, N, K, crow, c, v, layer);
}
void tiledspmm_mm(torch::Tensor B, torch::Tensor C, std::string layer, bool row_major) {
  layer_mm(find_layer(g_tiled_layers, layer), B, C, row_major);
}
void tiledspmm_clean() {
  for (auto& kv : g_tiled_layers) b200_plan_destroy(kv.second.plan);
  g_tiled_layers.clear();
}

// ---- engine extras used by matmuls.py (not in the reference's module) ------------------------
// dense -> CSR.  Large matrices go through the one-pass kernel (count warps / offset warp / fill warps, the dense
// matrix leaves HBM once; profiles/r2_dense2csr.md: 16384^2 at 90 / 99 / 99.9 % sparsity 0.35 / 0.28 / 0.22 ms against
// 0.49 / 0.43 / 0.42 ms for count + fill).  Its outputs are sized for one kept element in eight (the kept fraction is not
// known before the pass); a denser matrix is reported by the kernel with row_ptr complete, and the exact-size fill
// finishes the job.  The kept fraction of the previous conversion is remembered: after a dense one (> 1/8) the
// two-phase pair is used directly (it is the faster one there).  Small matrices stay on the two-phase pair
// (launch-bound either way).  set_onepass_min_elems(n): threshold in elements (default 2^20; < 0 = never).
int64_t g_onepass_min_elems = 1 << 20;
double g_last_kept_fraction = 0.0;
std::vector<torch::Tensor> dense_to_csr(torch::Tensor dense, bool compact) {
  check_dense(dense, "dense");
  c10::cuda::CUDAGuard guard(dense.device());
  const int64_t M = dense.size(0), K = dense.size(1);
  auto iopt = torch::dtype(torch::kInt32).device(dense.device());
  auto crow = torch::empty({M + 1}, iopt);
  int64_t nnz = 0;
  const int64_t ld = ld_of(dense);
  const bool onepass = g_onepass_min_elems >= 0 && M * K >= std::max<int64_t>(g_onepass_min_elems, 1) && K >= 64 &&
                       ld % 4 == 0 && reinterpret_cast<uintptr_t>(dense.data_ptr<float>()) % 16 == 0 &&
                       g_last_kept_fraction <= 0.125;
  if (onepass) {
    const int64_t cap = std::max<int64_t>(M * K / 8, 4096);
    auto col = torch::empty({cap}, iopt);
    auto val = torch::empty({cap}, dense.options());
    const size_t sb = b200_dense_to_csr_onepass_scratch_bytes(M, K);
    auto scratch = torch::empty({(int64_t)sb}, torch::dtype(torch::kUInt8).device(dense.device()));
    const int rc = b200_dense_to_csr_onepass(dense.data_ptr<float>(), M, K, ld, crow.data_ptr<int>(), col.data_ptr<int>(),
                                             val.data_ptr<float>(), cap, scratch.data_ptr(), sb, &nnz, cur_stream());
    if (rc == B200_OK) {
      g_last_kept_fraction = (double)nnz / (double)(M * K);
      if (compact) return {crow, col.narrow(0, 0, nnz).clone(), val.narrow(0, 0, nnz).clone()};
      return {crow, col.narrow(0, 0, nnz), val.narrow(0, 0, nnz)};
    }
    if (nnz <= cap) check(rc);  // a real failure; otherwise: denser than the guess, row_ptr is complete
  } else {
    check(b200_dense_to_csr_count(dense.data_ptr<float>(), M, K, ld, crow.data_ptr<int>(), &nnz, cur_stream()));
  }
  if (M * K > 0) g_last_kept_fraction = (double)nnz / (double)(M * K);
  auto col = torch::empty({nnz}, iopt);
  auto val = torch::empty({nnz}, dense.options());
  check(b200_dense_to_csr_fill(dense.data_ptr<float>(), M, K, ld, crow.data_ptr<int>(), col.data_ptr<int>(),
                               val.data_ptr<float>(), cur_stream()));
  return {crow, col, val};
}

// A stack of L dense [M,K] matrices ([L*M, K] view) -> block-diagonal CSR [L*M, L*K] on the device.
std::vector<torch::Tensor> dense_to_csr_block_diag(torch::Tensor dense2d, int64_t M) {
  auto out = dense_to_csr(dense2d, false);
  const int64_t rows = dense2d.size(0), K = dense2d.size(1);
  c10::cuda::CUDAGuard guard(dense2d.device());
  check(b200_csr_block_diag_cols(out[0].data_ptr<int>(), out[1].data_ptr<int>(), rows, M, K, cur_stream()));
  return out;
}

std::vector<torch::Tensor> csr_transpose(torch::Tensor offsets, torch::Tensor columns, torch::Tensor values,
                                         int64_t M, int64_t K) {
  const int64_t nnz = values.numel();
  check_csr(values, columns, offsets, nnz, M);
  c10::cuda::CUDAGuard guard(values.device());
  auto iopt = torch::dtype(torch::kInt32).device(values.device());
  auto t_crow = torch::empty({K + 1}, iopt);
  auto t_col = torch::empty({nnz}, iopt);
  auto t_val = torch::empty({nnz}, values.options());
  auto perm = torch::empty({nnz}, iopt);
  check(b200_csr_transpose(offsets.data_ptr<int>(), columns.data_ptr<int>(), values.data_ptr<float>(), M, K,
                           nnz, t_crow.data_ptr<int>(), t_col.data_ptr<int>(), t_val.data_ptr<float>(),
                           perm.data_ptr<int>(), cur_stream()));
  return {t_crow, t_col, t_val, perm};
}

torch::Tensor gather(torch::Tensor src, torch::Tensor perm) {
  TORCH_CHECK(src.is_cuda() && perm.is_cuda() && src.scalar_type() == torch::kFloat32 &&
              perm.scalar_type() == torch::kInt32 && src.is_contiguous() && perm.is_contiguous(),
              "gather takes contiguous CUDA float32 / int32 tensors");
  c10::cuda::CUDAGuard guard(src.device());
  auto out = torch::empty({perm.numel()}, src.options());
  check(b200_gather_f32(src.data_ptr<float>(), perm.data_ptr<int>(), perm.numel(), out.data_ptr<float>(),
                        cur_stream()));
  return out;
}

torch::Tensor sddmm(torch::Tensor offsets, torch::Tensor columns, int64_t M, int64_t K, torch::Tensor G,
                    torch::Tensor B) {
  const int64_t nnz = columns.numel();
  TORCH_CHECK(offsets.is_cuda() && columns.is_cuda() && offsets.scalar_type() == torch::kInt32 &&
              columns.scalar_type() == torch::kInt32 && offsets.is_contiguous() && columns.is_contiguous(),
              "CSR index arrays must be contiguous CUDA int32 tensors");
  TORCH_CHECK(offsets.numel() >= M + 1, "offsets shorter than M + 1");
  check_dense(G, "G");
  check_dense(B, "B");
  TORCH_CHECK(G.size(0) == M && B.size(0) == K && G.size(1) == B.size(1), "G must be [M,N] and B [K,N]");
  c10::cuda::CUDAGuard guard(G.device());
  auto out = torch::empty({nnz}, G.options());
  check(b200_sddmm_csr(offsets.data_ptr<int>(), columns.data_ptr<int>(), M, K, nnz, G.data_ptr<float>(),
                       ld_of(G), B.data_ptr<float>(), ld_of(B), G.size(1), out.data_ptr<float>(), cur_stream()));
  return out;
}

// ---- plan object for matmuls.py / layers.py --------------------------------------------------------------
// One inspected CSR operand: keeps the arrays alive, owns a b200_plan (cached merge-path boundaries, cached
// transposed CSR, densified tiles of the block-dense tensor-core path).  This is the inspect / execute / clean triple
// of the reference's layer registries (src/custom_mm.cpp:236-281) as an object whose lifetime follows the tensor.
struct PlanHandle {
  torch::Tensor offsets, columns, values;
  int64_t M = 0, K = 0, nnz = 0;
  b200_plan* plan = nullptr;

  PlanHandle(torch::Tensor crow, torch::Tensor col, torch::Tensor val, int64_t M_, int64_t K_)
      : offsets(std::move(crow)), columns(std::move(col)), values(std::move(val)), M(M_), K(K_) {
    nnz = values.numel();
    check_csr(values, columns, offsets, nnz, M);
    c10::cuda::CUDAGuard guard(values.device());
    check(b200_plan_create(offsets.data_ptr<int>(), columns.data_ptr<int>(), values.data_ptr<float>(), M, K, nnz,
                           &plan, cur_stream()));
  }
  ~PlanHandle() { b200_plan_destroy(plan); }
  PlanHandle(const PlanHandle&) = delete;
  PlanHandle& operator=(const PlanHandle&) = delete;

  // C = A * B (+bias, ReLU); flags: B200_OP_* (transposed operands are described by the flags, the tensors are
  // plain row-major 2-D)
  torch::Tensor forward(const torch::Tensor& B, torch::Tensor C, const c10::optional<torch::Tensor>& bias,
                        int64_t flags) {
    check_dense(B, "B");
    check_dense(C, "C");
    const bool bt = flags & B200_OP_B_TRANSPOSED, ct = flags & B200_OP_C_TRANSPOSED;
    const int64_t N = bt ? B.size(0) : B.size(1);
    TORCH_CHECK((bt ? B.size(1) : B.size(0)) == K, "dense operand does not match A's ", K, " columns");
    TORCH_CHECK((ct ? C.size(1) : C.size(0)) == M && (ct ? C.size(0) : C.size(1)) == N, "C has the wrong shape");
    const float* bp = nullptr;
    if (bias.has_value() && bias->defined()) {
      TORCH_CHECK(bias->is_cuda() && bias->scalar_type() == torch::kFloat32 && bias->is_contiguous() &&
                  bias->numel() == M, "bias must be a contiguous float32 CUDA vector of A_rows elements");
      bp = bias->data_ptr<float>();
    }
    c10::cuda::CUDAGuard guard(B.device());
    check(b200_plan_forward_ex(plan, B.data_ptr<float>(), ld_of(B), N, C.data_ptr<float>(), ld_of(C), bp,
                               (uint32_t)flags, cur_stream()));
    return C;
  }
  // dB = A^T * G
  torch::Tensor backward_B(const torch::Tensor& G, torch::Tensor dB, int64_t flags) {
    check_dense(G, "G");
    check_dense(dB, "dB");
    const bool bt = flags & B200_OP_B_TRANSPOSED, ct = flags & B200_OP_C_TRANSPOSED;
    const int64_t N = bt ? G.size(0) : G.size(1);
    TORCH_CHECK((bt ? G.size(1) : G.size(0)) == M, "G does not match A's ", M, " rows");
    TORCH_CHECK((ct ? dB.size(1) : dB.size(0)) == K && (ct ? dB.size(0) : dB.size(1)) == N, "dB has the wrong shape");
    c10::cuda::CUDAGuard guard(G.device());
    check(b200_plan_backward_B_ex(plan, G.data_ptr<float>(), ld_of(G), N, dB.data_ptr<float>(), ld_of(dB),
                                  (uint32_t)flags, cur_stream()));
    return dB;
  }
  // grad of the CSR values on A's pattern; with B200_OP_B_TRANSPOSED both operands are [N, .] (tensor-core path)
  torch::Tensor backward_A(const torch::Tensor& G, const torch::Tensor& B, int64_t flags) {
    check_dense(G, "G");
    check_dense(B, "B");
    const bool tr = flags & B200_OP_B_TRANSPOSED;
    const int64_t N = tr ? G.size(0) : G.size(1);
    TORCH_CHECK((tr ? G.size(1) : G.size(0)) == M && (tr ? B.size(1) : B.size(0)) == K &&
                (tr ? B.size(0) : B.size(1)) == N, "G must be [M,N] and B [K,N] (or [N,M] and [N,K] when transposed)");
    c10::cuda::CUDAGuard guard(G.device());
    auto out = torch::empty({nnz}, G.options());
    check(b200_plan_backward_A_ex(plan, G.data_ptr<float>(), ld_of(G), B.data_ptr<float>(), ld_of(B), N,
                                  out.data_ptr<float>(), (uint32_t)flags, cur_stream()));
    return out;
  }
  // ---- the FC layer in two calls (layers.cusparseLinear): everything between the Python autograd hooks happens here,
  // so a launch-bound layer (config 2) pays two extension calls per step instead of ~20 Python-dispatched ops ----
  // y[tokens, M] = act(x[tokens, K] * A^T + bias)
  torch::Tensor fc_forward(const torch::Tensor& x, const c10::optional<torch::Tensor>& bias, bool relu, bool small) {
    check_dense(x, "x");
    const int64_t tokens = x.size(0);
    const int64_t flags = relu ? B200_OP_RELU : 0;
    if (small || would_use_tc(tokens, 0, flags)) {
      auto y = torch::empty({tokens, M}, x.options());
      forward(x, y, bias, flags | B200_OP_B_TRANSPOSED | B200_OP_C_TRANSPOSED);
      return y;
    }
    auto yT = torch::empty({M, tokens}, x.options());
    forward(x.t().contiguous(), yT, bias, flags);
    return yT.t().contiguous();
  }
  // (d_values, d_x, d_bias) for g = dL/dy [tokens, M]; y is the saved output when the ReLU was fused
  std::vector<c10::optional<torch::Tensor>> fc_backward(torch::Tensor g, const torch::Tensor& x,
                                                        const c10::optional<torch::Tensor>& y, bool relu, bool small,
                                                        bool need_values, bool need_x, bool need_bias) {
    if (relu) g = g * (y.value() > 0);
    g = g.contiguous();
    check_dense(g, "grad_output");
    const int64_t tokens = g.size(0);
    std::vector<c10::optional<torch::Tensor>> out(3);
    if (need_values) {
      if (small || would_use_tc(tokens, 2, 0)) out[0] = backward_A(g, x, B200_OP_B_TRANSPOSED);
      else out[0] = backward_A(g.t().contiguous(), x.t().contiguous(), 0);
    }
    if (need_x) {
      if (small || would_use_tc(tokens, 1, 0)) {
        auto dx = torch::empty({tokens, K}, g.options());
        backward_B(g, dx, B200_OP_B_TRANSPOSED | B200_OP_C_TRANSPOSED);
        out[1] = dx;
      } else {
        auto dxT = torch::empty({K, tokens}, g.options());
        backward_B(g.t().contiguous(), dxT, 0);
        out[1] = dxT.t().contiguous();
      }
    }
    if (need_bias) out[2] = g.sum(0);
    return out;
  }
  void values_updated() { check(b200_plan_values_updated(plan)); }
  bool would_use_tc(int64_t N, int64_t side, int64_t flags) {
    return b200_plan_would_use_tc(plan, (int)side, N, (uint32_t)flags) != 0;
  }
  int64_t last_path() {
    int path = 0;
    check(b200_plan_tc_info(plan, 0, nullptr, nullptr, &path));
    return path;
  }
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "B200-native drop-in for the reference's custom_mm extension";
  m.def("init_cublas", &init_cublas_handle, "Create cuBLAS handle.");
  m.def("destroy_cublas", &destroy_cublas_handle, "Destroy cuBLAS handle.");
  m.def("init_cusparse", &init_engine, "Initialise the SpMM engine (name kept from the reference).");
  m.def("destroy_cusparse", &destroy_engine, "Release the SpMM engine's registries.");

  m.def("cublas_mmul", &cublas_mmul, "cuBLAS Torch Matrix Multiplication");
  m.def("cublas_bmm", &cublas_bmm, "cuBLAS Batched Torch Matrix Multiplication");
  m.def("cusparse_mmul", &cusparse_mmul, "CSR SpMM (TMA-staged merge-path kernel)");

  m.def("dummy_kernel", &dummy_kernel_launch, "Launch dummy kernel.");
  m.def("naive_spmm", &naive_spmm, "CSR SpMM (direct-load merge-path kernel)");

  m.def("tiledspmm_inspect_csr", &tiledspmm_inspect_csr, "Inspect function for TiledSpMM with CSR input");
  m.def("tiledspmm_inspect_coo", &tiledspmm_inspect_coo, "Inspect function for TiledSpMM with COO input");
  m.def("tiledspmm_mm", &tiledspmm_mm, "MM function for TiledSpMM", py::arg("B"), py::arg("C"),
        py::arg("layer"), py::arg("row_major") = false);
  m.def("tiledspmm_clean", &tiledspmm_clean, "Cleanup function for TiledSpMM");

  m.def("cusparse_inspect", &cusparse_inspect, "Inspect function for CuSPARSE with CSR input");
  m.def("cusparse_mmul_opt", &cusparse_mmul_opt, "MM function for CuSPARSE", py::arg("B"), py::arg("C"),
        py::arg("layer"), py::arg("row_major") = false);
  m.def("cusparse_clean", &cusparse_clean, "Cleanup function for CuSPARSE");

  // engine extras
  m.def("dense_to_csr", &dense_to_csr, "dense [M,K] -> (crow int32, col int32, val) on the device",
        py::arg("dense"), py::arg("compact") = false);
  m.def("set_onepass_min_elems", [](int64_t n) { g_onepass_min_elems = n; g_last_kept_fraction = 0.0; },
        "one-pass dense->CSR for matrices of at least n elements (< 0: never); also forgets the remembered density");
  m.def("dense_to_csr_block_diag", &dense_to_csr_block_diag,
        "stack of L dense [M,K] matrices given as [L*M,K] -> block-diagonal CSR (crow, col, val) of [L*M, L*K]");
  m.def("csr_transpose", &csr_transpose, "CSR(A) -> (crow, col, val, perm) of A^T");
  m.def("gather", &gather, "out[q] = src[perm[q]]");
  m.def("sddmm", &sddmm, "grad of CSR values: sum_j G[row,j] * B[col,j] on A's pattern");
  m.def("kernel_launch_count", []() { return (uint64_t)b200_kernel_launch_count(); });
  m.def("set_option", [](const std::string& n, int64_t v) { check(b200_set_option(n.c_str(), v)); });
  m.def("get_option", [](const std::string& n) { return b200_get_option(n.c_str()); });
  m.def("build_info", []() { return std::string(b200_build_info()); });
  py::class_<PlanHandle>(m, "Plan", "inspected CSR operand (b200_plan): forward / backward on cached state")
      .def(py::init<torch::Tensor, torch::Tensor, torch::Tensor, int64_t, int64_t>(), py::arg("crow"), py::arg("col"),
           py::arg("val"), py::arg("rows"), py::arg("cols"))
      .def("forward", &PlanHandle::forward, py::arg("B"), py::arg("C"), py::arg("bias") = py::none(),
           py::arg("flags") = 0)
      .def("backward_B", &PlanHandle::backward_B, py::arg("G"), py::arg("dB"), py::arg("flags") = 0)
      .def("backward_A", &PlanHandle::backward_A, py::arg("G"), py::arg("B"), py::arg("flags") = 0)
      .def("fc_forward", &PlanHandle::fc_forward, py::arg("x"), py::arg("bias") = py::none(), py::arg("relu") = false,
           py::arg("small") = false)
      .def("fc_backward", &PlanHandle::fc_backward)
      .def("values_updated", &PlanHandle::values_updated)
      .def("would_use_tc", &PlanHandle::would_use_tc, py::arg("N"), py::arg("side") = 0, py::arg("flags") = 0)
      .def("last_path", &PlanHandle::last_path);
  m.attr("OP_B_TRANSPOSED") = (int)B200_OP_B_TRANSPOSED;
  m.attr("OP_C_TRANSPOSED") = (int)B200_OP_C_TRANSPOSED;
  m.attr("OP_RELU") = (int)B200_OP_RELU;
  m.attr("OP_NO_TC") = (int)B200_OP_NO_TC;
  m.attr("OP_FORCE_TC") = (int)B200_OP_FORCE_TC;
}
