// transpose.cu — CSR(A) -> CSR(A^T) on the device, rows of A^T sorted by original
// row id (stable), i.e. exactly what `m1.transpose(-1,-2).to_sparse_csr()` builds on
// every backward call in the reference (matmuls.py:253).  The engine builds it once
// per plan and caches it; grad_B = A^T * grad_C then runs through the forward kernel.
//
// Inspector-time code (runs once per sparsity pattern): a stable LSB radix sort of
// the non-zero positions by column (CUB, keys limited to ceil(log2 K) bits), a
// binary search per column for the new row_ptr, and a gather of (row, value).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace b200 {
namespace {

__global__ void iota_kernel(int* out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i;
}

// t_row_ptr[c] = first position in sorted_col with value >= c
__global__ void colptr_kernel(const int* __restrict__ sorted_col, int nnz, int K,
                              int* __restrict__ t_row_ptr) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= K; c += gridDim.x * blockDim.x) {
    int lo = 0, hi = nnz;
    while (lo < hi) {
      int mid = (int)(((long long)lo + hi) >> 1);
      if (__ldg(sorted_col + mid) >= c) hi = mid; else lo = mid + 1;
    }
    t_row_ptr[c] = lo;
  }
}

// t_col[q] = row of original position perm[q]; t_val[q] = val[perm[q]]
__global__ void gather_rows_kernel(const int* __restrict__ row_ptr, int M, const int* __restrict__ perm,
                                   const float* __restrict__ val, int nnz,
                                   int* __restrict__ t_col, float* __restrict__ t_val) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += gridDim.x * blockDim.x) {
    const int p = __ldg(perm + q);
    int lo = 0, hi = M - 1;  // largest r with row_ptr[r] <= p  (rows may be empty)
    while (lo < hi) {
      int mid = (int)(((long long)lo + hi + 1) >> 1);
      if (__ldg(row_ptr + mid) <= p) lo = mid; else hi = mid - 1;
    }
    t_col[q] = lo;
    if (val != nullptr) t_val[q] = __ldg(val + p);
  }
}

__global__ void gather_kernel(const float* __restrict__ src, const int* __restrict__ perm, long long n,
                              float* __restrict__ dst) {
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n;
       q += (long long)gridDim.x * blockDim.x)
    dst[q] = __ldg(src + __ldg(perm + q));
}

inline int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int launch_gather(const float* src, const int* perm, long long n, float* dst, cudaStream_t s) {
  if (n == 0) return B200_OK;
  gather_kernel<<<grid_for(n), 256, 0, s>>>(src, perm, n, dst);
  B200_LAUNCH_OK("gather_kernel");
  return B200_OK;
}

int launch_transpose(const int* row_ptr, const int* col, const float* val, int M, int K, int nnz,
                     int* t_row_ptr, int* t_col, float* t_val, int* t_perm, cudaStream_t s) {
  if (nnz == 0) {
    B200_CUDA_OK(cudaMemsetAsync(t_row_ptr, 0, (size_t)(K + 1) * sizeof(int), s));
    return B200_OK;
  }
  int end_bit = 1;
  while (end_bit < 31 && (1LL << end_bit) < (long long)K) ++end_bit;

  size_t temp_bytes = 0;
  B200_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, col, (int*)nullptr, (int*)nullptr,
                                               (int*)nullptr, nnz, 0, end_bit, s));
  // one allocation: [sorted_col nnz][iota nnz][perm nnz (if caller gave none)][cub temp]
  const size_t n4 = ((size_t)nnz * sizeof(int) + 255) & ~(size_t)255;
  const size_t total = n4 * 3 + temp_bytes + 256;
  unsigned char* buf = nullptr;
  B200_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&buf), total, s));
  int* sorted_col = reinterpret_cast<int*>(buf);
  int* iota = reinterpret_cast<int*>(buf + n4);
  int* perm = t_perm ? t_perm : reinterpret_cast<int*>(buf + 2 * n4);
  void* temp = buf + 3 * n4;

  iota_kernel<<<grid_for(nnz), 256, 0, s>>>(iota, nnz);
  B200_LAUNCH_OK("iota_kernel");
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, col, sorted_col, iota, perm, nnz, 0,
                                                  end_bit, s);
  if (e != cudaSuccess) {
    cudaFreeAsync(buf, s);
    set_error("cub radix sort failed: %s", cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  colptr_kernel<<<grid_for((long long)K + 1), 256, 0, s>>>(sorted_col, nnz, K, t_row_ptr);
  B200_LAUNCH_OK("colptr_kernel");
  gather_rows_kernel<<<grid_for(nnz), 256, 0, s>>>(row_ptr, M, perm, val, nnz, t_col, t_val);
  B200_LAUNCH_OK("gather_rows_kernel");
  B200_CUDA_OK(cudaFreeAsync(buf, s));
  return B200_OK;
}

}  // namespace b200
