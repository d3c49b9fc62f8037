// l2bw.cu — microbenchmark behind DESIGN.md's "binding roof" numbers: how fast can 148 SMs pull
// 512-byte rows (one LDG.128 per lane) out of L2 / HBM, streamed or gathered at random?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2bw.bin tools/l2bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int U>
__global__ void __launch_bounds__(256) gather_rows(const float4* __restrict__ B, const int* __restrict__ idx,
                                                   long long n_idx, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long long i = warp * U; i + U <= n_idx; i += n_warps * U) {
    float4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldg(B + (long long)__ldg(idx + i + k) * 32 + lane);
#pragma unroll
    for (int k = 0; k < U; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
  }
  if (acc.x == 12345.678f) out[0] = acc;  // keep the loads alive
}

int main(int argc, char** argv) {
  const long long n_idx = 16LL << 20;  // 16M row reads of 512 B = 8 GiB per launch
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int mode = 0; mode < 3; ++mode) {            // 0 = sequential rows, 1 = random rows, 2 = R-MAT-like skew
    for (long long rows : {1LL << 16, 1LL << 17, 1LL << 18, 1LL << 20}) {  // 32 MB .. 512 MB of B
      if (mode == 2 && rows != (1LL << 20)) continue;
      std::vector<int> h(n_idx);
      unsigned long long s = 88172645463325252ULL;
      for (long long i = 0; i < n_idx; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        if (mode == 2) {  // every bit of the column index is 1 with probability ~0.24 (R-MAT a=.57 b=.19 c=.19)
          int c = 0;
          for (int b = 0; b < 20; ++b) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; c = (c << 1) | ((s & 0xffff) < 15729); }
          h[i] = c;
        } else
        h[i] = mode ? (int)(s % rows) : (int)(i % rows);
      }
      float4 *B, *out; int* idx;
      cudaMalloc(&B, rows * 512); cudaMalloc(&out, 64); cudaMalloc(&idx, n_idx * 4);
      cudaMemset(B, 0, rows * 512);
      cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
      for (int ctas_per_sm : {2, 3, 4, 8}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int it = 0; it < 2; ++it) gather_rows<8><<<sms * ctas_per_sm, 256>>>(B, idx, n_idx, out);
        cudaEventRecord(e0);
        for (int it = 0; it < 3; ++it) gather_rows<8><<<sms * ctas_per_sm, 256>>>(B, idx, n_idx, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        printf("%s B=%4lld MB ctas/sm=%d U=8: %.3f ms  %.2f TB/s  (%s)\n", mode == 2 ? "rmat-skew " : mode ? "random    " : "sequential",
               rows * 512 >> 20, ctas_per_sm, ms, n_idx * 512.0 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
      }
      cudaFree(B); cudaFree(out); cudaFree(idx);
    }
  }
  return 0;
}
