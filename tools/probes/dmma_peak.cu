// dmma_peak.cu — development probe: what mma.sync.m8n8k4.f64 (SASS DMMA) delivers on this GPU, against plain DFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int ACC>
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double d[ACC][2];
#pragma unroll
  for (int i = 0; i < ACC; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i * 1e-9; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += d[i][0] + d[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double d[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) d[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) d[i] = fma(d[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* out;
  CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_dmma<16><<<148 * 2, warps * 16>>>(out, iters, 1.0000001, 0.9999999);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double flops = 148.0 * 2 * (warps / 2.0) * iters * 16 * 512;
      if (rep) printf("{\"probe\": \"DMMA m8n8k4\", \"warps_per_sm\": %d, \"ms\": %.3f, \"TFLOPs\": %.2f}\n", warps, ms, flops / ms / 1e9);
    }
  }
  for (int warps : {8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_dfma<16><<<148 * 2, warps * 16>>>(out, iters, 1.0000001, 0.9999999);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double flops = 148.0 * 2 * (warps * 16.0) * iters * 16 * 2;
      if (rep) printf("{\"probe\": \"DFMA\", \"warps_per_sm\": %d, \"ms\": %.3f, \"TFLOPs\": %.2f}\n", warps, ms, flops / ms / 1e9);
    }
  }
  CK(cudaGetLastError());
  return 0;
}
