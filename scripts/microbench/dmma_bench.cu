// micro-benchmark: FP64 tensor-core mma.sync m8n8k4 vs DFMA throughput on sm_100a (per SM per clock)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  for (int t = 0; t < 8; ++t) { c[t][0] = 0; c[t][1] = 0; }
  double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, int iters) {
  double c[16];
  for (int t = 0; t < 16; ++t) c[t] = 0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 16; ++t) c[t] = fma(a, b, c[t]);
  }
  double s = 0; for (int t = 0; t < 16; ++t) s += c[t];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* d; cudaMalloc(&d, sizeof(double) * 148 * 16 * 1024);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const double clk = pr.clockRate * 1e3;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int which = 0; which < 2; ++which) {
      const int iters = 20000;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_dmma<<<pr.multiProcessorCount, warps * 32>>>(d, iters);
        else k_dfma<<<pr.multiProcessorCount, warps * 32>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma_per_thread = which == 0 ? (double)iters * 8 * 256 / 32 : (double)iters * 16;
      const double total = fma_per_thread * warps * 32 * pr.multiProcessorCount;
      printf("%s warps/SM %2d: %.3f ms  %.1f FMA/clk/SM (at %.0f MHz nominal)  %.2f TFLOP/s\n", which == 0 ? "DMMA m8n8k4" : "DFMA       ",
             warps, ms, total / (ms * 1e-3) / clk / pr.multiProcessorCount, clk / 1e6, 2 * total / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
