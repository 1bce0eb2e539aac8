// are the DFMA pipe and the DMMA sub-pipe independent?  half the warps of each CTA do DFMA, half DMMA
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_mix(double* out, int iters, int mode) {   // mode 0: all DMMA, 1: all DFMA, 2: even warps DMMA, odd warps DFMA
  const int warp = threadIdx.x >> 5;
  const bool mma = mode == 0 || (mode == 2 && (warp & 1) == 0);
  double s = 0;
  if (mma) {
    double c[8][2];
    for (int t = 0; t < 8; ++t) { c[t][0] = 0; c[t][1] = 0; }
    double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int t = 0; t < 8; ++t)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
    }
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  } else {
    double c[16];
    for (int t = 0; t < 16; ++t) c[t] = 0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int i = 0; i < iters * 4; ++i) {     // 4x iterations: 64 FMA per thread per outer iteration, like 8 DMMA
#pragma unroll
      for (int t = 0; t < 16; ++t) c[t] = fma(a, b, c[t]);
    }
    for (int t = 0; t < 16; ++t) s += c[t];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* d; cudaMalloc(&d, sizeof(double) * 148 * 1024);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const double clk = pr.clockRate * 1e3;
  const int warps = 16, iters = 10000;
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_mix<<<pr.multiProcessorCount, warps * 32>>>(d, iters, mode);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    // every thread does iters * 64 FMA-equivalents in both branches (8 DMMA x 256 / 32 = 64; 4 x 16 = 64)
    const double total = (double)iters * 64 * warps * 32 * pr.multiProcessorCount;
    printf("mode %d (%s): %.3f ms  %.1f FMA/clk/SM\n", mode, mode == 0 ? "all DMMA" : mode == 1 ? "all DFMA" : "half DMMA / half DFMA",
           ms, total / (ms * 1e-3) / clk / pr.multiProcessorCount);
  }
  return 0;
}
