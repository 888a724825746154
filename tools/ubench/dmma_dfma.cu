// Do DMMA (tensor) and DFMA (scalar FP64) share one pipe?  Even warps stream DMMA, odd warps stream DFMA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k(int iters, int mode, double* sink) {   // mode 0: all DMMA, 1: all DFMA, 2: even DMMA / odd DFMA
  const int warp = threadIdx.x >> 5;
  const bool tensor = (mode == 0) || (mode == 2 && (warp & 1) == 0);
  double d[8][2];
  for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = 1e-3 * i;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 0.999;
  if (tensor) {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(d[i][0], d[i][1], a, b);
  } else {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) { d[i][0] = fma(d[i][0], a, b); d[i][1] = fma(d[i][1], a, b); }
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1];
  if (s == 123.456) sink[threadIdx.x] = s;
}
int main() {
  double* sink; cudaMalloc(&sink, 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148, threads = 512;
  for (int mode = 0; mode < 3; ++mode) {
    k<<<grid, threads>>>(1000, mode, sink); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<grid, threads>>>(iters, mode, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = grid * (threads / 32);
    const double fl_t = 8.0 * 512.0 * iters, fl_s = 16.0 * 2 * 32 * iters;   // per warp
    double tf;
    if (mode == 0) tf = warps * fl_t;
    else if (mode == 1) tf = warps * fl_s;
    else tf = warps / 2 * (fl_t + fl_s);
    printf("mode %d: %.3f ms, total %.2f TFLOP/s%s\n", mode, ms, tf / (ms * 1e-3) / 1e12,
           mode == 2 ? " (half the warps DMMA, half DFMA; equal iteration counts)" : "");
  }
  return 0;
}
