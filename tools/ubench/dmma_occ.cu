// DMMA throughput vs warps per SM sub-partition and independent accumulators per warp (development aid).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(int iters, double* sink) {
  double d[NACC][2];
  for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 0.999;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(d[i][0], d[i][1], a, b);
  double s = 0; for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1];
  if (s == 123.456) sink[threadIdx.x] = s;
}
template <int NACC> void run(int threads, double* sink) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148;
  k<NACC><<<grid, threads>>>(1000, sink); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<NACC><<<grid, threads>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flop = (double)grid * (threads / 32) * iters * NACC * 512.0;
  printf("warps/SM %2d (per SMSP %.1f) NACC %2d: %.2f TFLOP/s\n", threads / 32, threads / 128.0, NACC, flop / (ms * 1e-3) / 1e12);
}
int main() {
  double* sink; cudaMalloc(&sink, 8 * 1024);
  for (int threads : {128, 256, 512, 1024}) { run<1>(threads, sink); run<2>(threads, sink); run<4>(threads, sink); run<8>(threads, sink); run<16>(threads, sink); }
  return 0;
}
