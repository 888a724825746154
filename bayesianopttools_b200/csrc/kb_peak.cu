// krig-b200: on-device roofline denominators that MEASURED_PEAKS.json does not
// carry: FP64 tensor (DMMA via mma.sync m8n8k4 f64) and scalar DFMA throughput.
#include "kb_common.cuh"
#include "kb_internal.h"

__global__ void __launch_bounds__(512) kb_peak_dmma_kernel(int iters, double* sink) {
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) kb_dmma(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(512) kb_peak_dfma_kernel(int iters, double* sink) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 1e-3 * (threadIdx.x + i);
  const double a = 0.999999, b = 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456) sink[0] = s;
}

cudaError_t kb_launch_peak(cudaStream_t st, int which, int iters, int grid, double* sink) {
  if (which == 0) kb_peak_dmma_kernel<<<grid, 512, 0, st>>>(iters, sink);
  else kb_peak_dfma_kernel<<<grid, 512, 0, st>>>(iters, sink);
  return cudaGetLastError();
}
