// DMMA main-loop structure test: 32x32 warp tiles fed from shared memory (8 LDS.64 per 16 DMMA), no barriers, no global
// traffic.  How close does the LDS->DMMA loop alone get to the DMMA peak at 2 / 4 warps per sub-partition?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MI, int NI, bool DBUF>
__global__ void __launch_bounds__(256, 2) k(int iters, double* sink) {
  __shared__ double sa[16][132];
  __shared__ double sb[16][132];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3, wr = warp >> 1, wc = warp & 1;
  for (int e = tid; e < 16 * 132; e += 256) { (&sa[0][0])[e] = 1.0 + e * 1e-9; (&sb[0][0])[e] = 0.5 + e * 1e-9; }
  __syncthreads();
  double acc[MI][NI][2];
  for (int mi = 0; mi < MI; ++mi) for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  if (!DBUF) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        double af[MI], bf[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) af[mi] = sa[ks * 4 + tq][wr * 8 * MI + mi * 8 + gq];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) bf[ni] = sb[ks * 4 + tq][wc * 8 * NI + ni * 8 + gq];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
      }
    }
  } else {
    double af[2][MI], bf[2][NI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) af[0][mi] = sa[tq][wr * 8 * MI + mi * 8 + gq];
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) bf[0][ni] = sb[tq][wc * 8 * NI + ni * 8 + gq];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int cur = ks & 1, nxt = cur ^ 1, kn = (ks + 1) & 3;
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) af[nxt][mi] = sa[kn * 4 + tq][wr * 8 * MI + mi * 8 + gq];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) bf[nxt][ni] = sb[kn * 4 + tq][wc * 8 * NI + ni * 8 + gq];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[cur][mi], bf[cur][ni]);
      }
    }
  }
  double s = 0; for (int mi = 0; mi < MI; ++mi) for (int ni = 0; ni < NI; ++ni) s += acc[mi][ni][0] + acc[mi][ni][1];
  if (s == 123.456) sink[tid] = s;
}
template <int MI, int NI, bool DBUF> void run(int ctas_per_sm, double* sink) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000, grid = 148 * ctas_per_sm;
  k<MI, NI, DBUF><<<grid, 256>>>(100, sink); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MI, NI, DBUF><<<grid, 256>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flop = (double)grid * 8 * iters * 4 * MI * NI * 512.0;
  printf("MI %d NI %d dbuf %d CTAs/SM %d: %.2f TFLOP/s\n", MI, NI, (int)DBUF, ctas_per_sm, flop / (ms * 1e-3) / 1e12);
}
int main() {
  double* sink; cudaMalloc(&sink, 8 * 1024);
  for (int c : {1, 2}) { run<4, 4, false>(c, sink); run<4, 4, true>(c, sink); run<2, 4, false>(c, sink); run<2, 4, true>(c, sink); }
  return 0;
}
