// Latency of the 64x64 factor+inverse (kb_diag.cuh) alone and under FP64-pipe contention from
// 8 co-resident warps issuing back-to-back DMMA (development aid).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define KB_TILE 64
#define KB_THREADS 256
#define KB_LDT 12
#define KB_DIAG_SYNC() asm volatile("bar.sync 1, 256;" ::: "memory")
__device__ long long g_stamp[16];
#define KB_DIAG_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_stamp[i] = clock64(); } while (0)
#include "../../bayesianopttools_b200/csrc/kb_diag.cuh"
__device__ __forceinline__ void kb_diag_factor_inverse_v0(double (*Cs)[KB_LDC], double (*Ds)[KB_LDC],
                                                       double* Ts_base, int* info, int pivot_base) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const unsigned FULL = 0xffffffffu;
  for (int e = tid; e < KB_TILE * KB_TILE; e += KB_THREADS) Ds[e >> 6][e & 63] = 0.0;
  KB_DIAG_SYNC();
  for (int jb = 0; jb < 4; ++jb) {
    const int c0 = jb * 16;
    {
      // 16x16 diagonal block: factor AND invert with one element of each per thread
      // (thread (r,c) keeps A[r][c] and W[r][c] in registers).  Per column j the pivot column
      // and the finished inverse row travel through a double-buffered 32-double shared
      // exchange: ONE block barrier per column, and every thread redoes the 4-op rsqrt so the
      // serial chain is  barrier -> rsqrt -> scale -> fma.  (A single-warp register/shuffle
      // version was issue-bound at ~4 clk per instruction.)
      const int r = tid >> 4, c = tid & 15;
      double a = Cs[c0 + r][c0 + c];
      double w = (r == c) ? 1.0 : 0.0;
      double* xch = Ts_base;                       // [2][32]: column of A | row of W
#pragma unroll 1
      for (int j = 0; j < 16; ++j) {
        double* buf = xch + (j & 1) * 32;
        if (c == j) buf[r] = a;
        if (r == j) buf[16 + c] = w;
        KB_DIAG_SYNC();
        double piv = buf[j];
        if (!(piv > 0.0)) {          // not positive definite (or NaN): flag, keep the batch alive
          if (tid == 0) atomicCAS(info, 0, pivot_base + c0 + j + 1);
          piv = 1.0;
        }
        const double inv = kb_rsqrt(piv);
        const double lr = (r == j) ? piv * inv : buf[r] * inv;     // L[r][j]
        const double lc = buf[c] * inv;                            // L[c][j]
        const double xj = buf[16 + c] * inv;                       // X[j][c]
        if (c == j) a = lr;
        else if (c > j) a = fma(-lr, lc, a);
        if (r == j) w = xj;
        else if (r > j) w = fma(-lr, xj, w);
      }
      if (c <= r) Cs[c0 + r][c0 + c] = a;
      Ds[c0 + r][c0 + c] = (c <= r) ? w : 0.0;
    }
    KB_DIAG_SYNC();
    if (jb == 3) break;
    // panel: rows below the block, P = A * inv(L_D)^T ; one warp per 8 rows, both 8-column halves
    const int nrb = (KB_TILE - 16 - c0) / 8;
    if (warp < nrb) {
      const int r0 = c0 + 16 + 8 * warp;
      double af[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) af[ks] = Cs[r0 + gq][c0 + ks * 4 + tq];
      double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) kb_dmma(p[ni][0], p[ni][1], af[ks], Ds[c0 + ni * 8 + gq][c0 + ks * 4 + tq]);
      __syncwarp();
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
        *reinterpret_cast<double2*>(&Cs[r0 + gq][c0 + ni * 8 + 2 * tq]) = make_double2(p[ni][0], p[ni][1]);
    }
    KB_DIAG_SYNC();
    // trailing update of the lower 8x8 tiles: C -= P P^T
    const int mt = nrb, ntile = mt * (mt + 1) / 2;
    for (int t = warp; t < ntile; t += 8) {
      int ri = 0;
      while ((ri + 1) * (ri + 2) / 2 <= t) ++ri;
      const int ci = t - ri * (ri + 1) / 2;
      const int r0 = c0 + 16 + 8 * ri, cc0 = c0 + 16 + 8 * ci;
      double2 cv = *reinterpret_cast<double2*>(&Cs[r0 + gq][cc0 + 2 * tq]);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        kb_dmma(cv.x, cv.y, -Cs[r0 + gq][c0 + ks * 4 + tq], Cs[cc0 + gq][c0 + ks * 4 + tq]);
      *reinterpret_cast<double2*>(&Cs[r0 + gq][cc0 + 2 * tq]) = cv;
    }
    KB_DIAG_SYNC();
  }
  // off-diagonal 16x16 blocks of the inverse, by block distance:
  //   X_ij = -inv(L_ii) * sum_{k=j}^{i-1} L_ik X_kj
  double* Ts = Ts_base + warp * 16 * KB_LDT;
  for (int delta = 1; delta <= 3; ++delta) {
    const int nblk = 4 - delta;
    if (warp < 2 * nblk) {
      const int j = warp >> 1, h = warp & 1, i = j + delta;
      double t0[2] = {0.0, 0.0}, t1[2] = {0.0, 0.0};
      for (int kb = j; kb < i; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double bv = Ds[kb * 16 + ks * 4 + tq][j * 16 + h * 8 + gq];
          kb_dmma(t0[0], t0[1], Cs[i * 16 + gq][kb * 16 + ks * 4 + tq], bv);
          kb_dmma(t1[0], t1[1], Cs[i * 16 + 8 + gq][kb * 16 + ks * 4 + tq], bv);
        }
      *reinterpret_cast<double2*>(&Ts[gq * KB_LDT + 2 * tq]) = make_double2(t0[0], t0[1]);
      *reinterpret_cast<double2*>(&Ts[(8 + gq) * KB_LDT + 2 * tq]) = make_double2(t1[0], t1[1]);
      __syncwarp();
      double x0[2] = {0.0, 0.0}, x1[2] = {0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const double bv = Ts[(ks * 4 + tq) * KB_LDT + gq];
        kb_dmma(x0[0], x0[1], Ds[i * 16 + gq][i * 16 + ks * 4 + tq], bv);
        kb_dmma(x1[0], x1[1], Ds[i * 16 + 8 + gq][i * 16 + ks * 4 + tq], bv);
      }
      *reinterpret_cast<double2*>(&Ds[i * 16 + gq][j * 16 + h * 8 + 2 * tq]) = make_double2(-x0[0], -x0[1]);
      *reinterpret_cast<double2*>(&Ds[i * 16 + 8 + gq][j * 16 + h * 8 + 2 * tq]) = make_double2(-x1[0], -x1[1]);
    }
    KB_DIAG_SYNC();
  }
}


__global__ void __launch_bounds__(512, 1) bench(int variant, int loaded, int reps, const double* A0, double* Lout,
                                                double* Dout, long long* cyc, int* info, double* sink) {
  extern __shared__ __align__(16) double sm[];
  double (*Cs)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(sm);
  double (*Ds)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(sm + 64 * KB_LDC);
  double* xch = sm + 2 * 64 * KB_LDC;
  volatile int* stop = reinterpret_cast<volatile int*>(xch + 1536);
  if (threadIdx.x == 0) *stop = 0;
  __syncthreads();
  if (threadIdx.x < 256) {
    long long tot = 0;
    for (int rep = 0; rep < reps; ++rep) {
      for (int e = threadIdx.x; e < 4096; e += 256) Cs[e >> 6][e & 63] = A0[e];
      KB_DIAG_SYNC();
      const long long t0 = clock64();
      if (variant == 0) kb_diag_factor_inverse_v0(Cs, Ds, xch, info, 0);
      else kb_diag_factor_inverse(Cs, Ds, xch, info, 0);
      tot += clock64() - t0;
    }
    if (threadIdx.x == 0) { cyc[blockIdx.x] = tot / reps; *stop = 1; }
    if (blockIdx.x == 0)
      for (int e = threadIdx.x; e < 4096; e += 256) {
        Lout[e] = ((e & 63) <= (e >> 6)) ? Cs[e >> 6][e & 63] : 0.0;
        Dout[e] = Ds[e >> 6][e & 63];
      }
  } else if (loaded) {
    double d[8][2];
    for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 0.999;
    while (!*stop) {
#pragma unroll
      for (int it = 0; it < 8; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) kb_dmma(d[i][0], d[i][1], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1];
    if (s == 123.456) sink[threadIdx.x] = s;
  }
}

int main() {
  const int n = 64;
  std::vector<double> M(n * n), A(n * n), L(n * n, 0.0), W(n * n, 0.0);
  srand(7);
  for (auto& v : M) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = (i == j) ? 2.0 : 0.0;
      for (int k = 0; k < n; ++k) s += M[i * n + k] * M[j * n + k];
      A[i * n + j] = s;
    }
  // host reference
  std::vector<double> T = A;
  for (int j = 0; j < n; ++j) {
    double s = T[j * n + j];
    for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
    L[j * n + j] = sqrt(s);
    for (int i = j + 1; i < n; ++i) {
      double t = T[i * n + j];
      for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / L[j * n + j];
    }
  }
  for (int c = 0; c < n; ++c)
    for (int i = c; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= L[i * n + k] * W[k * n + c];
      W[i * n + c] = s / L[i * n + i];
    }
  double *dA, *dL, *dD, *sink; long long* cyc; int* info;
  cudaMalloc(&dA, 8 * 4096); cudaMalloc(&dL, 8 * 4096); cudaMalloc(&dD, 8 * 4096); cudaMalloc(&sink, 8 * 512);
  cudaMalloc(&cyc, 8 * 1024); cudaMalloc(&info, 4); cudaMemset(info, 0, 4);
  cudaMemcpy(dA, A.data(), 8 * 4096, cudaMemcpyHostToDevice);
  const size_t smem = (2 * 64 * KB_LDC + 1536 + 2) * sizeof(double);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int variant = 0; variant < 2; ++variant)
    for (int loaded = 0; loaded < 1; ++loaded)
      for (int grid : {1, 148}) {
        bench<<<grid, 512, smem>>>(variant, loaded, 20, dA, dL, dD, cyc, info, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(grid);
        std::vector<double> hl(4096), hd(4096);
        cudaMemcpy(h.data(), cyc, 8 * grid, cudaMemcpyDeviceToHost);
        cudaMemcpy(hl.data(), dL, 8 * 4096, cudaMemcpyDeviceToHost);
        cudaMemcpy(hd.data(), dD, 8 * 4096, cudaMemcpyDeviceToHost);
        int hinfo; cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
        double el = 0, ed = 0, s = 0;
        for (int i = 0; i < 4096; ++i) { el = fmax(el, fabs(hl[i] - L[i])); if (((i & 63) >> 3) <= ((i >> 6) >> 3) || ((i & 63) >> 4) == ((i >> 6) >> 4)) ed = fmax(ed, fabs(hd[i] - W[i])); }
        for (int i = 0; i < grid; ++i) s += h[i];
        if (variant == 1) {
          long long st[16]; cudaMemcpyFromSymbol(st, g_stamp, sizeof(st));
          printf("  stamps:"); for (int i = 1; i <= 12; ++i) printf(" %lld", st[i] - st[i - 1]); printf("\n");
        }
        printf("variant %d loaded %d grid %3d: %8.0f cyc (%.2f us @1.965GHz)  errL %.2e errW %.2e info %d\n", variant,
               loaded, grid, s / grid, s / grid / 1965.0, el, ed, hinfo);
      }
  return 0;
}
