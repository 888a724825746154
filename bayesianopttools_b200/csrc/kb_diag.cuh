// krig-b200: factor + invert one 64x64 SPD tile held in shared memory (the serial step of the
// Cholesky chain, np.linalg.cholesky at likelihood_func.py:175 restricted to a diagonal tile).
// Shared by kb_chol.cu and tools/ubench/diagf.cu.  The 256 calling threads synchronise with
// KB_DIAG_SYNC() (default: the CTA barrier).
#pragma once
#include "kb_common.cuh"

#ifndef KB_DIAG_SYNC
#define KB_DIAG_SYNC() __syncthreads()
#endif
#define KB_LDC 68       // padded row stride of a 64 x 64 tile (== 4 mod 16: conflict-free DMMA fragments)
#define KB_LDT 12       // row stride of the per-warp 16x8 scratch tile (12 == 12 mod 16: conflict free)
#define KB_DIAG_SCRATCH (8 * 16 * KB_LDT)   // doubles of scratch behind Ds
#ifndef KB_DIAG_STAMP
#define KB_DIAG_STAMP(i)
#endif

// 1/sqrt(p) to full FP64 accuracy with a 4-deep dependent chain: hardware approximation
// (relative error ~2^-23) + one third-order step  y (1 + e/2 + 3e^2/8),  e = 1 - p y^2.
__device__ __forceinline__ double kb_rsqrt(double p) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  const double py = p * y;
  const double e = fma(-py, y, 1.0);
  const double ye = y * e;
  const double q = fma(e, 0.375, 0.5);
  return fma(ye, q, y);
}

// 64x64 SPD tile in Cs (lower part used) -> Cs = L (lower), Ds = L^-1 (lower, zeros above).
// Right-looking with 16-wide panels; every phase is shared by all 8 warps:
//   1. the 16x16 diagonal block is factored AND inverted with one element of each per thread;
//      column j and row j of the inverse travel through a double-buffered shared exchange with
//      ONE barrier per column.  The NEXT pivot's 1/sqrt is computed by every thread from the
//      exchanged column/diagonal (look-ahead), so the rsqrt chain overlaps the exchange instead
//      of adding to it.
//   2. panel  P = A W_D^T  and the finished block row of the inverse  W_i,: = W_D S_i,:   (DMMA)
//   3. trailing  C -= P P^T  and the running sums  S -= P W_i,:                           (DMMA)
// where S_ij = -sum_k L_ik W_kj accumulates in the not-yet-final part of Ds.
__device__ __forceinline__ void kb_diag_factor_inverse(double (*Cs)[KB_LDC], double (*Ds)[KB_LDC],
                                                       double* xch, int* info, int pivot_base) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  KB_DIAG_STAMP(0);
  for (int e = tid; e < KB_TILE * KB_TILE; e += KB_THREADS) Ds[e >> 6][e & 63] = 0.0;
  KB_DIAG_SYNC();
  KB_DIAG_STAMP(1);
  for (int jb = 0; jb < 4; ++jb) {
    const int c0 = jb * 16;
    {
      const int r = tid >> 4, c = tid & 15;
      double a = Cs[c0 + r][c0 + c];
      double w = (r == c) ? 1.0 : 0.0;
      double piv = Cs[c0][c0];
      if (!(piv > 0.0)) {            // not positive definite (or NaN): flag, keep the batch alive
        if (tid == 0) atomicCAS(info, 0, pivot_base + c0 + 1);
        piv = 1.0;
      }
      double inv = kb_rsqrt(piv);
#pragma unroll 1
      for (int j = 0; j < 16; ++j) {
        double* buf = xch + (j & 1) * 48;          // column j of A | row j of W | diagonal of A
        if (c == j) buf[r] = a;
        if (r == j) buf[16 + c] = w;
        if (r == c) buf[32 + r] = a;
        KB_DIAG_SYNC();
        const double colr = buf[r], colc = buf[c], wj = buf[16 + c];
        double invn = 1.0;
        if (j < 15) {                // look-ahead: pivot j+1 after the rank-1 update of column j
          const double ln = buf[j + 1] * inv;
          double pn = fma(-ln, ln, buf[32 + j + 1]);
          if (!(pn > 0.0)) {
            if (tid == 0) atomicCAS(info, 0, pivot_base + c0 + j + 2);
            pn = 1.0;
          }
          invn = kb_rsqrt(pn);
        }
        const double lr = colr * inv;              // L[r][j]  (r == j: p * rsqrt(p) = sqrt(p))
        const double lc = colc * inv;              // L[c][j]
        const double xj = wj * inv;                // W[j][c]
        if (c == j) a = lr;
        else if (c > j) a = fma(-lr, lc, a);
        if (r == j) w = xj;
        else if (r > j) w = fma(-lr, xj, w);
        inv = invn;
      }
      if (c <= r) Cs[c0 + r][c0 + c] = a;
      Ds[c0 + r][c0 + c] = (c <= r) ? w : 0.0;
    }
    KB_DIAG_SYNC();
    KB_DIAG_STAMP(2 + jb * 3);
    const int nrb = (KB_TILE - 16 - c0) / 8;       // 8-row blocks below the diagonal block
    // ---- phase 2: panel rows (warps 0..nrb-1) and the finished inverse block row (next 2*jb warps)
    if (warp < nrb) {
      const int r0 = c0 + 16 + 8 * warp;
      double af[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) af[ks] = Cs[r0 + gq][c0 + ks * 4 + tq];
      double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) kb_dmma(p[ni][0], p[ni][1], af[ks], Ds[c0 + ni * 8 + gq][c0 + ks * 4 + tq]);
      __syncwarp();
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
        *reinterpret_cast<double2*>(&Cs[r0 + gq][c0 + ni * 8 + 2 * tq]) = make_double2(p[ni][0], p[ni][1]);
    } else if (warp < nrb + 2 * jb) {
      const int cb = warp - nrb;                   // 8-column block of the finished part
      double x[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const double bv = Ds[c0 + ks * 4 + tq][cb * 8 + gq];
        if (ks < 2) kb_dmma(x[0][0], x[0][1], Ds[c0 + gq][c0 + ks * 4 + tq], bv);
        kb_dmma(x[1][0], x[1][1], Ds[c0 + 8 + gq][c0 + ks * 4 + tq], bv);
      }
      __syncwarp();
#pragma unroll
      for (int rb = 0; rb < 2; ++rb)
        *reinterpret_cast<double2*>(&Ds[c0 + rb * 8 + gq][cb * 8 + 2 * tq]) = make_double2(x[rb][0], x[rb][1]);
    }
    if (jb == 3) break;
    KB_DIAG_SYNC();
    KB_DIAG_STAMP(3 + jb * 3);
    // ---- phase 3: trailing update of A (lower 8x8 tiles) and of the running sums S ----
    const int ntile = nrb * (nrb + 1) / 2, ncb = 2 * jb + 2, total = ntile + nrb * ncb;
    for (int t = warp; t < total; t += 8) {
      if (t < ntile) {
        int ri = 0;
        while ((ri + 1) * (ri + 2) / 2 <= t) ++ri;
        const int ci = t - ri * (ri + 1) / 2;
        const int r0 = c0 + 16 + 8 * ri, cc0 = c0 + 16 + 8 * ci;
        double2 cv = *reinterpret_cast<double2*>(&Cs[r0 + gq][cc0 + 2 * tq]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          kb_dmma(cv.x, cv.y, -Cs[r0 + gq][c0 + ks * 4 + tq], Cs[cc0 + gq][c0 + ks * 4 + tq]);
        *reinterpret_cast<double2*>(&Cs[r0 + gq][cc0 + 2 * tq]) = cv;
      } else {
        const int u = t - ntile, ri = u / ncb, cb = u - ri * ncb;
        const int r0 = c0 + 16 + 8 * ri;
        double2 sv = *reinterpret_cast<double2*>(&Ds[r0 + gq][cb * 8 + 2 * tq]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          kb_dmma(sv.x, sv.y, -Cs[r0 + gq][c0 + ks * 4 + tq], Ds[c0 + ks * 4 + tq][cb * 8 + gq]);
        *reinterpret_cast<double2*>(&Ds[r0 + gq][cb * 8 + 2 * tq]) = sv;
      }
    }
    KB_DIAG_SYNC();
    KB_DIAG_STAMP(4 + jb * 3);
  }
  KB_DIAG_SYNC();
  KB_DIAG_STAMP(12);
}
