// krig-b200: one-tile designs (n <= 64, ordinary or low-order Kriging: at most 16 right-hand-side rows).
//
// KrigDemo, SOBO, MOBO and AK-MCS designs of the reference are this size, and there the general
// pipeline (decode -> build -> persistent Cholesky -> GLS: four launches, two memsets, a cooperative launch
// that brings up 296 CTAs for one tile) is pure launch latency.  Here ONE CTA per candidate decodes its
// hyper-parameters, evaluates the correlation tile straight into shared memory, factors and inverts it
// with the same in-shared routine the Cholesky chain uses (kb_diag.cuh), applies the inverse to the
// right-hand sides and forms the Schur block, and leaves in the augmented matrix exactly what the general
// path leaves (L, Z = [y F]^T L^-T, S = -Z Z^T), so the GLS kernel and everything after it are unchanged.
// References: likelihood_func.py:71-102 (decode, Psi), :171-180 (nugget, Cholesky, ln-det),
// :183-193 (the solves that Z replaces).
#include "kb_common.cuh"
#include "kb_internal.h"
#include "kb_diag.cuh"

#define KB_SMALL_Q 16

struct KbSmallSmem {
  double C[KB_TILE][KB_LDC];       // R, then L
  double D[KB_TILE][KB_LDC];       // L^-1 (lower blocks)
  double Y[KB_SMALL_Q][KB_LDC];    // right-hand-side rows  y, F columns
  double Z[KB_SMALL_Q][KB_LDC];    // Y L^-T
  double xch[KB_DIAG_SCRATCH];
  double gw[KB_MAX_D], ith[KB_MAX_D];
  double wk[4];
  double eps;
};

__global__ void __launch_bounds__(KB_THREADS)
kb_nll_small_kernel(KbModelDev m, KbAugGeom g, int nbhyp, const double* __restrict__ xpop, int nth, int trainvar,
                    int has_nugget, int has_weights, double fixed_nugget, int nkernel,
                    const int* __restrict__ kernel_ids, KbHyper hp, double* __restrict__ aug, int* __restrict__ info) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  KbSmallSmem& sm = *reinterpret_cast<KbSmallSmem*>(smem_raw);
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* x = xpop + (size_t)b * nbhyp;
  const int d = m.d, n = m.n;

  // ---- decode (as kb_decode_kernel; the per-candidate arrays the GLS kernel reads are written too) ----
  for (int dim = tid; dim < d; dim += KB_THREADS) {
    const double th = pow(10.0, x[dim]);
    const double ith = 1.0 / th, gw = 0.5 / (th * th);
    sm.gw[dim] = gw; sm.ith[dim] = ith;
    hp.gw[(size_t)b * d + dim] = gw;
    hp.ith[(size_t)b * d + dim] = ith;
  }
  if (tid == 0) {
    int pos = nth;
    double s2 = 0.0;
    if (trainvar) s2 = pow(10.0, x[pos++]);
    const double nug = has_nugget ? x[pos++] : fixed_nugget;
    double wk[4] = {0.0, 0.0, 0.0, 0.0};
    if (has_weights) {
      double sum = 0.0;
      for (int j = 0; j < nkernel; ++j) sum += x[pos + j];
      for (int j = 0; j < nkernel; ++j) wk[kernel_ids[j]] += x[pos + j] / sum;
    } else {
      wk[kernel_ids[0]] = 1.0;
    }
    for (int j = 0; j < 4; ++j) { sm.wk[j] = wk[j]; hp.wk[b * 4 + j] = wk[j]; }
    sm.eps = pow(10.0, nug);
    hp.eps[b] = sm.eps;
    hp.nugget[b] = nug;
    hp.sigma2[b] = s2;
    info[b] = 0;
  }
  __syncthreads();

  // ---- correlation tile (entries as kb_build_aug_kernel computes them) and right-hand-side rows ----
  {
    const int tr = tid >> 4, tc = tid & 15;
    KbPairAcc acc[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
        for (int e = 0; e < 2; ++e) kb_pair_init(acc[a][a2][e]);
    for (int dim = 0; dim < d; ++dim) {
      const double gw = sm.gw[dim], ith = sm.ith[dim];
      const double* xrow = m.XT + (size_t)dim * m.n_pad;
      double xi[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xi[a] = __ldg(xrow + tr * 4 + a);
      double2 xj[2];
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) xj[a2] = __ldg(reinterpret_cast<const double2*>(xrow + 2 * tc + 32 * a2));
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2) {
          kb_pair_dim(acc[a][a2][0], xi[a] - xj[a2].x, gw, ith, m.kmask);
          kb_pair_dim(acc[a][a2][1], xi[a] - xj[a2].y, gw, ith, m.kmask);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = tr * 4 + a;
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 2 * tc + 32 * a2 + e;
          double val;
          if (row < n && c < n) val = kb_pair_finish(acc[a][a2][e], sm.wk, m.kmask) + ((row == c) ? sm.eps : 0.0);
          else val = (row == c) ? 1.0 : 0.0;
          sm.C[row][c] = val;
        }
    }
    for (int e = tid; e < KB_SMALL_Q * KB_TILE; e += KB_THREADS) {
      const int q = e >> 6, c = e & 63;
      double val = 0.0;
      if (c < n) {
        if (q == 0) val = m.y[c];
        else if (q <= m.nind) val = m.F[(size_t)c * m.nind + (q - 1)];
      }
      sm.Y[q][c] = val;
    }
  }
  __syncthreads();

  // ---- L, L^-1 ----
  kb_diag_factor_inverse(sm.C, sm.D, sm.xch, &info[b], 0);
  __syncthreads();

  // ---- Z = Y L^-T (row j of Z: z_i = sum_{k <= i} Y[j][k] Linv[i][k]), then S = -Z Z^T (lower) ----
  {
    // four independent dot products per thread (rows i0, i0+16, i0+32, i0+48 of L^-1), interleaved;
    // L^-1 is only defined on and below its diagonal, so entries k > i are masked, not read
    const int j = tid >> 4, i0 = tid & 15;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < KB_TILE; ++k) {
      const double yk = sm.Y[j][k];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = i0 + 16 * t;
        if (k <= i) s[t] = fma(yk, sm.D[i][k], s[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) sm.Z[j][i0 + 16 * t] = s[t];
  }
  __syncthreads();
  // one step of residual correction, as the tile Cholesky's kb_refined_solve (kb_chol.cu): the product with the
  // explicit inverse alone leaves a residual ~ eps cond(L) |Y|.   r = Y - Z L^T (overwrites Y),   Z += r L^-T
  {
    const int j = tid >> 4, c0 = tid & 15;
    double r[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) r[t] = sm.Y[j][c0 + 16 * t];
    for (int i = 0; i < KB_TILE; ++i) {
      const double zi = sm.Z[j][i];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int c = c0 + 16 * t;
        if (i <= c) r[t] = fma(-zi, sm.C[c][i], r[t]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 4; ++t) sm.Y[j][c0 + 16 * t] = r[t];
  }
  __syncthreads();
  {
    const int j = tid >> 4, i0 = tid & 15;
    double s[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) s[t] = sm.Z[j][i0 + 16 * t];
    for (int k = 0; k < KB_TILE; ++k) {
      const double rk = sm.Y[j][k];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = i0 + 16 * t;
        if (k <= i) s[t] = fma(rk, sm.D[i][k], s[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) sm.Z[j][i0 + 16 * t] = s[t];
  }
  __syncthreads();
  double* A = aug + (size_t)b * g.stride;
  const int ld = g.ld;
  {
    const int a = tid >> 4, c = tid & 15;
    if (c <= a) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int k = 0; k < KB_TILE; k += 4) {
        s0 = fma(sm.Z[a][k], sm.Z[c][k], s0);
        s1 = fma(sm.Z[a][k + 1], sm.Z[c][k + 1], s1);
        s2 = fma(sm.Z[a][k + 2], sm.Z[c][k + 2], s2);
        s3 = fma(sm.Z[a][k + 3], sm.Z[c][k + 3], s3);
      }
      A[(size_t)(KB_TILE + a) * ld + KB_TILE + c] = -((s0 + s1) + (s2 + s3));
    }
  }
  // ---- leave L (strict upper part zero) and Z where the general path leaves them ----
  for (int e = tid; e < KB_TILE * KB_TILE / 2; e += KB_THREADS) {
    const int r = e >> 5, c = (e & 31) * 2;
    double2 l = *reinterpret_cast<double2*>(&sm.C[r][c]);
    if (c > r) l.x = 0.0;
    if (c + 1 > r) l.y = 0.0;
    *reinterpret_cast<double2*>(&A[(size_t)r * ld + c]) = l;
  }
  for (int e = tid; e < KB_SMALL_Q * KB_TILE / 2; e += KB_THREADS) {
    const int q = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(&A[(size_t)(KB_TILE + q) * ld + c]) = *reinterpret_cast<double2*>(&sm.Z[q][c]);
  }
}

bool kb_small_applies(const KbAugGeom& g, int d, bool kpls) {
  return g.nb == 1 && g.sb == 1 && g.ib == 0 && g.q <= KB_SMALL_Q && !kpls && d <= KB_MAX_D;
}

cudaError_t kb_launch_nll_small(cudaStream_t st, const KbModelDev& m, const KbAugGeom& g, int B, int nbhyp,
                                const double* xpop_dev, int nth, int trainvar, int has_nugget, int has_weights,
                                double fixed_nugget, int nkernel, const int* kernel_ids_dev, KbHyper hp, double* aug,
                                int* info) {
  const size_t smem = sizeof(KbSmallSmem);
  cudaError_t e = cudaFuncSetAttribute(kb_nll_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kb_nll_small_kernel<<<B, KB_THREADS, smem, st>>>(m, g, nbhyp, xpop_dev, nth, trainvar, has_nugget, has_weights,
                                                   fixed_nugget, nkernel, kernel_ids_dev, hp, aug, info);
  return cudaGetLastError();
}
