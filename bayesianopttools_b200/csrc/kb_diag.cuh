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
#define KB_DIAG_SCRATCH 96   // doubles of exchange scratch: 2 x (column | inverse row | diagonal)
#ifndef KB_DIAG_STAMP
#define KB_DIAG_STAMP(i)
#endif
// 16x16 diagonal blocks whose pivots span more than this ratio (as max L_jj / min L_jj) get the residual-corrected
// panel product in phase 2 (same criterion as KB_REFINE_RATIO of the tile Cholesky, kb_chol.cu).
#define KB_DIAG_REFINE_RATIO 100.0

// 1/sqrt(p) to full FP64 accuracy with a 4-deep dependent chain: hardware approximation
// (relative error ~2^-23) + one third-order step  y (1 + e/2 + 3e^2/8),  e = 1 - p y^2.
__device__ __forceinline__ double kb_rsqrt_seed(double p) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  return y;
}
__device__ __forceinline__ double kb_rsqrt_refine(double p, double y) {
  const double py = p * y;
  const double e = fma(-py, y, 1.0);
  const double ye = y * e;
  const double q = fma(e, 0.375, 0.5);
  return fma(ye, q, y);
}
__device__ __forceinline__ double kb_rsqrt(double p) { return kb_rsqrt_refine(p, kb_rsqrt_seed(p)); }
// 1/p: hardware approximation (~2^-20) + one third-order step  y (1 + e + e^2),  e = 1 - p y.
__device__ __forceinline__ double kb_rcp_seed(double p) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  return y;
}
__device__ __forceinline__ double kb_rcp_refine(double p, double y) {
  const double e = fma(-p, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}


// Work items of phase 3 per (panel, warp) as byte offsets of the (lane 0) operands:
//   x = A operand in Cs, y = B operand, z = output, w = flags (-1 none; bit 0: B and the
//   output live in Ds (running sums S) instead of Cs (trailing tile); bit 1: accumulate onto
//   the stored value (otherwise the sum starts at zero: first touch of a new S column)).
__constant__ int4 kb_diag_items[3][8][5] = {
  {{{8704, 8704, 8832, 2}, {21760, 17408, 22016, 2}, {30464, 13056, 30656, 2}, {13056, 64, 13120, 1}, {30464, 64, 30528, 1}},
   {{13056, 8704, 13184, 2}, {21760, 21760, 22080, 2}, {30464, 17408, 30720, 2}, {17408, 0, 17408, 1}, {0, 0, 0, -1}},
   {{13056, 13056, 13248, 2}, {26112, 8704, 26240, 2}, {30464, 21760, 30784, 2}, {17408, 64, 17472, 1}, {0, 0, 0, -1}},
   {{17408, 8704, 17536, 2}, {26112, 13056, 26304, 2}, {30464, 26112, 30848, 2}, {21760, 0, 21760, 1}, {0, 0, 0, -1}},
   {{17408, 13056, 17600, 2}, {26112, 17408, 26368, 2}, {30464, 30464, 30912, 2}, {21760, 64, 21824, 1}, {0, 0, 0, -1}},
   {{17408, 17408, 17664, 2}, {26112, 21760, 26432, 2}, {8704, 0, 8704, 1}, {26112, 0, 26112, 1}, {0, 0, 0, -1}},
   {{21760, 8704, 21888, 2}, {26112, 26112, 26496, 2}, {8704, 64, 8768, 1}, {26112, 64, 26176, 1}, {0, 0, 0, -1}},
   {{21760, 13056, 21952, 2}, {30464, 8704, 30592, 2}, {13056, 0, 13056, 1}, {30464, 0, 30464, 1}, {0, 0, 0, -1}}},
  {{{17536, 17536, 17664, 2}, {30592, 26240, 30848, 2}, {21888, 8832, 21888, 1}, {30592, 8832, 30592, 1}, {0, 0, 0, -1}},
   {{21888, 17536, 22016, 2}, {30592, 30592, 30912, 2}, {21888, 8896, 21952, 1}, {30592, 8896, 30656, 1}, {0, 0, 0, -1}},
   {{21888, 21888, 22080, 2}, {17536, 8704, 17408, 3}, {26240, 8704, 26112, 3}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{26240, 17536, 26368, 2}, {17536, 8768, 17472, 3}, {26240, 8768, 26176, 3}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{26240, 21888, 26432, 2}, {17536, 8832, 17536, 1}, {26240, 8832, 26240, 1}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{26240, 26240, 26496, 2}, {17536, 8896, 17600, 1}, {26240, 8896, 26304, 1}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{30592, 17536, 30720, 2}, {21888, 8704, 21760, 3}, {30592, 8704, 30464, 3}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{30592, 21888, 30784, 2}, {21888, 8768, 21824, 3}, {30592, 8768, 30528, 3}, {0, 0, 0, -1}, {0, 0, 0, -1}}},
  {{{26368, 26368, 26496, 2}, {26368, 17728, 26432, 1}, {0, 0, 0, -1}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{30720, 26368, 30848, 2}, {30720, 17408, 30464, 3}, {0, 0, 0, -1}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{30720, 30720, 30912, 2}, {30720, 17472, 30528, 3}, {0, 0, 0, -1}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{26368, 17408, 26112, 3}, {30720, 17536, 30592, 3}, {0, 0, 0, -1}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{26368, 17472, 26176, 3}, {30720, 17600, 30656, 3}, {0, 0, 0, -1}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{26368, 17536, 26240, 3}, {30720, 17664, 30720, 1}, {0, 0, 0, -1}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{26368, 17600, 26304, 3}, {30720, 17728, 30784, 1}, {0, 0, 0, -1}, {0, 0, 0, -1}, {0, 0, 0, -1}},
   {{26368, 17664, 26368, 1}, {0, 0, 0, -1}, {0, 0, 0, -1}, {0, 0, 0, -1}, {0, 0, 0, -1}}},
};

// 64x64 SPD tile in Cs (lower part used) -> Cs = L (lower triangle valid), Ds = L^-1 (lower 8x8
// blocks valid, the 16x16 diagonal blocks carry explicit zeros above the diagonal; blocks above
// are NOT written - every consumer stops at the diagonal block).
// Right-looking with 16-wide panels; every phase is shared by all 8 warps:
//   1. the 16x16 diagonal block is factored AND inverted with one element of each per thread;
//      column j and row j of the inverse travel through a double-buffered shared exchange with
//      ONE barrier per column.  The NEXT pivot is formed by every thread from the exchanged
//      column/diagonal (look-ahead) and its 1/sqrt is software pipelined across the barrier.
//      The step is issue bound, so it is written for instruction count: unconditional rank-1
//      updates (finished elements are latched with predicated moves, stale ones become
//      don't-care), one DSETP per pivot test.  A non-positive pivot poisons that candidate's
//      tile with NaN (nothing else) and is reported once per panel.
//   2. panel  P = A W_D^T  and the finished block row of the inverse  W_i,: = W_D S_i,:   (DMMA)
//   3. trailing  C -= P P^T  and the running sums  S -= P W_i,:   (DMMA; operand offsets from a
//      constant table, up to 5 independent 8x8 products in flight per warp)
// where S_ij = -sum_k L_ik W_kj accumulates in the not-yet-final part of Ds.
__device__ __forceinline__ void kb_diag_factor_inverse(double (*Cs)[KB_LDC], double (*Ds)[KB_LDC],
                                                       double* xch, int* info, int pivot_base) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const uint32_t cs_u = kb_smem_u32(&Cs[0][0]), ds_u = kb_smem_u32(&Ds[0][0]);
  const uint32_t laneA = (uint32_t)(gq * KB_LDC + tq) * 8u, laneBs = (uint32_t)(tq * KB_LDC + gq) * 8u,
                 laneO = (uint32_t)(gq * KB_LDC + 2 * tq) * 8u;
  KB_DIAG_STAMP(0);
  KB_DIAG_STAMP(1);
#pragma unroll 1
  for (int jb = 0; jb < 4; ++jb) {
    const int c0 = jb * 16;
    {
      const int r = tid >> 4, c = tid & 15;
      double a = Cs[c0 + r][c0 + c];
      double w = (r == c) ? 1.0 : 0.0;
      double piv = Cs[c0][c0];
      bool ok = piv > 0.0;           // not positive definite (or NaN): reported after the panel
      double seed = kb_rcp_seed(piv);
      double2* xa = reinterpret_cast<double2*>(xch);       // [2][16] (column j of A, diagonal of A)
      double* xw = xch + 64;                               // [2][16] row j of W
      // The loop works on UNSCALED columns: with q_j = 1 / pivot_j,
      //   A[r][c] -= A[r][j] A[c][j] q_j ,  W[r][c] -= A[r][j] W[j][c] q_j ,
      // each element freezes when its column (A) / row (W) becomes the pivot, and the
      // 1/sqrt(pivot) scaling is applied once after the loop.  Per column and thread: 9 FP64
      // instructions (the step is FP64-issue bound: 256 threads share 64 lanes/clk).
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double2* ba = xa + (j & 1) * 16;
        double* bw = xw + (j & 1) * 16;
        if (c == j) ba[r].x = a;
        if (r == c) ba[r].y = a;
        if (r == j) bw[c] = w;
        KB_DIAG_SYNC();
        const double colr = ba[r].x, colc = ba[c].x, wj = bw[c];
        const double2 nx = ba[(j + 1) & 15];             // (A[j+1][j], A[j+1][j+1])
        const double q = kb_rcp_refine(piv, seed);       // 1 / pivot j
        const double t = colr * q;
        if (c > j) a = fma(-t, colc, a);
        if (r > j) w = fma(-t, wj, w);
        if (j < 15) {                // look-ahead: pivot j+1 after the rank-1 update of column j
          const double t2 = nx.x * q;
          piv = fma(-t2, nx.x, nx.y);
          ok = ok && (piv > 0.0);
          seed = kb_rcp_seed(piv);
        }
      }
      // pivots -> everyone; scale: L[r][c] = A_frozen[r][c] / sqrt(p_c), W[r][c] = W_frozen[r][c] / sqrt(p_r)
      double* xp = xch + 48;         // free part of exchange buffer 0.. (last step used buffer 1)
      KB_DIAG_SYNC();
      if (r == c) xch[r] = a;
      KB_DIAG_SYNC();
      const double ic = kb_rsqrt(xch[c]), ir = kb_rsqrt(xch[r]);
      (void)xp;
      if (c <= r) Cs[c0 + r][c0 + c] = a * ic;
      Ds[c0 + r][c0 + c] = (c <= r) ? w * ir : 0.0;
      if (!ok) {                     // rare: first pivot of this panel that is not a positive number
        if (tid == 0) {
          int bad = 0;
          for (int j = 15; j >= 0; --j) {
            const double pj = xch[j];
            if (!(pj > 0.0) || !(pj < 1.7e308)) bad = j;
          }
          atomicCAS(info, 0, pivot_base + c0 + bad + 1);
        }
      }
    }
    KB_DIAG_SYNC();
    KB_DIAG_STAMP(2 + jb * 3);
    const int nrb = (KB_TILE - 16 - c0) / 8;       // 8-row blocks below the diagonal block
    // ---- phase 2: panel rows (warps 0..nrb-1) and the finished inverse block row (next 2*jb warps)
    if (warp < nrb) {
      const int r0 = c0 + 16 + 8 * warp;
      // pivot spread of the 16x16 block just factored (xch[0..15] = its unscaled pivots L_jj^2)
#ifdef KB_DIAG_NO_REFINE                // A/B build aid
      const bool refine = false;
#else
      double plo = xch[lane & 15], phi = plo;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        plo = fmin(plo, __shfl_xor_sync(0xffffffffu, plo, o));
        phi = fmax(phi, __shfl_xor_sync(0xffffffffu, phi, o));
      }
      const bool refine = phi > (KB_DIAG_REFINE_RATIO * KB_DIAG_REFINE_RATIO) * plo;
#endif
      double af[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) af[ks] = Cs[r0 + gq][c0 + ks * 4 + tq];
      double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) kb_dmma(p[ni][0], p[ni][1], af[ks], Ds[c0 + ni * 8 + gq][c0 + ks * 4 + tq]);
      if (refine) {
        // The product with the explicit 16x16 inverse leaves a residual ~ eps cond(L_D) |A|: when the pivots of
        // the block span orders of magnitude (smooth kernel, tiny nugget, first panel of the first tile) take
        // one step of residual correction   P <- P + (A - P L_D^T) W_D^T   -- all inside the warp that owns the
        // 8 rows.  Well-conditioned blocks never come here.
        double rs[2][2];
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const double2 a = *reinterpret_cast<const double2*>(&Cs[r0 + gq][c0 + ni * 8 + 2 * tq]);
          rs[ni][0] = a.x;
          rs[ni][1] = a.y;
        }
        __syncwarp();
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
          *reinterpret_cast<double2*>(&Cs[r0 + gq][c0 + ni * 8 + 2 * tq]) = make_double2(p[ni][0], p[ni][1]);
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double x = Cs[r0 + gq][c0 + ks * 4 + tq];
#pragma unroll
          for (int ni = 0; ni < 2; ++ni) {
            const int kk = ks * 4 + tq, nn = ni * 8 + gq;             // B[k][n] = L_D[n][k], lower triangular
            const double l = (kk <= nn) ? Cs[c0 + nn][c0 + kk] : 0.0;
            kb_dmma(rs[ni][0], rs[ni][1], -x, l);
          }
        }
        __syncwarp();
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
          *reinterpret_cast<double2*>(&Cs[r0 + gq][c0 + ni * 8 + 2 * tq]) = make_double2(rs[ni][0], rs[ni][1]);
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double r = Cs[r0 + gq][c0 + ks * 4 + tq];
#pragma unroll
          for (int ni = 0; ni < 2; ++ni) kb_dmma(p[ni][0], p[ni][1], r, Ds[c0 + ni * 8 + gq][c0 + ks * 4 + tq]);
        }
      }
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
    {
      constexpr int NB = 5;
      uint32_t pa[NB], pb[NB], po[NB], kss[NB];
      bool on[NB];
      double2 acc[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int4 it = kb_diag_items[jb][warp][i];
        on[i] = it.w >= 0;
        const bool inD = (it.w & 1) != 0;
        pa[i] = cs_u + (uint32_t)it.x + laneA;
        pb[i] = (inD ? ds_u + laneBs : cs_u + laneA) + (uint32_t)it.y;
        po[i] = (inD ? ds_u : cs_u) + (uint32_t)it.z + laneO;
        kss[i] = inD ? 4u * KB_LDC * 8u : 32u;
        acc[i] = make_double2(0.0, 0.0);
        if (on[i] && (it.w & 2))
          asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(acc[i].x), "=d"(acc[i].y) : "r"(po[i]));
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int i = 0; i < NB; ++i)
          if (on[i]) {
            double av, bv;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(av) : "r"(pa[i] + ks * 32u));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(bv) : "r"(pb[i] + ks * kss[i]));
            kb_dmma(acc[i].x, acc[i].y, -av, bv);
          }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NB; ++i)
        if (on[i]) asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(po[i]), "d"(acc[i].x), "d"(acc[i].y) : "memory");
    }
    KB_DIAG_SYNC();
    KB_DIAG_STAMP(4 + jb * 3);
  }
  KB_DIAG_SYNC();
  KB_DIAG_STAMP(12);
}
