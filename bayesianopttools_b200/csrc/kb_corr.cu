// krig-b200 K1: hyper-parameter decode + correlation-matrix build.
//
// Replaces calckernel/gaussian/exponential/matern32/matern52
// (surrogate_models/supports/kernelfunc.py:4-82) and the Psi assembly +
// nugget of likelihood()/maincalc (likelihood_func.py:93-102,171), batched over
// the hyper-parameter population.  Training sites are kept structure-of-arrays
// (XT[dim][point]) so one 64-point tile of one dimension is a contiguous 512-byte
// run: each tile is staged into shared memory with 1-D TMA bulk copies
// (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier, and every thread
// writes its 4x4 micro-tile with 16-byte stores that coalesce to 256-byte runs.
#include "kb_common.cuh"
#include "kb_internal.h"

// ---------------------------------------------------------------------------
// decode: log10 hyper-parameter rows -> per-dimension weights
// (likelihood_func.py:71-79, nuggetset :121-165, KPLS fold kernelfunc.py:43-49)
// ---------------------------------------------------------------------------
__global__ void kb_decode_kernel(int B, int nbhyp, const double* __restrict__ xpop, int d, int nth,
                                 int trainvar, int has_nugget, int has_weights, double fixed_nugget,
                                 int nkernel, const int* __restrict__ kernel_ids,
                                 const double* __restrict__ pls, KbHyper hp) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const double* x = xpop + (size_t)b * nbhyp;
  for (int dim = threadIdx.x; dim < d; dim += blockDim.x) {
    double gw, ith;
    if (pls == nullptr) {
      const double th = pow(10.0, x[dim]);
      ith = 1.0 / th;
      gw = 0.5 / (th * th);
    } else {
      double s = 0.0;
      for (int k = 0; k < nth; ++k) {
        const double th = pow(10.0, x[k]);
        const double w = pls[dim * nth + k];
        s += (w * w) / (th * th);
      }
      gw = 0.5 * s;
      ith = 0.0;
    }
    hp.gw[(size_t)b * d + dim] = gw;
    hp.ith[(size_t)b * d + dim] = ith;
  }
  if (threadIdx.x == 0) {
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
    for (int j = 0; j < 4; ++j) hp.wk[b * 4 + j] = wk[j];
    if (pls != nullptr)
      for (int k = 0; k < nth && k < KB_KPLS_MAX_H; ++k) {
        const double th = pow(10.0, x[k]);
        hp.gk[b * KB_KPLS_MAX_H + k] = 0.5 / (th * th);
      }
    hp.eps[b] = pow(10.0, nug);
    hp.nugget[b] = nug;
    hp.sigma2[b] = s2;
  }
}

cudaError_t kb_launch_decode(cudaStream_t st, int B, int nbhyp, const double* xpop_dev, int d, int nth,
                             int trainvar, int has_nugget, int has_weights, double fixed_nugget,
                             int nkernel, const int* kernel_ids_dev, const double* pls_dev, KbHyper hp) {
  kb_decode_kernel<<<B, 128, 0, st>>>(B, nbhyp, xpop_dev, d, nth, trainvar, has_nugget, has_weights,
                                      fixed_nugget, nkernel, kernel_ids_dev, pls_dev, hp);
  return cudaGetLastError();
}

// X (n x d row-major) -> XT (d x n_pad), zero padded
__global__ void kb_transpose_pad_kernel(const double* __restrict__ X, int n, int d,
                                        double* __restrict__ XT, int n_pad) {
  const long long total = (long long)d * n_pad;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int dim = (int)(e / n_pad), k = (int)(e % n_pad);
    XT[e] = (k < n) ? X[(size_t)k * d + dim] : 0.0;
  }
}
cudaError_t kb_launch_transpose_pad(cudaStream_t st, const double* X, int n, int d, double* XT, int n_pad) {
  const long long total = (long long)d * n_pad;
  int grid = (int)((total + 255) / 256);
  if (grid > 1184) grid = 1184;
  if (grid < 1) grid = 1;
  kb_transpose_pad_kernel<<<grid, 256, 0, st>>>(X, n, d, XT, n_pad);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// 64x64 correlation tile.  Thread (tr = tid/16, tc = tid%16) owns rows
// tr*4 + a (a<4) and columns 2*tc + 32*a2 + e (a2<2, e<2).
// ---------------------------------------------------------------------------
// PRE: single-kernel design whose staged coordinates were already multiplied by the dimension's weight (sqrt(gw) for the
// Gaussian, 1/theta otherwise; kb_pair_dim_scaled): two FP64 operations per pair and dimension instead of three.
template <int MASK, bool PRE = false>
__device__ __forceinline__ void kb_corr_tile(const double* __restrict__ sXi, const double* __restrict__ sXj,
                                             int d, const double* __restrict__ sgw,
                                             const double* __restrict__ sith, const double* __restrict__ swk,
                                             double (&out)[4][2][2]) {
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  KbPairAcc acc[4][2][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
      for (int e = 0; e < 2; ++e) kb_pair_init(acc[a][a2][e]);
  for (int dim = 0; dim < d; ++dim) {
    const double gw = sgw[dim], ith = sith[dim];
    double xi[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xi[a] = sXi[dim * KB_TILE + tr * 4 + a];
    double2 xj[2];
#pragma unroll
    for (int a2 = 0; a2 < 2; ++a2)
      xj[a2] = *reinterpret_cast<const double2*>(&sXj[dim * KB_TILE + 2 * tc + 32 * a2]);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) {
        if (PRE) {
          kb_pair_dim_scaled(acc[a][a2][0], xi[a] - xj[a2].x, MASK);
          kb_pair_dim_scaled(acc[a][a2][1], xi[a] - xj[a2].y, MASK);
        } else {
          kb_pair_dim(acc[a][a2][0], xi[a] - xj[a2].x, gw, ith, MASK);
          kb_pair_dim(acc[a][a2][1], xi[a] - xj[a2].y, gw, ith, MASK);
        }
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
      for (int e = 0; e < 2; ++e) out[a][a2][e] = kb_pair_finish(acc[a][a2][e], swk, MASK);
}

// Stage XT[:, ti*64 .. +64) and XT[:, tj*64 .. +64) with TMA bulk copies.
// sX layout: [2][d][64]; bar is a shared mbarrier (one use, parity 0).
__device__ __forceinline__ void kb_stage_x_tiles(const double* __restrict__ XTi, int ldi, int ti,
                                                 const double* __restrict__ XTj, int ldj, int tj, int d,
                                                 double* sX, uint64_t* bar) {
  if (threadIdx.x == 0) {
    kb_mbar_init(bar, 1);
    kb_mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    kb_mbar_expect_tx(bar, (uint32_t)(2 * d * KB_TILE * sizeof(double)));
    for (int dim = 0; dim < d; ++dim) {
      kb_tma_bulk_g2s(sX + dim * KB_TILE, XTi + (size_t)dim * ldi + (size_t)ti * KB_TILE,
                      KB_TILE * sizeof(double), bar);
      kb_tma_bulk_g2s(sX + (d + dim) * KB_TILE, XTj + (size_t)dim * ldj + (size_t)tj * KB_TILE,
                      KB_TILE * sizeof(double), bar);
    }
  }
  kb_mbar_wait(bar, 0);
}

// ---------------------------------------------------------------------------
// Population build of the augmented matrices (one 64x64 tile per CTA).
// ---------------------------------------------------------------------------
#ifndef KB_BUILD_PRESCALE
#define KB_BUILD_PRESCALE 1          // A/B build aid: 0 keeps the three-operation form in the population build
#endif
template <int MASK>
__global__ void __launch_bounds__(KB_THREADS)
kb_build_aug_kernel(KbModelDev m, KbAugGeom g, KbHyper hp, double* __restrict__ aug, int t_begin) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sX = reinterpret_cast<double*>(smem_raw);            // [2][d][64]
  double* sgw = sX + 2 * m.d * KB_TILE;                        // [d]
  double* sith = sgw + m.d;                                    // [d]
  double* swk = sith + m.d;                                    // [4]
  uint64_t* bar = reinterpret_cast<uint64_t*>(swk + 4);

  const int b = blockIdx.y;
  double* A = aug + (size_t)b * g.stride;
  const int nR = g.nb * (g.nb + 1) / 2;
  int t = blockIdx.x + t_begin;
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;

  if (t < nR) {
    // lower-triangular tile (ti, tj), tj <= ti
    int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    const int tj = t - ti * (ti + 1) / 2;
    // single-kernel designs: the staged coordinates are multiplied by the candidate's per-dimension weight once
    // (128 d products per tile), which takes one of the three FP64 operations out of the 4096 d pair-dimension steps
    constexpr bool PRE = KB_BUILD_PRESCALE && (MASK == 1 || MASK == 2 || MASK == 4 || MASK == 8);
    for (int i = threadIdx.x; i < m.d; i += blockDim.x) {
      const double gwv = hp.gw[(size_t)b * m.d + i], ithv = hp.ith[(size_t)b * m.d + i];
      sgw[i] = PRE ? ((MASK == 1) ? sqrt(gwv) : ithv) : gwv;
      sith[i] = ithv;
    }
    if (threadIdx.x < 4) swk[threadIdx.x] = hp.wk[b * 4 + threadIdx.x];
    kb_stage_x_tiles(m.XT, m.n_pad, ti, m.XT, m.n_pad, tj, m.d, sX, bar);
    __syncthreads();
    if (PRE) {
      for (int e = threadIdx.x; e < 2 * m.d * KB_TILE; e += blockDim.x) {
        int dim = e / KB_TILE;
        if (dim >= m.d) dim -= m.d;
        sX[e] *= sgw[dim];
      }
      __syncthreads();
    }
    double out[4][2][2];
    kb_corr_tile<MASK, PRE>(sX, sX + m.d * KB_TILE, m.d, sgw, sith, swk, out);
    const double eps = hp.eps[b];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = ti * KB_TILE + tr * 4 + a;
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) {
        const int col = tj * KB_TILE + 2 * tc + 32 * a2;
        double2 v;
        double* pv = &v.x;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = col + e;
          double val;
          if (row < m.n && c < m.n) val = out[a][a2][e] + ((row == c) ? eps : 0.0);
          else if (g.seated && row >= m.n && row < m.n + g.q) {
            // seated right-hand sides: row n is y, rows n+1.. the trend columns; zero Schur corner
            const int qr = row - m.n;
            val = (c < m.n) ? (qr == 0 ? m.y[c] : m.F[(size_t)c * m.nind + (qr - 1)]) : 0.0;
          }
          else val = (row == c) ? 1.0 : 0.0;       // identity padding
          pv[e] = val;
        }
        *reinterpret_cast<double2*>(&A[(size_t)row * g.ld + col]) = v;
      }
    }
  } else {
    // right-hand-side rows: y, F columns (as rows), identity rows (fit mode)
    t -= nR;
    const int ncolt = g.nb + g.sb;
    const int ti = g.nb + t / ncolt, tj = t % ncolt;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = tr * 4 + a;
      const int row = ti * KB_TILE + r;
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) {
        const int col = tj * KB_TILE + 2 * tc + 32 * a2;
        double2 v;
        double* pv = &v.x;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = col + e;
          double val = 0.0;
          if (c < m.n) {
            if (ti < g.nb + g.sb) {
              const int q = (ti - g.nb) * KB_TILE + r;
              if (q == 0) val = m.y[c];
              else if (q <= m.nind) val = m.F[(size_t)c * m.nind + (q - 1)];
            } else {
              const int qi = (ti - g.nb - g.sb) * KB_TILE + r;
              if (qi == c) val = 1.0;
            }
          }
          pv[e] = val;
        }
        *reinterpret_cast<double2*>(&A[(size_t)row * g.ld + col]) = v;
      }
    }
  }
}

static size_t kb_corr_smem(int d) {
  return (size_t)(2 * d * KB_TILE + 2 * d + 4) * sizeof(double) + 16;
}

#define KB_DISPATCH_MASK(MASKVAR, CALL)                                  \
  switch (MASKVAR) {                                                     \
    case 1: { constexpr int M_ = 1; CALL; } break;                       \
    case 2: { constexpr int M_ = 2; CALL; } break;                       \
    case 3: { constexpr int M_ = 3; CALL; } break;                       \
    case 4: { constexpr int M_ = 4; CALL; } break;                       \
    case 5: { constexpr int M_ = 5; CALL; } break;                       \
    case 6: { constexpr int M_ = 6; CALL; } break;                       \
    case 7: { constexpr int M_ = 7; CALL; } break;                       \
    case 8: { constexpr int M_ = 8; CALL; } break;                       \
    case 9: { constexpr int M_ = 9; CALL; } break;                       \
    case 10: { constexpr int M_ = 10; CALL; } break;                     \
    case 11: { constexpr int M_ = 11; CALL; } break;                     \
    case 12: { constexpr int M_ = 12; CALL; } break;                     \
    case 13: { constexpr int M_ = 13; CALL; } break;                     \
    case 14: { constexpr int M_ = 14; CALL; } break;                     \
    default: { constexpr int M_ = 15; CALL; } break;                     \
  }

template <int MASK>
static cudaError_t kb_build_aug_launch(cudaStream_t st, const KbModelDev& m, const KbAugGeom& g, int B,
                                       KbHyper hp, double* aug, int skip_r_tiles) {
  const size_t smem = kb_corr_smem(m.d);
  cudaError_t e = cudaFuncSetAttribute(kb_build_aug_kernel<MASK>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int nR = g.nb * (g.nb + 1) / 2;
  const int ntiles = nR + (g.sb + g.ib) * (g.nb + g.sb);
  const int t_begin = skip_r_tiles ? nR : 0;
  if (ntiles - t_begin <= 0) return cudaSuccess;     // seated right-hand sides: nothing left to build
  dim3 grid(ntiles - t_begin, B);
  kb_build_aug_kernel<MASK><<<grid, KB_THREADS, smem, st>>>(m, g, hp, aug, t_begin);
  return cudaGetLastError();
}

cudaError_t kb_launch_build_aug(cudaStream_t st, const KbModelDev& m, const KbAugGeom& g, int B,
                                KbHyper hp, double* aug, int skip_r_tiles) {
  cudaError_t err = cudaSuccess;
  KB_DISPATCH_MASK(m.kmask, err = kb_build_aug_launch<M_>(st, m, g, B, hp, aug, skip_r_tiles));
  return err;
}

// ---------------------------------------------------------------------------
// KPLS fast path.  kernelfunc.py:43-49 evaluates exp(-1/2 sum_k [sum_i w_ik^2 (a_i-b_i)^2] / theta_k^2):
// the inner sums do not depend on theta, so they are computed ONCE per design (h values per pair,
// tile-major: dtiles[tile][k][64][64]) and a candidate's correlation tile costs h FMAs and one exp
// per entry instead of 3 d flops.  Thread (tr, tc) owns the same 4x4 entries as kb_corr_tile.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void kb_tile_of(int t, int& ti, int& tj) {
  ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  tj = t - ti * (ti + 1) / 2;
}

__global__ void __launch_bounds__(KB_THREADS)
kb_kpls_dist_kernel(KbModelDev m, int h, const double* __restrict__ pls, double* __restrict__ dtiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sX = reinterpret_cast<double*>(smem_raw);            // [2][d][64]
  double* sw2 = sX + 2 * m.d * KB_TILE;                        // [d][h] squared PLS weights
  uint64_t* bar = reinterpret_cast<uint64_t*>(sw2 + m.d * KB_KPLS_MAX_H);
  int ti, tj;
  kb_tile_of(blockIdx.x, ti, tj);
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  for (int e = threadIdx.x; e < m.d * h; e += blockDim.x) {
    const double w = pls[e];
    sw2[(e / h) * KB_KPLS_MAX_H + (e % h)] = w * w;
  }
  kb_stage_x_tiles(m.XT, m.n_pad, ti, m.XT, m.n_pad, tj, m.d, sX, bar);
  __syncthreads();
  const double* sXi = sX;
  const double* sXj = sX + m.d * KB_TILE;
  for (int k = 0; k < h; ++k) {
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
    for (int dim = 0; dim < m.d; ++dim) {
      const double w2 = sw2[dim * KB_KPLS_MAX_H + k];
      double xi[4], xj[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xi[a] = sXi[dim * KB_TILE + tr * 4 + a];
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) {
        const double2 v = *reinterpret_cast<const double2*>(&sXj[dim * KB_TILE + 2 * tc + 32 * a2]);
        xj[2 * a2] = v.x; xj[2 * a2 + 1] = v.y;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double df = xi[a] - xj[c];
          acc[a][c] = fma(df * df, w2, acc[a][c]);
        }
    }
    double* out = dtiles + ((size_t)blockIdx.x * h + k) * (KB_TILE * KB_TILE);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2)
        *reinterpret_cast<double2*>(&out[(tr * 4 + a) * KB_TILE + 2 * tc + 32 * a2]) =
            make_double2(acc[a][2 * a2], acc[a][2 * a2 + 1]);
  }
}

cudaError_t kb_launch_kpls_dist(cudaStream_t st, const KbModelDev& m, int h, const double* pls, double* dtiles) {
  const size_t smem = (size_t)(2 * m.d * KB_TILE + m.d * KB_KPLS_MAX_H) * sizeof(double) + 16;
  cudaError_t e = cudaFuncSetAttribute(kb_kpls_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kb_kpls_dist_kernel<<<m.nb * (m.nb + 1) / 2, KB_THREADS, smem, st>>>(m, h, pls, dtiles);
  return cudaGetLastError();
}

#define KB_KPLS_GROUP 8    // candidates per CTA: the D tiles are read once per group
__global__ void __launch_bounds__(KB_THREADS)
kb_build_aug_kpls_kernel(KbModelDev m, KbAugGeom g, int B, int h, KbHyper hp, const double* __restrict__ dtiles,
                         double* __restrict__ aug) {
  __shared__ double sgk[KB_KPLS_GROUP][KB_KPLS_MAX_H];
  __shared__ double seps[KB_KPLS_GROUP];
  int ti, tj;
  kb_tile_of(blockIdx.x, ti, tj);
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  const int b0 = blockIdx.y * KB_KPLS_GROUP;
  const int nb_here = (B - b0 < KB_KPLS_GROUP) ? (B - b0) : KB_KPLS_GROUP;
  if (threadIdx.x < KB_KPLS_GROUP * KB_KPLS_MAX_H) {
    const int bb = threadIdx.x / KB_KPLS_MAX_H, k = threadIdx.x % KB_KPLS_MAX_H;
    sgk[bb][k] = (bb < nb_here && k < h) ? hp.gk[(b0 + bb) * KB_KPLS_MAX_H + k] : 0.0;
    if (k == 0) seps[bb] = (bb < nb_here) ? hp.eps[b0 + bb] : 0.0;
  }
  __syncthreads();
  const bool inside = (ti + 1) * KB_TILE <= m.n;            // no padding row/column in this tile
  const double* dt = dtiles + (size_t)blockIdx.x * h * (KB_TILE * KB_TILE);
#pragma unroll 1
  for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll 1
    for (int a = 0; a < 4; ++a) {
      const int r = tr * 4 + a, c = 2 * tc + 32 * a2;
      double2 dv[KB_KPLS_MAX_H];
#pragma unroll
      for (int k = 0; k < KB_KPLS_MAX_H; ++k)
        dv[k] = (k < h) ? __ldg(reinterpret_cast<const double2*>(&dt[(size_t)k * (KB_TILE * KB_TILE) + r * KB_TILE + c]))
                        : make_double2(0.0, 0.0);
      const int row = ti * KB_TILE + r, col = tj * KB_TILE + c;
      for (int bb = 0; bb < nb_here; ++bb) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < KB_KPLS_MAX_H; ++k) {
          s0 = fma(dv[k].x, sgk[bb][k], s0);
          s1 = fma(dv[k].y, sgk[bb][k], s1);
        }
        double2 v = make_double2(kb_exp_nonpos(-s0), kb_exp_nonpos(-s1));
        if (row == col) v.x += seps[bb];
        if (row == col + 1) v.y += seps[bb];
        if (!inside) {
          // padding: identity, except the seated right-hand-side rows n .. n+q-1 (y, trend columns; zero corner)
          const bool rhs_row = g.seated && row >= m.n && row < m.n + g.q;
          const int qr = row - m.n;
          if (row >= m.n || col >= m.n) {
            if (rhs_row) v.x = (col < m.n) ? (qr == 0 ? m.y[col] : m.F[(size_t)col * m.nind + (qr - 1)]) : 0.0;
            else v.x = (row == col) ? 1.0 : 0.0;
          }
          if (row >= m.n || col + 1 >= m.n) {
            if (rhs_row) v.y = (col + 1 < m.n) ? (qr == 0 ? m.y[col + 1] : m.F[(size_t)(col + 1) * m.nind + (qr - 1)]) : 0.0;
            else v.y = (row == col + 1) ? 1.0 : 0.0;
          }
        }
        *reinterpret_cast<double2*>(&aug[(size_t)(b0 + bb) * g.stride + (size_t)row * g.ld + col]) = v;
      }
    }
}

cudaError_t kb_launch_build_aug_kpls(cudaStream_t st, const KbModelDev& m, const KbAugGeom& g, int B, int h,
                                     KbHyper hp, const double* dtiles, double* aug) {
  dim3 grid(g.nb * (g.nb + 1) / 2, (B + KB_KPLS_GROUP - 1) / KB_KPLS_GROUP);
  kb_build_aug_kpls_kernel<<<grid, KB_THREADS, 0, st>>>(m, g, B, h, hp, dtiles, aug);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return kb_launch_build_aug(st, m, g, B, hp, aug, 1);       // right-hand-side / identity rows
}

// ---------------------------------------------------------------------------
// General (n x m) correlation matrix: calckernel() itself (kernelfunc.py:4-34).
// XNT is [d][ldn], XMT is [d][ldm], both zero padded to a multiple of 64.
// ---------------------------------------------------------------------------
template <int MASK>
__global__ void __launch_bounds__(KB_THREADS)
kb_corr_general_kernel(int n, int mcols, int d, const double* __restrict__ XNT, int ldn,
                       const double* __restrict__ XMT, int ldm, const double* __restrict__ gw,
                       const double* __restrict__ ith, const double* __restrict__ wk,
                       double* __restrict__ out, int ldo) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sX = reinterpret_cast<double*>(smem_raw);
  double* sgw = sX + 2 * d * KB_TILE;
  double* sith = sgw + d;
  double* swk = sith + d;
  uint64_t* bar = reinterpret_cast<uint64_t*>(swk + 4);
  const int ti = blockIdx.y, tj = blockIdx.x;
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    sgw[i] = gw[i];
    sith[i] = ith[i];
  }
  if (threadIdx.x < 4) swk[threadIdx.x] = wk[threadIdx.x];
  kb_stage_x_tiles(XNT, ldn, ti, XMT, ldm, tj, d, sX, bar);
  __syncthreads();
  double o[4][2][2];
  kb_corr_tile<MASK>(sX, sX + d * KB_TILE, d, sgw, sith, swk, o);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = ti * KB_TILE + tr * 4 + a;
    if (row >= n) continue;
#pragma unroll
    for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = tj * KB_TILE + 2 * tc + 32 * a2 + e;
        if (c < mcols) out[(size_t)row * ldo + c] = o[a][a2][e];
      }
  }
}

template <int MASK>
static cudaError_t kb_corr_general_launch(cudaStream_t st, int n, int m, int d, const double* XNT, int ldn,
                                          const double* XMT, int ldm, const double* gw, const double* ith,
                                          const double* wk, double* out, int ldo) {
  const size_t smem = kb_corr_smem(d);
  cudaError_t e = cudaFuncSetAttribute(kb_corr_general_kernel<MASK>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((m + KB_TILE - 1) / KB_TILE, (n + KB_TILE - 1) / KB_TILE);
  kb_corr_general_kernel<MASK><<<grid, KB_THREADS, smem, st>>>(n, m, d, XNT, ldn, XMT, ldm, gw, ith, wk,
                                                               out, ldo);
  return cudaGetLastError();
}

cudaError_t kb_launch_corr_general(cudaStream_t st, int n, int m, int d, const double* XNT, int ldn,
                                   const double* XMT, int ldm, const double* gw, const double* ith,
                                   const double* wk, int kmask, double* out, int ldo) {
  cudaError_t err = cudaSuccess;
  KB_DISPATCH_MASK(kmask, err = kb_corr_general_launch<M_>(st, n, m, d, XNT, ldn, XMT, ldm, gw, ith, wk,
                                                            out, ldo));
  return err;
}
