// krig-b200 K4: fused batched predictor / acquisition sweep.
//
// Replaces prediction() (surrogate_models/supports/prediction.py:155-289) and the
// AK-MCS population reductions (reliability_analysis/akmcs.py:208-238) for
// millions of sites per call.  One persistent CTA handles strips of 64 sites:
//   A. standardise the sites (samplingplan.py:84-88 / prediction.py:160) and build
//      the cross-correlation strip psi (n x 64)            (prediction.py:187-201)
//   B. one DMMA GEMM against the k-major fit state  AT = [L^-T | alpha | R^-1 F]:
//        rows <  n_pad : v = L^-1 psi, only sum v^2 is kept  (prediction.py:222,224)
//        row  n_pad    : psi' R^-1 (y - F BE)                (prediction.py:209-212)
//        rows n_pad+1..: F' R^-1 psi                         (prediction.py:225)
//      (lower-triangular L^-1: row block I only needs k < 128 (I+1))
//   C. per site: mean, SSqr, s, EI / PoI / PoF / LCB / EBE / U  (prediction.py:227-285)
//      and the running arg-min / failure counters (akmcs.py:208-238).
// psi never leaves L2: the strip lives in a CTA-private scratch slab that is
// written once and streamed back through a 4-stage cp.async ring.
#include <float.h>

#include "kb_common.cuh"
#include "kb_internal.h"

#define PK_BM 128
#define PK_KC 16
#define PK_STAGES 3
#define PK_LDA 132
#define PK_TK 16          // rows of the training set per psi-phase tile
#define PK_QIN 8          // extra rows (alpha, R^-1 F columns) folded into the psi phase up to this many

// Shared memory: [header][xs d x 64][gw d][1/theta d][big]
//   streaming kernel: big = A ring [stages][16][132] | psi ring [stages][16][68]
//   resident kernel : big = L^-1 (k-major) [n_pad][n_pad+4] | psi strip [n_pad][68] | extra-row partials
struct KbPredSmem {
  double vv[4][128];
  KbSweepResult red[4];
  double wk[4];
};

int kb_predict_strip(bool resident) { (void)resident; return 64; }
int kb_predict_threads(bool resident) { (void)resident; return 256; }

// stage_x: raw sites of the next strip are prefetched into two shared buffers (2 x 64 x d doubles).
// For wide inputs that staging costs a resident CTA (or does not fit at all), so the host turns it off
// and the strip is read straight from global memory when it is standardised.
size_t kb_predict_smem_bytes(int d, bool resident, int n_pad, int qx, int stage_x) {
  const size_t mt = kb_predict_strip(resident);
  size_t big;
  if (resident)
    big = sizeof(double) * ((size_t)n_pad * (n_pad + 4) + (size_t)n_pad * (mt + 4) + 4 * (size_t)(qx < PK_QIN ? qx : PK_QIN) * mt);
  else
    big = sizeof(double) * PK_STAGES * PK_KC * (PK_LDA + mt + 4);
  return sizeof(KbPredSmem) +
         ((size_t)d * (mt + 2) + 2 * (size_t)d * PK_TK + 2 * PK_TK * PK_QIN + (stage_x ? 2 * mt * (size_t)d : 0)) * sizeof(double) + big;
}

// The resident variant needs the whole inverse factor in shared memory.
bool kb_predict_can_reside(int d, int n_pad, int qx, int nind_trend) {
  return n_pad <= 128 && nind_trend == 0 && qx <= PK_QIN &&
         kb_predict_smem_bytes(d, true, n_pad, qx, 1) <= 220 * 1024;
}

__device__ __forceinline__ double kb_legendre_on(int o, double x) {
  // orthonormal Legendre on [-1,1]: P_o(x) * sqrt(2 o + 1)   (trendfunction.py:98-101)
  double pm = 1.0, p = x;
  if (o == 0) return 1.0;
  for (int k = 1; k < o; ++k) {
    const double pn = ((2 * k + 1) * x * p - k * pm) / (k + 1);
    pm = p;
    p = pn;
  }
  return p * sqrt(2.0 * o + 1.0);
}

struct KbRunning {
  double best_val, min_u;
  long long best_idx, min_u_idx;
  long long c0, c1, c2;
};

// Per-site outputs from the four sufficient statistics (prediction.py:209-285, akmcs.py:208-238):
//   fpc = p(x)' BE,  psia = psi' R^-1 (y - F BE),  vv = psi' R^-1 psi,  t2 = u' (F'R^-1F)^-1 u.
// erf / exp are evaluated only when an output or the arg-min needs them.
__device__ __forceinline__ void kb_acq_site(const KbPredictArgs& P, double fpc, double psia, double vv, double t2,
                                            long long gi, KbRunning& run) {
  double f = fpc + psia;
  if (P.ymode == KB_YMAP_DEFAULT) f = (f * 0.5 + 0.5) * P.ya + P.yb;      // stdtoreal, prediction.py:300-313
  else if (P.ymode == KB_YMAP_STD) f = P.yb + P.ya * f;
  const double ssqr = P.sigma2 * ((1.0 - vv) + t2);
  const double s = sqrt(fabs(ssqr));
  const bool want_ei = P.out[KB_OUT_EI] || P.out[KB_OUT_POI] || P.acq == KB_OUT_EI || P.acq == KB_OUT_POI;
  const bool want_pof = P.out[KB_OUT_POF] || P.acq == KB_OUT_POF;
  double ei = 0.0, poi = 0.0, pof = 0.0;
  if (want_ei) {
    const double diff = P.ybest - f;
    const double z = diff / s;
    const double cdf = 0.5 + 0.5 * erf(0.7071067811865475 * z);
    ei = -(diff * cdf + s / 2.5066282746310002 * exp(-0.5 * (diff * diff) / ssqr) + DBL_MIN);
    poi = -cdf;
  }
  if (want_pof) pof = 0.5 + 0.5 * erf(0.7071067811865475 * ((double)P.limit_sign * (f - P.limit) / s));
  const double lcb = f - P.sigmalcb * ssqr;
  const double U = fabs(f) / s;
  double vals[KB_OUT_COUNT];
  vals[KB_OUT_PRED] = f; vals[KB_OUT_SSQR] = ssqr; vals[KB_OUT_S] = s; vals[KB_OUT_EI] = ei;
  vals[KB_OUT_POI] = poi; vals[KB_OUT_POF] = pof; vals[KB_OUT_LCB] = lcb; vals[KB_OUT_EBE] = -ssqr;
  vals[KB_OUT_U] = U; vals[KB_OUT_FPC] = fpc;
#pragma unroll
  for (int k = 0; k < KB_OUT_COUNT; ++k)
    if (P.out[k]) P.out[k][gi] = vals[k];
  double av = vals[KB_OUT_PRED];
#pragma unroll
  for (int k = 0; k < KB_OUT_COUNT; ++k)
    if (P.acq == k) av = vals[k];
  const long long ggi = gi + P.idx_offset;
  if (av < run.best_val) { run.best_val = av; run.best_idx = ggi; }
  if (U < run.min_u) { run.min_u = U; run.min_u_idx = ggi; }
  run.c0 += (f <= 0.0);
  run.c1 += (f - 1.96 * s <= 0.0);
  run.c2 += (f + 1.96 * s <= 0.0);
}

__device__ __forceinline__ void kb_run_init(KbRunning& run) {
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  run.best_val = INF; run.min_u = INF;
  run.best_idx = 0x7fffffffffffffffLL; run.min_u_idx = 0x7fffffffffffffffLL;
  run.c0 = run.c1 = run.c2 = 0;
}

// warp-level reduction of the running (min, lowest index) pairs and counters; lane 0 holds the result
__device__ __forceinline__ KbSweepResult kb_run_warp_reduce(KbRunning run) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, run.best_val, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, run.best_idx, o);
    if (ov < run.best_val || (ov == run.best_val && oi < run.best_idx)) { run.best_val = ov; run.best_idx = oi; }
    const double uv = __shfl_xor_sync(0xffffffffu, run.min_u, o);
    const long long ui = __shfl_xor_sync(0xffffffffu, run.min_u_idx, o);
    if (uv < run.min_u || (uv == run.min_u && ui < run.min_u_idx)) { run.min_u = uv; run.min_u_idx = ui; }
    run.c0 += __shfl_xor_sync(0xffffffffu, run.c0, o);
    run.c1 += __shfl_xor_sync(0xffffffffu, run.c1, o);
    run.c2 += __shfl_xor_sync(0xffffffffu, run.c2, o);
  }
  KbSweepResult r;
  r.best_val = run.best_val; r.best_idx = run.best_idx;
  r.min_u = run.min_u; r.min_u_idx = run.min_u_idx;
  r.n_fail = run.c0; r.n_fail_minus = run.c1; r.n_fail_plus = run.c2; r.n_points = 0;
  return r;
}

__device__ __forceinline__ void kb_result_merge(KbSweepResult& r, const KbSweepResult& q) {
  if (q.best_val < r.best_val || (q.best_val == r.best_val && q.best_idx < r.best_idx)) {
    r.best_val = q.best_val; r.best_idx = q.best_idx;
  }
  if (q.min_u < r.min_u || (q.min_u == r.min_u && q.min_u_idx < r.min_u_idx)) {
    r.min_u = q.min_u; r.min_u_idx = q.min_u_idx;
  }
  r.n_fail += q.n_fail; r.n_fail_minus += q.n_fail_minus; r.n_fail_plus += q.n_fail_plus;
}

// RESIDENT = false: any n; psi strips live in an L2-resident scratch slab and stream back through a
//   cp.async ring together with the fit state.
// RESIDENT = true : n_pad <= 128 (AK-MCS / early SOBO models): the inverse factor stays in shared
//   memory for the life of the CTA, the psi strip never leaves shared memory, the triangular
//   product runs without a single barrier and its row blocks are dealt to the warps in
//   boustrophedon order so every warp carries the same number of DMMAs.
template <int MASK, bool RESIDENT, bool REFINE = false>
__global__ void __launch_bounds__(256, 2)
kb_predict_kernel(KbPredictArgs P) {
  // 256 threads, 64-site strips, 2 CTAs/SM (a single 512-thread CTA with 128-site strips was measured
  // 10 % slower: the psi phase and the DMMA product of ONE CTA cannot overlap).
  constexpr int MT = 64;                         // sites per strip
  constexpr int NT = 256;                        // threads
  constexpr int LDB = MT + 4;                    // padded psi row (== 4 mod 16)
  constexpr int WCN = MT / 32;                   // warps across the strip
  extern __shared__ __align__(128) unsigned char smem_raw[];
  KbPredSmem& sm = *reinterpret_cast<KbPredSmem*>(smem_raw);
  double* xs = reinterpret_cast<double*>(smem_raw + sizeof(KbPredSmem));   // [d][64]
  double* sgw = xs + P.d * MT;
  double* sith = sgw + P.d;
  double* xtile = sith + P.d;                                             // [2][d][16] training-site tiles
  double* etile = xtile + 2 * P.d * PK_TK;                                // [2][16][PK_QIN] extra columns
  double* xraw = etile + 2 * PK_TK * PK_QIN;                                // [2][64 d] raw sites, prefetched
  double* big = xraw + (P.stage_x ? 2 * MT * P.d : 0);
  double (*sa)[PK_KC][PK_LDA] = reinterpret_cast<double (*)[PK_KC][PK_LDA]>(big);
  double (*sb)[PK_KC][LDB] = reinterpret_cast<double (*)[PK_KC][LDB]>(big + PK_STAGES * PK_KC * PK_LDA);
  const int lda_r = P.n_pad + 4;
  double* ATs = big;                                                        // [n_pad][n_pad + 4]
  double (*psiS)[LDB] = reinterpret_cast<double (*)[LDB]>(big + (size_t)P.n_pad * lda_r);
  double* exs_r = big + (size_t)P.n_pad * lda_r + (size_t)P.n_pad * LDB;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int wr = warp / WCN, wc = warp % WCN;
  const int n = P.n, d = P.d, n_pad = P.n_pad;
  constexpr bool refine = !RESIDENT && REFINE;     // (a build of its own: the common path keeps its registers)
  double* psi = RESIDENT ? nullptr : P.psi_scratch + (size_t)blockIdx.x * n_pad * MT * (refine ? 2 : 1);
  double* vbuf = psi + (size_t)n_pad * MT;        // refine: v0 = W psi of this strip
  double* ex = P.ex_scratch + (size_t)blockIdx.x * P.qx * MT;
  if (RESIDENT) {
    for (int e = tid; e < n_pad * (n_pad / 2); e += NT) {
      const int k = e / (n_pad / 2), seg = (e - k * (n_pad / 2)) * 2;
      kb_cp_async16(&ATs[(size_t)k * lda_r + seg], P.AT + (size_t)k * P.ldat + seg);
    }
    kb_cp_async_commit();
    kb_cp_async_wait<0>();
  }

  for (int i = tid; i < d; i += blockDim.x) {
    sgw[i] = P.gw[i];
    sith[i] = P.ith[i];
  }
  if (tid < 4) sm.wk[tid] = P.wk[tid];

  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  KbRunning run;
  run.best_val = INF; run.min_u = INF;
  run.best_idx = 0x7fffffffffffffffLL; run.min_u_idx = 0x7fffffffffffffffLL;
  run.c0 = run.c1 = run.c2 = 0;

  const long long nstrips = (P.m + MT - 1) / MT;
  // raw sites of the NEXT strip are prefetched (cp.async, two buffers) while this one is processed
  const bool x_aligned = ((reinterpret_cast<uintptr_t>(P.x) & 15) == 0);
  auto prefetch_x = [&](long long strip, int buf) {
    if (P.stage_x && x_aligned && strip < nstrips && (strip + 1) * MT <= P.m) {
      const double* src = P.x + (size_t)strip * MT * d;
      double* dst = xraw + (size_t)buf * MT * d;
      for (int e = tid; e < MT * d / 2; e += NT) kb_cp_async16(dst + 2 * e, src + 2 * e);
    }
    kb_cp_async_commit();
  };
  prefetch_x(blockIdx.x, 0);
  int xbuf = 0;
  for (long long strip = blockIdx.x; strip < nstrips; strip += gridDim.x, xbuf ^= 1) {
    const long long p0 = strip * MT;
    const int npt = (int)((P.m - p0 < MT) ? (P.m - p0) : MT);
    kb_cp_async_wait<0>();
    __syncthreads();
    prefetch_x(strip + gridDim.x, xbuf ^ 1);
    // ---- A1. standardise the strip ---------------------------------------
    const bool staged = P.stage_x && x_aligned && npt == MT;
    const double* xr = xraw + (size_t)xbuf * MT * d;
    for (int e = tid; e < MT * d; e += blockDim.x) {
      const int p = e / d, dim = e - p * d;
      double v = 0.0;
      if (p < npt) {
        v = staged ? xr[e] : P.x[(size_t)(p0 + p) * d + dim];
        if (P.norm_type == KB_NORM_DEFAULT) {
          const double lo = P.p0[dim], hi = P.p1[dim];
          v = ((v - lo) / (hi - lo) - 0.5) * 2.0;
        } else if (P.norm_type == KB_NORM_STD) {
          v = (v - P.p0[dim]) / P.p1[dim];
        }
      }
      xs[dim * MT + p] = v;
    }
    __syncthreads();
    // ---- A2. psi strip -> CTA-private scratch (stays in L2) ---------------
    // With few extra rows (ordinary / low-order Kriging) psi' alpha and F' R^-1 psi are
    // accumulated right here (2 DFMA per pair) instead of costing a 128-row GEMM block.
    const bool qin = (P.qx <= PK_QIN);
    // per row-group partial sums of the extra rows; aliases the (idle) GEMM pipeline buffers
    double (*exs)[PK_QIN][MT] = reinterpret_cast<double (*)[PK_QIN][MT]>(RESIDENT ? exs_r : big);
    {
      // Training sites and the extra columns of the fit state arrive in 16-row tiles through a
      // double-buffered cp.async stage; a thread owns one site and 4 rows of the tile, i.e. four
      // independent correlation chains (the phase is FP64-issue bound, not latency bound).
      const int p = tid & (MT - 1), kg = tid / MT;
      double exa[PK_QIN];
#pragma unroll
      for (int j = 0; j < PK_QIN; ++j) exa[j] = 0.0;
      const int ntile = n_pad / PK_TK;
      auto issue_tile = [&](int t, int buf) {
        double* xt = xtile + (size_t)buf * d * PK_TK;
        for (int e = tid; e < d * (PK_TK / 2); e += NT) {
          const int dim = e / (PK_TK / 2), seg = (e - dim * (PK_TK / 2)) * 2;
          kb_cp_async16(&xt[dim * PK_TK + seg], P.XT + (size_t)dim * n_pad + t * PK_TK + seg);
        }
        if (qin) {
          double* et = etile + (size_t)buf * PK_TK * PK_QIN;
          const int nseg = (P.qx + 1) >> 1;
          for (int e = tid; e < PK_TK * nseg; e += NT) {
            const int r = e / nseg, seg = (e - r * nseg) * 2;
            kb_cp_async16(&et[r * PK_QIN + seg], P.AT + (size_t)(t * PK_TK + r) * P.ldat + n_pad + seg);
          }
        }
        kb_cp_async_commit();
      };
      issue_tile(0, 0);
      for (int t = 0; t < ntile; ++t) {
        kb_cp_async_wait<0>();
        __syncthreads();
        if (t + 1 < ntile) issue_tile(t + 1, (t + 1) & 1);
        const double* xt = xtile + (size_t)(t & 1) * d * PK_TK + kg * 4;
        const double* et = etile + (size_t)(t & 1) * PK_TK * PK_QIN + kg * 4 * PK_QIN;
        KbPairAcc acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) kb_pair_init(acc[i]);
#pragma unroll 2
        for (int dim = 0; dim < d; ++dim) {
          const double x = xs[dim * MT + p], gw = sgw[dim], ith = sith[dim];
          const double2 xa = *reinterpret_cast<const double2*>(&xt[dim * PK_TK]);
          const double2 xb = *reinterpret_cast<const double2*>(&xt[dim * PK_TK + 2]);
          kb_pair_dim(acc[0], x - xa.x, gw, ith, MASK);
          kb_pair_dim(acc[1], x - xa.y, gw, ith, MASK);
          kb_pair_dim(acc[2], x - xb.x, gw, ith, MASK);
          kb_pair_dim(acc[3], x - xb.y, gw, ith, MASK);
        }
        const int k0 = t * PK_TK + kg * 4;
        double vals[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) vals[i] = kb_pair_finish(acc[i], sm.wk, MASK);   // four interleaved exp chains
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double val = vals[i];
          if (k0 + i >= n) val = 0.0;
          if (qin) {
#pragma unroll
            for (int j = 0; j < PK_QIN; ++j)
              if (j < P.qx) exa[j] = fma(val, et[i * PK_QIN + j], exa[j]);
          }
          if (RESIDENT) psiS[k0 + i][p] = val;
          else psi[(size_t)(k0 + i) * MT + p] = val;
        }
      }
      if (qin) {
#pragma unroll
        for (int j = 0; j < PK_QIN; ++j)
          if (j < P.qx) (RESIDENT ? exs_r[((size_t)kg * P.qx + j) * MT + p] : exs[kg][j][p]) = exa[j];
      }
    }
    __syncthreads();
    if (qin) {
      for (int e = tid; e < P.qx * MT; e += NT) {
        const int j = e / MT, p = e - j * MT;
        if (RESIDENT)
          ex[(size_t)j * MT + p] = (exs_r[((size_t)0 * P.qx + j) * MT + p] + exs_r[((size_t)1 * P.qx + j) * MT + p]) +
                                      (exs_r[((size_t)2 * P.qx + j) * MT + p] + exs_r[((size_t)3 * P.qx + j) * MT + p]);
        else
          ex[(size_t)j * MT + p] = (exs[0][j][p] + exs[1][j][p]) + (exs[2][j][p] + exs[3][j][p]);
      }
      __syncthreads();
    }

    // ---- B. V = AT^T-blocks x psi on the FP64 tensor pipe -----------------
    double sq[4][2];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) sq[ni][0] = sq[ni][1] = 0.0;

    if (RESIDENT) {
      // rows of L^-1 in 8-row blocks; row group wr takes blocks wr, 7-wr, 8+wr, 15-wr (equal work);
      // block rb needs k < 8 (rb + 1) only
      const int n8 = (n + 7) & ~7, nrb = n8 >> 3;
      int rbs[4], klim[4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        rbs[mi] = (mi & 1) ? (mi * 4 + 3 - wr) : (mi * 4 + wr);
        klim[mi] = (rbs[mi] < nrb) ? 2 * (rbs[mi] + 1) : 0;
      }
      double acc[4][4][2];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      const int nks = n8 >> 2;
      for (int ks = 0; ks < nks; ++ks) {
        double bf[4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ni] = psiS[ks * 4 + tq][wc * 32 + ni * 8 + gq];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
          if (ks < klim[mi]) {
            const double af = ATs[(size_t)(ks * 4 + tq) * lda_r + rbs[mi] * 8 + gq];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) kb_dmma(acc[mi][ni][0], acc[mi][ni][1], af, bf[ni]);
          }
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          sq[ni][0] = fma(acc[mi][ni][0], acc[mi][ni][0], sq[ni][0]);
          sq[ni][1] = fma(acc[mi][ni][1], acc[mi][ni][1], sq[ni][1]);
        }
    } else {
    // refine (ill-conditioned fit): three passes of the same tile stream,
    //   0: v0 = W psi -> vbuf     1: r = psi - L v0 -> psi (in place)     2: v = v0 + W r, sums of squares
    // (a pass reads only what earlier passes wrote, row block I of pass 1 its own psi rows)
    const int npass = refine ? 3 : 1;
    for (int pass = 0; pass < npass; ++pass) {
    const double* Amat = (pass == 1) ? P.LT : P.AT;
    const double* Bmat = (pass == 1) ? vbuf : psi;
    const int nblk = (qin || pass > 0) ? (n_pad + PK_BM - 1) / PK_BM : P.nrows / PK_BM;
      // One continuous chunk stream over all row blocks: the ring keeps prefetching the next
      // block's first chunks while the current block drains and runs its (register-only) epilogue.
      auto nch_of = [&](int I) { return (((I + 1) * PK_BM > n_pad) ? n_pad : (I + 1) * PK_BM) / PK_KC; };
      // load cursor: running per-thread source pointers (one add per chunk; the chunk/row-block
      // bookkeeping is kept off the address path so the copies issue in the shadow of the DMMAs)
      int Il = 0, cl = 0, ncl = nch_of(0);
      const size_t a_step = (size_t)PK_KC * P.ldat, a_j = (size_t)4 * P.ldat;
      const double* gA0 = Amat + (size_t)(tid >> 6) * P.ldat + (tid & 63) * 2;
      const double* gB0 = Bmat + (size_t)(tid >> 5) * MT + (tid & 31) * 2;
      const double* gA = gA0;
      const double* gB = gB0;
      const uint32_t sA0 = kb_smem_u32(&sa[0][tid >> 6][(tid & 63) * 2]);
      const uint32_t sB0 = kb_smem_u32(&sb[0][tid >> 5][(tid & 31) * 2]);
      auto issue_next = [&](int stage) {
        if (Il < nblk) {
          const uint32_t da = sA0 + stage * (uint32_t)(PK_KC * PK_LDA * 8), db = sB0 + stage * (uint32_t)(PK_KC * LDB * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(da + j * (uint32_t)(4 * PK_LDA * 8)), "l"(gA + j * a_j));
#pragma unroll
          for (int j = 0; j < 2; ++j)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(db + j * (uint32_t)(8 * LDB * 8)), "l"(gB + j * (8 * MT)));
          gA += a_step;
          gB += PK_KC * MT;
          if (++cl == ncl) {
            cl = 0;
            ++Il;
            ncl = nch_of(Il);
            gA = gA0 + (size_t)Il * PK_BM;
            gB = gB0;
          }
        }
        kb_cp_async_commit();
      };
#pragma unroll
      for (int s = 0; s < PK_STAGES - 1; ++s) issue_next(s);
      int gi = 0;                                       // chunks consumed so far
      for (int I = 0; I < nblk; ++I) {
        const int nch = nch_of(I);
        double acc[4][4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        for (int c = 0; c < nch; ++c, ++gi) {
          kb_cp_async_wait<PK_STAGES - 2>();
          __syncthreads();
          const int stg = gi % PK_STAGES;
          // L^-1 is lower triangular: the 32 rows of this warp need k < 128 I + 32 (wr + 1) only, so
          // in the diagonal range of the row block the upper warps drop out (their share of the
          // tensor pipe goes to the co-resident CTA)
          if (c * PK_KC >= I * PK_BM + 32 * (wr + 1)) {
            issue_next((gi + PK_STAGES - 1) % PK_STAGES);
            continue;
          }
#pragma unroll
          for (int ks = 0; ks < PK_KC / 4; ++ks) {
            double af[4], bf[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) af[mi] = sa[stg][ks * 4 + tq][wr * 32 + mi * 8 + gq];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) bf[ni] = sb[stg][ks * 4 + tq][wc * 32 + ni * 8 + gq];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) kb_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
            // refill the stage freed by the barrier above once the first products are queued
            if (ks == 0) issue_next((gi + PK_STAGES - 1) % PK_STAGES);
          }
        }
      // block epilogue: squares for L^-1 rows, raw values for the extra rows
  #pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          const int row = I * PK_BM + wr * 32 + mi * 8 + gq;
          if (row < n_pad) {
            if (refine && pass < 2) {
  #pragma unroll
              for (int ni = 0; ni < 4; ++ni) {
                const size_t o = (size_t)row * MT + wc * 32 + ni * 8 + 2 * tq;
                if (pass == 0) {
                  *reinterpret_cast<double2*>(&vbuf[o]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                } else {
                  const double2 pv = __ldcg(reinterpret_cast<const double2*>(&psi[o]));
                  *reinterpret_cast<double2*>(&psi[o]) = make_double2(pv.x - acc[mi][ni][0], pv.y - acc[mi][ni][1]);
                }
              }
            } else {
  #pragma unroll
              for (int ni = 0; ni < 4; ++ni) {
                double v0 = acc[mi][ni][0], v1 = acc[mi][ni][1];
                if (refine) {
                  const double2 vv0 = __ldcg(reinterpret_cast<const double2*>(
                      &vbuf[(size_t)row * MT + wc * 32 + ni * 8 + 2 * tq]));
                  v0 += vv0.x;
                  v1 += vv0.y;
                }
                sq[ni][0] = fma(v0, v0, sq[ni][0]);
                sq[ni][1] = fma(v1, v1, sq[ni][1]);
              }
            }
          } else if (!qin && pass == 0) {        // (the extra rows belong to the first pass: A = the fit state, B = psi)
            const int j = row - n_pad;
            if (j < P.qx) {
  #pragma unroll
              for (int ni = 0; ni < 4; ++ni) {
                const int col = wc * 32 + ni * 8 + 2 * tq;
                ex[(size_t)j * MT + col] = acc[mi][ni][0];
                ex[(size_t)j * MT + col + 1] = acc[mi][ni][1];
              }
            }
          }
        }
      }
      kb_cp_async_wait<0>();
      __syncthreads();
    }   // pass
    }
    // column sums of squares: reduce over the 8 row-groups of the warp, then over wr
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = sq[ni][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (gq == 0) sm.vv[wr][wc * 32 + ni * 8 + 2 * tq + e] = v;
      }
    __syncthreads();

    // ---- C. per-site outputs ------------------------------------------------
    if (tid < MT && tid < npt) {
      const int p = tid;
      const double vv = sm.vv[0][p] + sm.vv[1][p] + sm.vv[2][p] + sm.vv[3][p];
      double fpc, t2;
      if (P.idx == nullptr) {
        fpc = P.beta[0];
        const double u = (ex[MT + p] - 1.0) / P.LM[0];
        t2 = u * u;
      } else {
        fpc = 0.0;
        t2 = 0.0;
        const int nind = P.nind;
        for (int a = 0; a < nind; ++a) {
          double pc = 1.0;
          for (int dim = 0; dim < d; ++dim) {
            const int o = P.idx[a * d + dim];
            if (o) pc *= kb_legendre_on(o, xs[dim * MT + p]);
          }
          fpc = fma(pc, P.beta[a], fpc);
          double u = ex[(size_t)(1 + a) * MT + p] - pc;
          for (int c = 0; c < a; ++c) u = fma(-P.LM[(size_t)a * nind + c], ex[(size_t)(1 + c) * MT + p], u);
          u /= P.LM[(size_t)a * nind + a];
          ex[(size_t)(1 + a) * MT + p] = u;   // w = LM^-1 u, in place
          t2 = fma(u, u, t2);
        }
      }
      kb_acq_site(P, fpc, ex[p], vv, t2, p0 + p, run);
    }
  }

  // ---- CTA partial: (min, lowest index) + counters over threads 0..63 ------
  __syncthreads();
  if (warp < MT / 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, run.best_val, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, run.best_idx, o);
      if (ov < run.best_val || (ov == run.best_val && oi < run.best_idx)) { run.best_val = ov; run.best_idx = oi; }
      const double uv = __shfl_xor_sync(0xffffffffu, run.min_u, o);
      const long long ui = __shfl_xor_sync(0xffffffffu, run.min_u_idx, o);
      if (uv < run.min_u || (uv == run.min_u && ui < run.min_u_idx)) { run.min_u = uv; run.min_u_idx = ui; }
      run.c0 += __shfl_xor_sync(0xffffffffu, run.c0, o);
      run.c1 += __shfl_xor_sync(0xffffffffu, run.c1, o);
      run.c2 += __shfl_xor_sync(0xffffffffu, run.c2, o);
    }
    if (lane == 0) {
      KbSweepResult r;
      r.best_val = run.best_val; r.best_idx = run.best_idx;
      r.min_u = run.min_u; r.min_u_idx = run.min_u_idx;
      r.n_fail = run.c0; r.n_fail_minus = run.c1; r.n_fail_plus = run.c2; r.n_points = 0;
      sm.red[warp] = r;
    }
  }
  __syncthreads();
  if (tid == 0) {
    KbSweepResult r = sm.red[0];
    for (int wv = 1; wv < MT / 32; ++wv) {
      const KbSweepResult q = sm.red[wv];
      if (q.best_val < r.best_val || (q.best_val == r.best_val && q.best_idx < r.best_idx)) {
        r.best_val = q.best_val; r.best_idx = q.best_idx;
      }
      if (q.min_u < r.min_u || (q.min_u == r.min_u && q.min_u_idx < r.min_u_idx)) {
        r.min_u = q.min_u; r.min_u_idx = q.min_u_idx;
      }
      r.n_fail += q.n_fail; r.n_fail_minus += q.n_fail_minus; r.n_fail_plus += q.n_fail_plus;
    }
    P.partials[blockIdx.x] = r;
  }
}

// ---------------------------------------------------------------------------
// Warp-autonomous variant for small models (n <= 128, d <= 40, ordinary Kriging): the AK-MCS
// population sweep (akmcs.py:77-89 over 1e8 points with tens of training points) is bound by the
// exp() of the correlations, not by memory or the tensor pipe, so the goal is to keep every
// issue slot busy: no CTA barrier in the loop, a warp owns 32 sites (one per lane), builds its
// psi tile in a warp-private shared buffer, multiplies by the resident L^-1 with DMMA and
// finishes its sites, all with __syncwarp only.
// ---------------------------------------------------------------------------
#define PW_LDP 36          // padded psi row: 32 sites + 4 (== 4 mod 16 -> conflict-free B fragments)
#define PW_MAXD 8           // widest design of the 8-warp kernel (128 registers per thread)
#define PW_MAXD_WIDE 16     // ... of the 4-warp kernel (255 registers per thread): runtime-d loop
#define PW_MAXD_XWIDE 40    // ... of its extra-wide builds (hidimenRA, testcase/RA/testcase.py:81-88: 40 inputs): the lane's
                            // 40 coordinates still live in registers; widths padded to 24 / 32 / 40 (kb_pw_rows)

// Up to 64 training sites: 8 warps per CTA, L^-1 as a full k-major matrix, two CTAs per SM.  65 .. 128 sites:
// 4 warps per CTA (one per SM sub-partition), L^-1 packed as its lower 8x8 blocks, one CTA per SM -- the psi
// tiles of the warps (n8 x 36 doubles each) are what fills the shared memory.
static int kb_predict_warp_nw(int n8, int d) { return (n8 <= 64 && d <= PW_MAXD) ? 8 : 4; }
// Rows of the shared copy of the design (and entries of the weight vectors): the width the kernel build that serves d
// columns is compiled for.  Runtime-d builds pad to 8 / 16 / 24 / 32 / 40 with zero rows and zero weights, so their
// dimension loop carries no predicate at all (a per-dimension `dim < d` branch kept the chains of a lane from
// interleaving: the d = 40 sweep ran at 35 % of the FP64 issue rate).
__host__ __device__ __forceinline__ int kb_pw_rows(int d) {
  return d <= 3 ? d : (d <= 8 ? 8 : (d <= 16 ? 16 : ((d + 7) & ~7)));
}
static size_t kb_predict_warp_at_doubles(int n8, int nw) {
  const int nrb = n8 >> 3;
  return nw == 8 ? (size_t)n8 * (n8 + 4) : (size_t)(nrb * (nrb + 1) / 2) * 64;
}
size_t kb_predict_warp_smem_bytes(int d, int n8, int qx) {
  const int nw = kb_predict_warp_nw(n8, d);
  const int dr = kb_pw_rows(d);
  return sizeof(double) * ((size_t)dr * n8 + 2 * (size_t)dr + (size_t)qx * n8 + kb_predict_warp_at_doubles(n8, nw) +
                           nw * (size_t)n8 * PW_LDP + 8 * 32) + sizeof(KbSweepResult) * 8 + 64;
}
bool kb_predict_can_warp(int d, int n, int qx, int nind_trend) {
  const int n8 = (n + 7) & ~7;
  static const char* big_env = getenv("KB_PW_BIG");       // debug aid: KB_PW_BIG=0 keeps n > 64 on the CTA kernel
  const int nmax = (big_env && atoi(big_env) == 0) ? 64 : 128;
  static const char* xw_env = getenv("KB_PW_XWIDE");   // debug aid: KB_PW_XWIDE=0 keeps d > 16 on the CTA kernel
  const int dmax = (big_env && atoi(big_env) == 0) ? PW_MAXD
                   : ((xw_env && atoi(xw_env) == 0) ? PW_MAXD_WIDE : PW_MAXD_XWIDE);
  // (from 4 columns up a group's 32 d raw coordinates are transposed through one warp's n8 x PW_LDP psi tile)
  return n8 <= nmax && d <= dmax && nind_trend == 0 && qx == 2 && (d < 4 || 32 * d <= n8 * PW_LDP) &&
         kb_predict_warp_smem_bytes(d, n8, qx) + 1024 <= 227 * 1024;
}
int kb_predict_warp_ctas_per_sm(int d, int n, int qx) {
  return (2 * (kb_predict_warp_smem_bytes(d, (n + 7) & ~7, qx) + 1024) <= 227 * 1024) ? 2 : 1;
}

// D > 0: the design has exactly D columns (compile-time: the per-dimension weights live in registers, no
// predicated dimension loop); D = 0: any d <= PW_MAXD.  U = correlation chains in flight per lane.
// The triangular product takes the 8-row blocks in pairs (6 fragment loads per 8 DMMA instead of 10).
// SUBST (ill-conditioned fits, KbPredictArgs::refine): v = L^-1 psi by BLOCKED FORWARD SUBSTITUTION instead of the product
// with the explicit inverse -- 8-row block rb:  v_rb = W_rb,rb (psi_rb - sum_{kb < rb} L_rb,kb v_kb), the off-diagonal
// blocks from the factor itself and only the 8x8 diagonal blocks of the inverse (well conditioned each); the same DMMA
// count, v overwrites psi in the warp's tile block by block.  As accurate as the reference's triangular solve
// (tools/emulate_predictor.py: blk16 / trsv), where the explicit-inverse product is 5-30x further from an 80-bit evaluation
// once the pivots reach down to the nugget.
template <int MASK, int D, int U, int NW, int DW, bool SUBST = false>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1)
kb_predict_warp_kernel(KbPredictArgs P) {
  constexpr bool PACKED = (NW != 8) || SUBST;   // L^-1 as lower 8x8 blocks [rb (rb + 1) / 2 + kb][k % 8][row % 8]
  constexpr int NTH = NW * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // DD: the width this build is compiled for.  The runtime-d builds (DW = 8, 16, 24, 32, 40 = kb_pw_rows(d)) see
  // zero rows and zero weights beyond d: no predicate in the dimension loop.  (The host sizes the shared memory for
  // kb_pw_rows(d) >= DD rows.)
  constexpr int DD = D ? D : DW;
  // single-kernel designs: the weight of a dimension is folded into the coordinates once (x sqrt(gw) for the Gaussian,
  // x / theta otherwise), which leaves a subtract and an FMA (or add) per pair and dimension instead of three
  // (also in the SUBST builds: measured on the ill-conditioned designs of tools/illcond_check.py, the distance of the
  // variance from an 80-bit evaluation is the same with and without the fold)
  constexpr bool PRE = (MASK == 1 || MASK == 2 || MASK == 4 || MASK == 8);
  const int n = P.n, d = D ? D : P.d, n8 = (n + 7) & ~7, lda = n8 + 4;
  KbSweepResult* red = reinterpret_cast<KbSweepResult*>(smem_raw);
  double* swk = reinterpret_cast<double*>(smem_raw + sizeof(KbSweepResult) * 8);   // [4] (+4 pad)
  double* XTs = swk + 8;                    // [DD][n8]
  double* sgw = XTs + (size_t)DD * n8;      // [DD]
  double* sith = sgw + DD;                  // [DD]
  double* exts = sith + DD;                 // [2][n8]  alpha | R^-1 1
  double* ATs = exts + 2 * (size_t)n8;      // [n8][n8 + 4]  k-major L^-1, or its packed lower blocks
  double* psiAll = ATs + (PACKED ? (size_t)((n8 >> 3) * ((n8 >> 3) + 1) / 2) * 64 : (size_t)n8 * lda);  // [NW][n8][PW_LDP]
  double* vvAll = psiAll + NW * (size_t)n8 * PW_LDP;   // [NW warps][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;

  for (int e = tid; e < DD * n8; e += NTH) {
    const int dim = e / n8, k = e - dim * n8;
    double v = 0.0;
    if (dim < d && k < n) {
      v = P.XT[(size_t)dim * P.n_pad + k];
      if (PRE) v *= (MASK == 1) ? sqrt(P.gw[dim]) : P.ith[dim];
    }
    XTs[e] = v;
  }
  for (int e = tid; e < DD; e += NTH) {
    // pre-scaled builds keep the coordinate scale of each dimension in sgw
    sgw[e] = (e < d) ? (PRE ? ((MASK == 1) ? sqrt(P.gw[e]) : P.ith[e]) : P.gw[e]) : 0.0;
    sith[e] = (e < d) ? P.ith[e] : 0.0;
  }
  if (tid < 4) swk[tid] = P.wk[tid];
  for (int e = tid; e < 2 * n8; e += NTH) {
    const int j = e / n8, k = e - j * n8;
    exts[e] = (k < n) ? P.AT[(size_t)k * P.ldat + P.n_pad + j] : 0.0;
  }
  for (int e = tid; e < n8 * n8; e += NTH) {
    const int k = e / n8, i = e - k * n8;
    if (PACKED) {
      const int rbk = i >> 3, kbk = k >> 3;
      // (SUBST: the diagonal slots hold the inverse's diagonal blocks, the others the FACTOR's blocks)
      if (kbk <= rbk)
        ATs[(size_t)(rbk * (rbk + 1) / 2 + kbk) * 64 + (k & 7) * 8 + (i & 7)] =
            (SUBST && kbk < rbk) ? P.LT[(size_t)k * P.ldat + i] : P.AT[(size_t)k * P.ldat + i];
    } else {
      ATs[(size_t)k * lda + i] = P.AT[(size_t)k * P.ldat + i];
    }
  }
  __syncthreads();
  // A fragment of 8-row block rb at k-step ks (rows rb*8 + gq, k = ks*4 + tq)
  auto a_frag = [&](int rb, int ks) -> double {
    return PACKED ? ATs[(size_t)(rb * (rb + 1) / 2 + (ks >> 1)) * 64 + ((ks & 1) * 4 + tq) * 8 + gq]
                  : ATs[(size_t)(ks * 4 + tq) * lda + rb * 8 + gq];
  };

  double (*psiW)[PW_LDP] = reinterpret_cast<double (*)[PW_LDP]>(psiAll + (size_t)warp * n8 * PW_LDP);
  double* vvW = vvAll + warp * 32;
  const double beta0 = P.beta[0], lm0 = P.LM[0];
  const int nrb = n8 >> 3;
  KbRunning run;
  kb_run_init(run);
  // per-dimension weights in registers for the narrow builds (the wide ones read them from shared memory); with
  // pre-scaled coordinates rgw holds the scale of the site coordinates and rith is unused
  constexpr bool WREG = (D > 0) || (PRE && DD <= 16);
  double rgw[WREG ? DD : 1], rith[WREG ? DD : 1];
  if (WREG) {
#pragma unroll
    for (int dim = 0; dim < DD; ++dim) {
      rgw[dim] = sgw[dim];
      rith[dim] = sith[dim];
    }
  }

  const long long ngroups = (P.m + 31) / 32;
  // The raw coordinates of the warp's NEXT 32 sites are requested right after the psi phase, when xv is dead, and land
  // during the triangular product: with one or two warps per SM sub-partition nothing else hides that global-memory
  // latency (ncu, n = 100, d = 40: 27 % of the warps' samples sat on the first use of these loads).  From 4 columns up
  // the 32 d doubles of the group are read as ONE contiguous block (lane l takes elements l, l + 32, ...: full lines per
  // request instead of 32 sectors of 32 different rows) and transposed through the warp's own psi tile, which is idle
  // between the product of one group and the psi phase of the next.
  constexpr bool COAL = (DD >= 4);
  double xv[DD];
  auto load_raw = [&](long long g) {
    if (COAL) {
      const long long base = g * 32 * d, total = P.m * (long long)d;
#pragma unroll
      for (int j = 0; j < DD; ++j) {
        const long long e = base + j * 32 + lane;
        xv[j] = (j < d && e < total) ? P.x[e] : 0.0;
      }
    } else {
      const long long st = g * 32 + lane;
#pragma unroll
      for (int dim = 0; dim < DD; ++dim) xv[dim] = (dim < d && st < P.m) ? P.x[(size_t)st * d + dim] : 0.0;
    }
  };
  const long long gstride = (long long)gridDim.x * NW;
  if ((long long)blockIdx.x * NW + warp < ngroups) load_raw((long long)blockIdx.x * NW + warp);
  for (long long grp = (long long)blockIdx.x * NW + warp; grp < ngroups; grp += gstride) {
    const long long site = grp * 32 + lane;
    const bool valid = site < P.m;
    // ---- this lane's site, standardised ------------------------------------
    if (COAL) {
      double* stg = &psiW[0][0];                 // 32 d <= n8 * PW_LDP doubles (kb_predict_can_warp)
#pragma unroll
      for (int j = 0; j < DD; ++j)
        if (j < d) stg[j * 32 + lane] = xv[j];
      __syncwarp();
#pragma unroll
      for (int dim = 0; dim < DD; ++dim) xv[dim] = (dim < d) ? stg[lane * d + dim] : 0.0;
      __syncwarp();
    }
    // (sites past the end hold zeros: whatever the maps below make of them stays in their own lane and is dropped)
    if (P.norm_type == KB_NORM_DEFAULT) {
#pragma unroll
      for (int dim = 0; dim < DD; ++dim)
        if (dim < d) {
          const double lo = P.p0[dim], hi = P.p1[dim];
          xv[dim] = ((xv[dim] - lo) / (hi - lo) - 0.5) * 2.0;
        }
    } else if (P.norm_type == KB_NORM_STD) {
#pragma unroll
      for (int dim = 0; dim < DD; ++dim)
        if (dim < d) xv[dim] = (xv[dim] - P.p0[dim]) / P.p1[dim];
    }
    if (PRE) {
#pragma unroll
      for (int dim = 0; dim < DD; ++dim) xv[dim] *= WREG ? rgw[dim] : sgw[dim];
    }
    // ---- psi column of this site; psi' alpha and 1' R^-1 psi on the way --------
    double ex0 = 0.0, ex1 = 0.0;
    // U training sites per pass (n8 is a multiple of 8, U of 2 or 4): all shared-memory reads first, then the U
    // independent correlation chains (branch-free exp: the compiler interleaves them), then the stores
    for (int k = 0; k < n8; k += U) {
      KbPairAcc acc[U];
      double e0[U], e1[U], val[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        kb_pair_init(acc[u]);
        e0[u] = exts[k + u];
        e1[u] = exts[n8 + k + u];
      }
#pragma unroll
      for (int dim = 0; dim < DD; ++dim) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double diff = xv[dim] - XTs[dim * n8 + k + u];
          if (PRE) kb_pair_dim_scaled(acc[u], diff, MASK);
          else if (WREG) kb_pair_dim(acc[u], diff, rgw[dim], rith[dim], MASK);
          else kb_pair_dim(acc[u], diff, sgw[dim], sith[dim], MASK);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        val[u] = kb_pair_finish(acc[u], swk, MASK);
        if (k + u >= n) val[u] = 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ex0 = fma(val[u], e0[u], ex0);
        ex1 = fma(val[u], e1[u], ex1);
        psiW[k + u][lane] = val[u];
      }
    }
    __syncwarp();
    if (grp + gstride < ngroups) load_raw(grp + gstride);
    // ---- v = L^-1 psi for the 32 sites, keep the column sums of squares --------
    double sq[4][2];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) sq[cb][0] = sq[cb][1] = 0.0;
    int rb = 0;
    if (SUBST) {
      // finish one 8-row block: right-hand side psi_rb - acc (in place), v_rb = W_rb,rb rhs, v_rb over psi_rb, squares
      auto finish = [&](int r8, double (&acc)[4][2]) {
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          double2* pr = reinterpret_cast<double2*>(&psiW[r8 * 8 + gq][cb * 8 + 2 * tq]);
          const double2 pv = *pr;
          *pr = make_double2(pv.x - acc[cb][0], pv.y - acc[cb][1]);
          acc[cb][0] = acc[cb][1] = 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int ks = 2 * r8; ks < 2 * r8 + 2; ++ks) {
          const double af = a_frag(r8, ks);                 // diagonal slot: the inverse's 8x8 block
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) kb_dmma(acc[cb][0], acc[cb][1], af, psiW[ks * 4 + tq][cb * 8 + gq]);
        }
        __syncwarp();
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          *reinterpret_cast<double2*>(&psiW[r8 * 8 + gq][cb * 8 + 2 * tq]) = make_double2(acc[cb][0], acc[cb][1]);
          sq[cb][0] = fma(acc[cb][0], acc[cb][0], sq[cb][0]);
          sq[cb][1] = fma(acc[cb][1], acc[cb][1], sq[cb][1]);
        }
        __syncwarp();
      };
      // two row blocks per pass share the v fragments of the columns left of both (as the plain product does); the
      // second one then takes the first one's fresh v_rb through the factor's block (rb + 1, rb)
      for (; rb + 1 < nrb; rb += 2) {
        double acc0[4][2], acc1[4][2];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) acc0[cb][0] = acc0[cb][1] = acc1[cb][0] = acc1[cb][1] = 0.0;
        for (int ks = 0; ks < 2 * rb; ++ks) {
          const double a0 = a_frag(rb, ks), a1 = a_frag(rb + 1, ks);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            const double bf = psiW[ks * 4 + tq][cb * 8 + gq];
            kb_dmma(acc0[cb][0], acc0[cb][1], a0, bf);
            kb_dmma(acc1[cb][0], acc1[cb][1], a1, bf);
          }
        }
        finish(rb, acc0);
#pragma unroll
        for (int ks = 2 * rb; ks < 2 * rb + 2; ++ks) {
          const double a1 = a_frag(rb + 1, ks);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) kb_dmma(acc1[cb][0], acc1[cb][1], a1, psiW[ks * 4 + tq][cb * 8 + gq]);
        }
        finish(rb + 1, acc1);
      }
      for (; rb < nrb; ++rb) {
        double acc[4][2];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) acc[cb][0] = acc[cb][1] = 0.0;
        for (int ks = 0; ks < 2 * rb; ++ks) {             // sum_{kb < rb} L_rb,kb v_kb  (rows < 8 rb of the tile hold v)
          const double af = a_frag(rb, ks);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) kb_dmma(acc[cb][0], acc[cb][1], af, psiW[ks * 4 + tq][cb * 8 + gq]);
        }
        finish(rb, acc);
      }
    }
    {
      // two 8-row blocks per pass: the psi fragments of a k-step serve both
      for (; rb + 1 < nrb; rb += 2) {
        double acc[2][4][2];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) acc[h][cb][0] = acc[h][cb][1] = 0.0;
        const int kmax = 2 * (rb + 1);               // block rb needs k < 8 (rb + 1), block rb + 1 eight more
        for (int ks = 0; ks < kmax; ++ks) {
          const double a0 = a_frag(rb, ks);
          const double a1 = a_frag(rb + 1, ks);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            const double bf = psiW[ks * 4 + tq][cb * 8 + gq];
            kb_dmma(acc[0][cb][0], acc[0][cb][1], a0, bf);
            kb_dmma(acc[1][cb][0], acc[1][cb][1], a1, bf);
          }
        }
#pragma unroll
        for (int ks = kmax; ks < kmax + 2; ++ks) {
          const double a1 = a_frag(rb + 1, ks);
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) kb_dmma(acc[1][cb][0], acc[1][cb][1], a1, psiW[ks * 4 + tq][cb * 8 + gq]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            sq[cb][0] = fma(acc[h][cb][0], acc[h][cb][0], sq[cb][0]);
            sq[cb][1] = fma(acc[h][cb][1], acc[h][cb][1], sq[cb][1]);
          }
      }
    }
    for (; rb < nrb; ++rb) {
      double acc[4][2];
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) acc[cb][0] = acc[cb][1] = 0.0;
      const int kmax = 2 * (rb + 1);
      for (int ks = 0; ks < kmax; ++ks) {
        const double af = a_frag(rb, ks);
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) kb_dmma(acc[cb][0], acc[cb][1], af, psiW[ks * 4 + tq][cb * 8 + gq]);
      }
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        sq[cb][0] = fma(acc[cb][0], acc[cb][0], sq[cb][0]);
        sq[cb][1] = fma(acc[cb][1], acc[cb][1], sq[cb][1]);
      }
    }
#pragma unroll
    for (int cb = 0; cb < 4; ++cb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = sq[cb][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (gq == 0) vvW[cb * 8 + 2 * tq + e] = v;
      }
    __syncwarp();
    const double vv = vvW[lane];
    __syncwarp();
    // ---- outputs of this lane's site ------------------------------------------
    if (valid) {
      const double u = (ex1 - 1.0) / lm0;
      kb_acq_site(P, beta0, ex0, vv, u * u, site, run);
    }
  }
  // ---- CTA partial -------------------------------------------------------------
  const KbSweepResult wr_ = kb_run_warp_reduce(run);
  if (lane == 0) red[warp] = wr_;
  __syncthreads();
  if (tid == 0) {
    KbSweepResult r = red[0];
    for (int w = 1; w < NW; ++w) kb_result_merge(r, red[w]);
    P.partials[blockIdx.x] = r;
  }
}

// Deterministic final reduction over the per-CTA partials (lowest index on ties; the counters are integers):
// one warp, lanes stride over the partials, then a shuffle tree.
__global__ void kb_sweep_final_kernel(const KbSweepResult* __restrict__ partials, int nparts,
                                      long long npoints, KbSweepResult* __restrict__ result) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  KbRunning run;
  kb_run_init(run);
  for (int i = threadIdx.x; i < nparts; i += 32) {
    const KbSweepResult q = partials[i];
    if (q.best_val < run.best_val || (q.best_val == run.best_val && q.best_idx < run.best_idx)) {
      run.best_val = q.best_val; run.best_idx = q.best_idx;
    }
    if (q.min_u < run.min_u || (q.min_u == run.min_u && q.min_u_idx < run.min_u_idx)) {
      run.min_u = q.min_u; run.min_u_idx = q.min_u_idx;
    }
    run.c0 += q.n_fail; run.c1 += q.n_fail_minus; run.c2 += q.n_fail_plus;
  }
  KbSweepResult r = kb_run_warp_reduce(run);
  if (threadIdx.x == 0) {
    r.n_points = npoints;
    *result = r;
  }
}

template <int MASK, bool RESIDENT, bool REFINE = false>
static cudaError_t kb_predict_launch(cudaStream_t st, const KbPredictArgs& a, int grid, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kb_predict_kernel<MASK, RESIDENT, REFINE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kb_predict_kernel<MASK, RESIDENT, REFINE><<<grid, 256, smem, st>>>(a);
  return cudaGetLastError();
}

#define KB_DISPATCH_MASK_P(MASKVAR, CALL)                                \
  switch (MASKVAR) {                                                     \
    case 1: { constexpr int M_ = 1; CALL; } break;                       \
    case 2: { constexpr int M_ = 2; CALL; } break;                       \
    case 4: { constexpr int M_ = 4; CALL; } break;                       \
    case 8: { constexpr int M_ = 8; CALL; } break;                       \
    default: { constexpr int M_ = 15; CALL; } break;                     \
  }

template <int MASK, int D, int U, int NW, int DW, bool SUBST = false>
static cudaError_t kb_predict_warp_launch_w(cudaStream_t st, const KbPredictArgs& a, int grid, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kb_predict_warp_kernel<MASK, D, U, NW, DW, SUBST>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kb_predict_warp_kernel<MASK, D, U, NW, DW, SUBST><<<grid, NW * 32, smem, st>>>(a);
  return cudaGetLastError();
}
template <int MASK, int D, int U, int DW = 8, bool SUBST = false>
static cudaError_t kb_predict_warp_launch_v(cudaStream_t st, const KbPredictArgs& a, int grid, size_t smem) {
  if (((a.n + 7) & ~7) <= 64 && a.d <= PW_MAXD) return kb_predict_warp_launch_w<MASK, D, U, 8, DW, SUBST>(st, a, grid, smem);
  // one warp per SM sub-partition: twice the correlation chains per lane (registers are plentiful at 128
  // threads; measured +3 % at d = 2, +17 % at d = 6)
  return kb_predict_warp_launch_w<MASK, D, 2 * U, 4, DW, SUBST>(st, a, grid, smem);
}
// The kernel is compiled for every design width it serves exactly (1 .. PW_MAXD; the reliability problems AK-MCS
// sweeps are 2-D) and, beyond that, for the padded widths 8 / 16 / 24 / 32 / 40 (kb_pw_rows); KB_PW_D=0 forces the
// padded runtime-d build for narrow designs too (debug aid).  Ill-conditioned fits (a.refine) take the blocked-substitution
// builds: exact for 1 .. 3 columns (the dense low-dimensional designs where it happens), padded beyond.
template <int MASK, bool SUBST>
static cudaError_t kb_predict_warp_launch_s(cudaStream_t st, const KbPredictArgs& a, int grid, size_t smem) {
  static const char* env_d = getenv("KB_PW_D");
  const int dsel = (env_d && atoi(env_d) == 0 && a.d > 3) ? 0 : a.d;
  switch (dsel) {
    case 1: return kb_predict_warp_launch_v<MASK, 1, 4, 8, SUBST>(st, a, grid, smem);
    case 2: return kb_predict_warp_launch_v<MASK, 2, 4, 8, SUBST>(st, a, grid, smem);
    case 3: return kb_predict_warp_launch_v<MASK, 3, 4, 8, SUBST>(st, a, grid, smem);
    default: break;
  }
  if constexpr (MASK != 15 && !SUBST) {       // composite kernels carry both weight sets: wider designs would spill
    switch (dsel) {
      case 4: return kb_predict_warp_launch_v<MASK, 4, 4>(st, a, grid, smem);
      case 5: return kb_predict_warp_launch_v<MASK, 5, 2>(st, a, grid, smem);
      case 6: return kb_predict_warp_launch_v<MASK, 6, 2>(st, a, grid, smem);
      case 7: return kb_predict_warp_launch_v<MASK, 7, 2>(st, a, grid, smem);
      case 8: return kb_predict_warp_launch_v<MASK, 8, 2>(st, a, grid, smem);
      default: break;
    }
  }
  // padded runtime-d builds.  A lane's chains are serial over the dimensions (one dependent FMA per dimension), and
  // with one warp per SM sub-partition nothing else hides that latency: a pass over U pairs issues for 4 U DD cycles
  // while each chain needs DD FMA latencies, so U = 8 for the single-kernel builds (the composite kernel issues four
  // times the work per dimension and keeps U = 2 .. 4 within its registers).
  switch (kb_pw_rows(a.d)) {
    case 8: return kb_predict_warp_launch_v<MASK, 0, (MASK == 15 ? 2 : 4), 8, SUBST>(st, a, grid, smem);
    case 16: return kb_predict_warp_launch_w<MASK, 0, (MASK == 15 ? 4 : 8), 4, 16, SUBST>(st, a, grid, smem);
    case 24: return kb_predict_warp_launch_w<MASK, 0, (MASK == 15 ? 2 : 8), 4, 24, SUBST>(st, a, grid, smem);
    case 32: return kb_predict_warp_launch_w<MASK, 0, (MASK == 15 ? 2 : 8), 4, 32, SUBST>(st, a, grid, smem);
    default: return kb_predict_warp_launch_w<MASK, 0, (MASK == 15 ? 2 : 8), 4, PW_MAXD_XWIDE, SUBST>(st, a, grid, smem);
  }
}
template <int MASK>
static cudaError_t kb_predict_warp_launch(cudaStream_t st, const KbPredictArgs& a, int grid, size_t smem) {
  return a.refine ? kb_predict_warp_launch_s<MASK, true>(st, a, grid, smem)
                  : kb_predict_warp_launch_s<MASK, false>(st, a, grid, smem);
}

cudaError_t kb_launch_predict(cudaStream_t st, const KbPredictArgs& a, int grid, KbSweepResult* result_dev) {
  const size_t smem = (a.resident == 2) ? kb_predict_warp_smem_bytes(a.d, (a.n + 7) & ~7, a.qx)
                                        : kb_predict_smem_bytes(a.d, a.resident != 0, a.n_pad, a.qx, a.stage_x);
  cudaError_t err = cudaSuccess;
  if (a.resident == 2) {
    KB_DISPATCH_MASK_P(a.kmask, (err = kb_predict_warp_launch<M_>(st, a, grid, smem)));
  } else if (a.resident) {
    KB_DISPATCH_MASK_P(a.kmask, (err = kb_predict_launch<M_, true>(st, a, grid, smem)));
  } else if (a.refine) {
    KB_DISPATCH_MASK_P(a.kmask, (err = kb_predict_launch<M_, false, true>(st, a, grid, smem)));
  } else {
    KB_DISPATCH_MASK_P(a.kmask, (err = kb_predict_launch<M_, false>(st, a, grid, smem)));
  }
  if (err != cudaSuccess) return err;
  if (result_dev) {
    kb_sweep_final_kernel<<<1, 32, 0, st>>>(a.partials, grid, a.m, result_dev);
    err = cudaGetLastError();
  }
  return err;
}

// ---------------------------------------------------------------------------------------------------
// A handful of sites (m <= KB_FEW_MAX) on a large model.  The reference's local acquisition optimisers call
// `prediction` one design at a time (acquifunc_opt.py:45-76, EHVIcomputation.py:40-45); the strip kernel
// above would push that single strip through ONE CTA (n^2 x 64 flops and the whole inverse factor through
// one SM: 0.48 ms at n = 1000).  Here the work is spread over n_pad / 64 CTAs twice:
//   F1  CTA c evaluates psi for training sites [64 c, 64 c + 64) x all m sites and its share of
//       psi' alpha, psi' R^-1 F;
//   F2  CTA c owns columns [64 c, 64 c + 64) of L^-1 psi (a tall-skinny product streamed once from L2/HBM)
//       and their sum of squares; the last CTA to finish (atomic ticket) adds the partials in a fixed order
//       and finishes the sites exactly as the strip kernel does.
// ---------------------------------------------------------------------------------------------------
template <int MASK>
__global__ void __launch_bounds__(256) kb_predict_few_psi_kernel(KbPredictArgs P, KbFewScratch S) {
  extern __shared__ __align__(16) double fsm[];
  const int d = P.d, n_pad = P.n_pad, slice = blockIdx.y;
  const int m = (int)((P.m - (long long)slice * KB_FEW_MAX < KB_FEW_MAX) ? (P.m - (long long)slice * KB_FEW_MAX) : KB_FEW_MAX);
  const double* xsl = P.x + (size_t)slice * KB_FEW_MAX * d;          // this slice's raw sites
  double* Spsi = S.psi + (size_t)slice * n_pad * KB_FEW_MAX;
  double* Sex = S.ex + (size_t)slice * gridDim.x * P.qx * KB_FEW_MAX;
  double* xs = fsm;                        // [KB_FEW_MAX][d] standardised sites
  double* sgw = xs + KB_FEW_MAX * d;
  double* sith = sgw + d;
  double* swk = sith + d;                  // [4]
  double* psis = swk + 4;                  // [64][KB_FEW_MAX]
  const int tid = threadIdx.x, cb = blockIdx.x;
  for (int e = tid; e < KB_FEW_MAX * d; e += 256) {
    const int sidx = e / d, dim = e - sidx * d;
    double v = 0.0;
    if (sidx < m) {
      v = xsl[(size_t)sidx * d + dim];
      if (P.norm_type == KB_NORM_DEFAULT) {
        const double lo = P.p0[dim], hi = P.p1[dim];
        v = ((v - lo) / (hi - lo) - 0.5) * 2.0;
      } else if (P.norm_type == KB_NORM_STD) {
        v = (v - P.p0[dim]) / P.p1[dim];
      }
    }
    xs[e] = v;
  }
  for (int i = tid; i < d; i += 256) { sgw[i] = P.gw[i]; sith[i] = P.ith[i]; }
  if (tid < 4) swk[tid] = P.wk[tid];
  __syncthreads();
  const int kl = tid & 63, sg = tid >> 6, k = cb * 64 + kl;
  KbPairAcc acc[KB_FEW_MAX / 4];
#pragma unroll
  for (int t = 0; t < KB_FEW_MAX / 4; ++t) kb_pair_init(acc[t]);
  for (int dim = 0; dim < d; ++dim) {
    const double xk = __ldg(P.XT + (size_t)dim * n_pad + k), gw = sgw[dim], ith = sith[dim];
#pragma unroll
    for (int t = 0; t < KB_FEW_MAX / 4; ++t) kb_pair_dim(acc[t], xs[(sg + 4 * t) * d + dim] - xk, gw, ith, MASK);
  }
#pragma unroll
  for (int t = 0; t < KB_FEW_MAX / 4; ++t) {
    const int sidx = sg + 4 * t;
    double val = kb_pair_finish(acc[t], swk, MASK);
    if (k >= P.n || sidx >= m) val = 0.0;
    psis[kl * KB_FEW_MAX + sidx] = val;
    Spsi[(size_t)k * KB_FEW_MAX + sidx] = val;
  }
  __syncthreads();
  // this block's share of the extra columns: ex[j][s] += sum_k psi[k][s] AT[k][n_pad + j]
  for (int e = tid; e < P.qx * KB_FEW_MAX; e += 256) {
    const int j = e / KB_FEW_MAX, sidx = e - j * KB_FEW_MAX;
    const double* col = P.AT + (size_t)cb * 64 * P.ldat + n_pad + j;
    double a0 = 0.0, a1 = 0.0;
    for (int r = 0; r < 64; r += 2) {
      a0 = fma(psis[r * KB_FEW_MAX + sidx], __ldg(col + (size_t)r * P.ldat), a0);
      a1 = fma(psis[(r + 1) * KB_FEW_MAX + sidx], __ldg(col + (size_t)(r + 1) * P.ldat), a1);
    }
    Sex[((size_t)cb * P.qx + j) * KB_FEW_MAX + sidx] = a0 + a1;
  }
}

#define KB_FEW_CW 16      // columns of L^-1 psi per block of the second kernel
#define KB_FEW_KG 16      // row groups (threads per column)
__global__ void __launch_bounds__(256) kb_predict_few_gemv_kernel(KbPredictArgs P, KbFewScratch S,
                                                                  KbSweepResult* __restrict__ result, int nexb) {
  __shared__ double part[KB_FEW_KG][KB_FEW_CW][KB_FEW_MAX + 1];
  __shared__ int s_last;
  const int tid = threadIdx.x, cb = blockIdx.x, ncb = gridDim.x;
  const int il = tid & (KB_FEW_CW - 1), kg = tid / KB_FEW_CW, i = cb * KB_FEW_CW + il, slice = blockIdx.y;
  const int m = (int)((P.m - (long long)slice * KB_FEW_MAX < KB_FEW_MAX) ? (P.m - (long long)slice * KB_FEW_MAX) : KB_FEW_MAX);
  const double* Spsi = S.psi + (size_t)slice * P.n_pad * KB_FEW_MAX;
  double* Sex = S.ex + (size_t)slice * nexb * P.qx * KB_FEW_MAX;
  double* Svv = S.vv + (size_t)slice * ncb * KB_FEW_MAX;
  const double* xsl = P.x + (size_t)slice * KB_FEW_MAX * P.d;
  double acc[KB_FEW_MAX];
#pragma unroll
  for (int t = 0; t < KB_FEW_MAX; ++t) acc[t] = 0.0;
  const int kend = cb * KB_FEW_CW + KB_FEW_CW - 1;       // L^-1 is lower triangular: rows k <= i only
  const double* col = P.AT + i;
#pragma unroll 4
  for (int k = kg; k <= kend; k += KB_FEW_KG) {
    const double a = __ldg(col + (size_t)k * P.ldat);
    const double2* pk = reinterpret_cast<const double2*>(Spsi + (size_t)k * KB_FEW_MAX);
#pragma unroll
    for (int t = 0; t < KB_FEW_MAX / 2; ++t) {
      const double2 pv = __ldg(pk + t);                  // written by the previous kernel, read-only here
      acc[2 * t] = fma(a, pv.x, acc[2 * t]);
      acc[2 * t + 1] = fma(a, pv.y, acc[2 * t + 1]);
    }
  }
#pragma unroll
  for (int t = 0; t < KB_FEW_MAX; ++t) part[kg][il][t] = acc[t];
  __syncthreads();
  if (tid < KB_FEW_MAX) {
    double s2 = 0.0;
    for (int c = 0; c < KB_FEW_CW; ++c) {
      double v = 0.0;
#pragma unroll
      for (int g = 0; g < KB_FEW_KG; ++g) v += part[g][c][tid];
      s2 = fma(v, v, s2);
    }
    Svv[(size_t)cb * KB_FEW_MAX + tid] = s2;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(S.counter + slice, 1) == ncb - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last block: fixed-order sums of the partials, then the per-site outputs (as section C above) ----
  KbRunning run;
  kb_run_init(run);
  if (tid < m) {
    const int sidx = tid, d = P.d, qx = P.qx;
    double vv = 0.0;
    for (int c = 0; c < ncb; ++c) vv += __ldcg(&Svv[(size_t)c * KB_FEW_MAX + sidx]);
    double* ex0 = Sex + sidx;                        // block 0's slots hold the totals: ex0[j * KB_FEW_MAX]
    for (int j = 0; j < qx; ++j) {
      double t = 0.0;
      for (int c = 0; c < nexb; ++c) t += __ldcg(&Sex[((size_t)c * qx + j) * KB_FEW_MAX + sidx]);
      ex0[(size_t)j * KB_FEW_MAX] = t;
    }
    double fpc, t2;
    if (P.idx == nullptr) {
      fpc = P.beta[0];
      const double u = (ex0[KB_FEW_MAX] - 1.0) / P.LM[0];
      t2 = u * u;
    } else {
      fpc = 0.0;
      t2 = 0.0;
      const int nind = P.nind;
      for (int a = 0; a < nind; ++a) {
        double pc = 1.0;
        for (int dim = 0; dim < d; ++dim) {
          const int o = P.idx[a * d + dim];
          if (o) {
            double v = xsl[(size_t)sidx * d + dim];
            if (P.norm_type == KB_NORM_DEFAULT) {
              const double lo = P.p0[dim], hi = P.p1[dim];
              v = ((v - lo) / (hi - lo) - 0.5) * 2.0;
            } else if (P.norm_type == KB_NORM_STD) {
              v = (v - P.p0[dim]) / P.p1[dim];
            }
            pc *= kb_legendre_on(o, v);
          }
        }
        fpc = fma(pc, P.beta[a], fpc);
        double u = ex0[(size_t)(1 + a) * KB_FEW_MAX] - pc;
        for (int c = 0; c < a; ++c) u = fma(-P.LM[(size_t)a * nind + c], ex0[(size_t)(1 + c) * KB_FEW_MAX], u);
        u /= P.LM[(size_t)a * nind + a];
        ex0[(size_t)(1 + a) * KB_FEW_MAX] = u;
        t2 = fma(u, u, t2);
      }
    }
    kb_acq_site(P, fpc, ex0[0], vv, t2, (long long)slice * KB_FEW_MAX + sidx, run);
  }
  if (tid < 32) {
    KbSweepResult r = kb_run_warp_reduce(run);
    if (tid == 0) {
      r.n_points = m;
      S.slices[slice] = r;                 // merged over the slices by kb_sweep_final_kernel
      if (result && gridDim.y == 1) { r.n_points = P.m; *result = r; }
      S.counter[slice] = 0;
    }
  }
}

template <int MASK>
static cudaError_t kb_predict_few_psi_launch(cudaStream_t st, const KbPredictArgs& a, const KbFewScratch& S, int ncb,
                                             int ny, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kb_predict_few_psi_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  kb_predict_few_psi_kernel<MASK><<<dim3(ncb, ny), 256, smem, st>>>(a, S);
  return cudaGetLastError();
}

int kb_predict_few_slices(long long m) { return (int)((m + KB_FEW_MAX - 1) / KB_FEW_MAX); }

// doubles of scratch for ny slices of KB_FEW_MAX sites
size_t kb_predict_few_scratch_doubles(int n_pad, int qx, int ny) {
  const size_t ncb = n_pad / 64, ncb2 = n_pad / KB_FEW_CW;
  const size_t per_slice = (size_t)n_pad * KB_FEW_MAX + ncb * qx * KB_FEW_MAX + ncb2 * KB_FEW_MAX;
  return per_slice * ny + (size_t)ny * (sizeof(KbSweepResult) / sizeof(double)) + (size_t)(ny + 1) / 2 + 2;
}

cudaError_t kb_launch_predict_few(cudaStream_t st, const KbPredictArgs& a, double* scratch, int ny_cap,
                                  KbSweepResult* result_dev) {
  const int ncb = a.n_pad / 64, ncb2 = a.n_pad / KB_FEW_CW, ny = kb_predict_few_slices(a.m);
  if (ny > ny_cap) return cudaErrorInvalidValue;
  KbFewScratch S;
  S.psi = scratch;
  S.ex = S.psi + (size_t)ny_cap * a.n_pad * KB_FEW_MAX;
  S.vv = S.ex + (size_t)ny_cap * ncb * a.qx * KB_FEW_MAX;
  S.slices = reinterpret_cast<KbSweepResult*>(S.vv + (size_t)ny_cap * ncb2 * KB_FEW_MAX);
  S.counter = reinterpret_cast<int*>(S.slices + ny_cap);
  const size_t smem = sizeof(double) * ((size_t)KB_FEW_MAX * a.d + 2 * a.d + 4 + 64 * KB_FEW_MAX);
  cudaError_t err = cudaSuccess;
  KB_DISPATCH_MASK_P(a.kmask, (err = kb_predict_few_psi_launch<M_>(st, a, S, ncb, ny, smem)));
  if (err != cudaSuccess) return err;
  kb_predict_few_gemv_kernel<<<dim3(ncb2, ny), 256, 0, st>>>(a, S, result_dev, ncb);
  err = cudaGetLastError();
  if (err == cudaSuccess && result_dev && ny > 1) {
    kb_sweep_final_kernel<<<1, 32, 0, st>>>(S.slices, ny, a.m, result_dev);
    err = cudaGetLastError();
  }
  return err;
}

// Largest persistent grid for the predict kernel: SMs x resident CTAs per SM.
int kb_predict_max_grid(int kmask, int d, int num_sms, int resident, int n_pad, int qx, int stage_x) {
  const size_t smem = kb_predict_smem_bytes(d, resident != 0, n_pad, qx, stage_x);
  if (smem > 227 * 1024) return 0;
  int occ = 0;
  cudaError_t err = cudaSuccess;
  if (resident) {
    KB_DISPATCH_MASK_P(kmask, {
      err = cudaFuncSetAttribute(kb_predict_kernel<M_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err == cudaSuccess)
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb_predict_kernel<M_, true>, 256, smem);
    });
  } else {
    KB_DISPATCH_MASK_P(kmask, {
      err = cudaFuncSetAttribute(kb_predict_kernel<M_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err == cudaSuccess)
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb_predict_kernel<M_, false>, 256, smem);
    });
  }
  if (err != cudaSuccess) { cudaGetLastError(); return 0; }
  return occ * num_sms;
}
