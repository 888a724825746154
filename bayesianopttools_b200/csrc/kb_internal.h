// krig-b200 internal declarations shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/krigb200.h"

#ifndef KB_TILE
#define KB_TILE 64
#define KB_THREADS 256
#endif
#define KB_MAX_D 200  // dimensions the shared-memory X tiles are sized for

// Bulk tile task kinds of the Cholesky task graph (kb_chol.cu); task.w = kind | first K tile column << 8.
enum { KB_TASK_CHAIN = 0, KB_TASK_OFF = 1, KB_TASK_SCHUR = 2, KB_TASK_PRE = 3, KB_TASK_THIN = 4,
       KB_TASK_SCHUR_THIN = 5, KB_TASK_OFF2 = 6, KB_TASK_PRE_A = 7, KB_TASK_SCHUR_THIN_A = 8, KB_TASK_OFFS = 9 };
#define KB_CHOL_CTL_WORDS (8 + 1024)

// Geometry of one augmented matrix  [ R ; rhs rows ]  (row-major, lower tiles only).
//   tile rows 0..nb-1        : R (n padded to nb*64 with identity)
//   tile rows nb..nb+sb-1    : y, F columns as ROWS (they become L^-1 [y F]) and,
//                              in tile columns nb.., the Schur block -Z Z^T
//   tile rows nb+sb..NT-1    : identity rows (fit mode only) -> L^-T
struct KbAugGeom {
  int n, n_pad, nb;   // training points, padded, tile rows of R
  int nind, q;        // trend columns, q = 1 + nind
  int sb;             // rhs tile rows that take part in the Schur block
  int ib;             // identity tile rows (0 in population mode, nb in fit mode)
  int NT;             // nb + sb + ib
  int ld;             // (nb + sb) * 64 doubles per row
  size_t stride;      // doubles per matrix = NT*64*ld
  int seated;         // population mode, q <= 16, (n mod 64) + q <= 64: the right-hand sides [y F]^T sit in the
                      // padding rows n .. n+q-1 of the last tile row of R (sb = 0, no right-hand-side or Schur
                      // tasks: the regular tile tasks produce Z and the last chain step the Schur corner)
};

// Per-candidate decoded hyper-parameters (device arrays, B rows each).
struct KbHyper {
  double* gw;      // [B][d]   1/(2 theta^2)  (KPLS folded)
  double* ith;     // [B][d]   1/theta
  double* wk;      // [B][4]   composite weight per kernel id
  double* eps;     // [B]      10^nugget
  double* sigma2;  // [B]      10^x[nth] when trainvar
  double* nugget;  // [B]      log10 nugget actually used
  double* gk;      // [B][8]   KPLS only: 1/(2 theta_k^2) per PLS component (k < h <= 8)
};
#define KB_KPLS_MAX_H 8

struct KbModelDev {
  int n, d, nind, n_pad, nb;
  int kmask;             // OR of 1<<kernel_id
  const double* XT;      // [d][n_pad]  training sites, structure-of-arrays, zero padded
  const double* y;       // [n]
  const double* F;       // [n][nind]
};

// --- kb_corr.cu --------------------------------------------------------------
cudaError_t kb_launch_decode(cudaStream_t st, int B, int nbhyp, const double* xpop_dev, int d, int nth,
                             int trainvar, int has_nugget, int has_weights, double fixed_nugget,
                             int nkernel, const int* kernel_ids_dev, const double* pls_dev /*d*h or null*/,
                             KbHyper hp);
cudaError_t kb_launch_build_aug(cudaStream_t st, const KbModelDev& m, const KbAugGeom& g, int B,
                                KbHyper hp, double* aug, int skip_r_tiles);
// KPLS fast path (kernelfunc.py:43-49): the theta-independent per-component squared distances
// D_k[a][b] = sum_i w_ik^2 (x_ai - x_bi)^2 of every lower tile are computed once per design ...
cudaError_t kb_launch_kpls_dist(cudaStream_t st, const KbModelDev& m, int h, const double* pls, double* dtiles);
// ... and each candidate's correlation tiles are exp(-sum_k D_k gk_k) from them.
cudaError_t kb_launch_build_aug_kpls(cudaStream_t st, const KbModelDev& m, const KbAugGeom& g, int B, int h,
                                     KbHyper hp, const double* dtiles, double* aug);
cudaError_t kb_launch_corr_general(cudaStream_t st, int n, int m, int d, const double* XNT, int ldn,
                                   const double* XMT, int ldm, const double* gw, const double* ith,
                                   const double* wk, int kmask, double* out, int ldo);
cudaError_t kb_launch_transpose_pad(cudaStream_t st, const double* X, int n, int d, double* XT, int n_pad);

// --- kb_chol.cu --------------------------------------------------------------
// sync_words: kb_chol_sync_words(g, B) ints (queue head, role tickets, row progress, PRE flags),
// zeroed by the launcher on the stream before every launch.
#ifdef __cplusplus
#include <vector>
size_t kb_chol_sync_words(const KbAugGeom& g, int B);
void kb_chol_build_tasks(const KbAugGeom& g, int B, int num_ctas, std::vector<int4>& out);
#endif
cudaError_t kb_launch_chol(cudaStream_t st, const KbAugGeom& g, int B, double* aug, double* dinv,
                           int* sync_words, const int4* tasks, int ntasks, int* info, int num_sms);

// --- kb_gls.cu ---------------------------------------------------------------
cudaError_t kb_launch_gls(cudaStream_t st, const KbAugGeom& g, int B, const double* aug, int trainvar,
                          const double* sigma2_in, double* Mbuf, double* beta, double* sigma2_out,
                          double* nll, int* info, double* rt_out /*[n_pad] or null (B==1)*/,
                          double* lndet_out);
// pivot spread of a fitted factor (min / max of diag L over the real rows -> out[0], out[1]) and its k-major copy
cudaError_t kb_launch_diag_spread(cudaStream_t st, const KbAugGeom& g, const double* aug, double* out2);
cudaError_t kb_launch_fit_refine_extra(cudaStream_t st, const KbAugGeom& g, const double* aug, const double* rt,
                                       double* AT, const double* LT, int ldat, double* res);
cudaError_t kb_launch_fit_extract_factor(cudaStream_t st, const KbAugGeom& g, const double* aug, double* LT, int ldat);
cudaError_t kb_launch_fit_extract(cudaStream_t st, const KbAugGeom& g, const double* aug, const double* rt,
                                  double* AT, int ldat);

cudaError_t kb_launch_loocv(cudaStream_t st, int n, int n_pad, int nind, const double* AT, int ldat,
                            const double* LM, const double* y, double* pred);

// --- kb_predict.cu -----------------------------------------------------------
// max / min of diag L beyond which a fit's sweeps use the residual-corrected product (calibration: below 250 the
// explicit-inverse product and a triangular solve agree to rounding; at ~1000, pivots at the nugget, they differ 5-30x)
#define KB_PRED_REFINE_SPREAD 300.0
struct KbPredictArgs {
  // model
  int n, d, nind, n_pad, kmask;
  const double* XT;        // [d][n_pad]
  const double* gw;        // [d]
  const double* ith;       // [d]
  const double* wk;        // [4]
  const double* AT;        // [n_pad][ldat]  k-major: [L^-T | alpha | R^-1 F | 0]
  int ldat, nrows;         // nrows = rows of the A operand (multiple of 128)
  const double* beta;      // [nind]
  const double* LM;        // [nind][nind] Cholesky factor of F' R^-1 F (lower)
  const int* idx;          // [nind][d] trend multi-index (null when nind==1 and order 0)
  int maxorder;
  double sigma2;
  // input normalisation: type 0: ((x-p0)/(p1-p0)-0.5)*2 ; type 1: (x-p0)/p1 ; type 2: none
  int norm_type;
  const double* p0;
  const double* p1;
  // points
  long long m;
  const double* x;         // [m][d] raw units (device)
  long long idx_offset;    // global index of x[0] (multi-GPU shards)
  // mean in real units when the responses were normalised (KB_YMAP_*, prediction.py:213-217)
  int ymode;
  double ya, yb;
  // acquisition parameters
  double ybest, limit, sigmalcb;
  int limit_sign;          // +1 for '>' '>=' ; -1 for '<' '<='
  int acq;                 // which output drives the arg-min reduction
  // outputs (device, each m doubles or null)
  double* out[KB_OUT_COUNT];
  // scratch
  double* psi_scratch;     // [grid][n_pad][strip]   (streaming variant only; strip = 128)
  double* ex_scratch;      // [grid][qx][strip]
  int stage_x;             // raw sites of the next strip prefetched into shared memory (narrow inputs)
  int qx;                  // extra rows (1 + nind)
  int resident;            // 1: inverse factor and psi strip stay in shared memory (n_pad <= 128);
                           // 2: warp-autonomous variant (n <= 64, d <= 8, ordinary Kriging)
  KbSweepResult* partials; // [grid]
  // residual-corrected product for ill-conditioned fits (kb_fit_state sets it from the pivot spread of the factor):
  //   v0 = W psi,  r = psi - L v0,  v = v0 + W r     (W = L^-1 explicit; three passes of the same tile stream)
  int refine;
  const double* LT;        // [n_pad][ldat]  k-major factor: LT[k][i] = L[i][k]  (refine only)
};
cudaError_t kb_launch_predict(cudaStream_t st, const KbPredictArgs& a, int grid, KbSweepResult* result_dev);
// few-sites path: slices of KB_FEW_MAX sites on a model that does not fit the resident variants
#define KB_FEW_MAX 16
#define KB_FEW_MAX_SLICES 96
struct KbFewScratch {
  double* psi;            // [slices][n_pad][KB_FEW_MAX]
  double* ex;             // [slices][n_pad/64][qx][KB_FEW_MAX]  per-block partial sums of the extra columns
  double* vv;             // [slices][n_pad/16][KB_FEW_MAX]      per-block partial sums of squares
  KbSweepResult* slices;  // [slices] reductions of each slice
  int* counter;           // [slices] finished blocks of the second kernel (reset by the last one)
};
int kb_predict_few_slices(long long m);
size_t kb_predict_few_scratch_doubles(int n_pad, int qx, int ny);
cudaError_t kb_launch_predict_few(cudaStream_t st, const KbPredictArgs& a, double* scratch, int ny_cap,
                                  KbSweepResult* result_dev);
size_t kb_predict_smem_bytes(int d, bool resident, int n_pad, int qx, int stage_x);
bool kb_predict_can_reside(int d, int n_pad, int qx, int nind_trend);
int kb_predict_strip(bool resident);
bool kb_predict_can_warp(int d, int n, int qx, int nind_trend);
size_t kb_predict_warp_smem_bytes(int d, int n8, int qx);
int kb_predict_warp_ctas_per_sm(int d, int n, int qx);
int kb_predict_max_grid(int kmask, int d, int num_sms, int resident, int n_pad, int qx, int stage_x);

// --- kb_rng.cu ---------------------------------------------------------------
cudaError_t kb_launch_points_generate(cudaStream_t st, double* x, long long npts, int d, int dist, double p0, double p1,
                                      const double* lb_dev, const double* ub_dev, unsigned long long seed,
                                      long long first_index);

cudaError_t kb_launch_points_lowdisc(cudaStream_t st, double* x, long long npts, int d, int kind,
                                     const unsigned int* dirs_dev, const int* primes_dev, const double* lb_dev,
                                     const double* ub_dev, long long first_index);

// kb_small.cu: one-tile designs (n <= 64, at most 16 right-hand-side rows) in a single launch
bool kb_small_applies(const KbAugGeom& g, int d, bool kpls);
cudaError_t kb_launch_nll_small(cudaStream_t st, const KbModelDev& m, const KbAugGeom& g, int B, int nbhyp,
                                const double* xpop_dev, int nth, int trainvar, int has_nugget, int has_weights,
                                double fixed_nugget, int nkernel, const int* kernel_ids_dev, KbHyper hp, double* aug,
                                int* info);

// kb_ehvi.cu: 2-objective expected hypervolume improvement (exi2d.py:6-83)
size_t kb_ehvi_table_doubles(int k);
void kb_ehvi_build_table(int k, const double* front, const double* ref, double* table);
int kb_ehvi_grid(long long m, int num_sms);
cudaError_t kb_launch_ehvi2d(cudaStream_t st, long long m, const double* mu1, const double* s1, const double* mu2,
                             const double* s2, const double* table_dev, int k, long long idx_offset, double* out,
                             void* partials, void* result, int accumulate, int num_sms);

// --- kb_peak.cu --------------------------------------------------------------
cudaError_t kb_launch_peak(cudaStream_t st, int which, int iters, int grid, double* sink);

// kb_cmaes.cu: device-resident CMA-ES generation loop
// state layout (doubles) -- see kb_cmaes_state_doubles
struct KbCmaView {
  double *mean, *pc, *ps, *C, *A, *D, *scale, *lb, *ub, *w, *best_x, *z, *y, *x, *hist;
  double* scal;   // [0] sigma [1] best_f [2] gen [3] stop [4] evals
};

__host__ __device__ inline size_t kb_cma_view(double* base, int N, int lam, int mu, int maxiter, KbCmaView* v) {
  size_t o = 0;
  auto take = [&](size_t n) { double* p = base ? base + o : nullptr; o += n; return p; };
  KbCmaView t;
  t.scal = take(8);
  t.mean = take(N); t.pc = take(N); t.ps = take(N); t.D = take(N); t.scale = take(N); t.lb = take(N); t.ub = take(N);
  t.best_x = take(N); t.w = take(mu);
  t.C = take((size_t)N * N); t.A = take((size_t)N * N);
  t.z = take((size_t)lam * N); t.y = take((size_t)lam * N); t.x = take((size_t)lam * N);
  t.hist = take(maxiter + 1);
  if (v) *v = t;
  return o;
}


struct KbCmaParams {
  int N, lam, mu, maxiter, nbhyp, has_bounds, window;
  double cc, cs, c1, cmu, damps, chiN, mueff, tolfun, tolx;
  unsigned long long seed;
};

size_t kb_cmaes_state_doubles(int N, int lam, int mu, int maxiter);
cudaError_t kb_launch_cmaes_step(cudaStream_t st, const KbCmaParams& P, double* state, const double* f_in, double* xpop,
                                 int first);
