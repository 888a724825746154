/* krigb200.h -- C-ABI of libkrigb200.so: the B200 (sm_100a) Kriging hot path.
 *
 * The reference (fazaghifari/BayesianOptTools, "KADAL") is pure Python and has
 * no FFI; its operator surface for this path is the set of Python callables
 * cited at each entry point below.  A maintainer binds these symbols with
 * ctypes.  Conventions:
 *   - plain C, int status return (KB_OK == 0), no exceptions cross the boundary;
 *     kb_last_error() gives the message of the last failure on this thread
 *   - the caller owns every host buffer; the library owns all device memory
 *     behind the opaque kb_model handle
 *   - one handle is used from one host thread at a time; one CUDA stream per
 *     handle (kb_model_set_stream); entry points named *_dev take DEVICE
 *     pointers, enqueue on the handle's stream and do not synchronise
 *   - all matrices are row-major FP64 (numpy default)
 */
#ifndef KRIGB200_H
#define KRIGB200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct kb_model kb_model;
typedef int kb_status;

enum {
  KB_OK = 0,
  KB_ERR_INVALID = 1,   /* bad argument */
  KB_ERR_CUDA = 2,      /* CUDA runtime error (message in kb_last_error) */
  KB_ERR_STATE = 3,     /* e.g. predict before fit */
  KB_ERR_NOT_PD = 4     /* correlation matrix not positive definite (fit only) */
};

/* kernel ids: surrogate_models/supports/kernelfunc.py:22-33 */
enum { KB_GAUSSIAN = 0, KB_EXPONENTIAL = 1, KB_MATERN32 = 2, KB_MATERN52 = 3 };
/* model types: KrigInfo["type"], likelihood_func.py:93,98 */
enum { KB_TYPE_KRIGING = 0, KB_TYPE_KPLS = 1 };
/* input normalisation of prediction sites: prediction.py:155-160 */
enum { KB_NORM_DEFAULT = 0, KB_NORM_STD = 1, KB_NORM_NONE = 2 };

/* predictor outputs: prediction.py:238-285; U = |pred|/s is akmcs.py:223 */
enum {
  KB_OUT_PRED = 0, KB_OUT_SSQR = 1, KB_OUT_S = 2, KB_OUT_EI = 3, KB_OUT_POI = 4,
  KB_OUT_POF = 5, KB_OUT_LCB = 6, KB_OUT_EBE = 7, KB_OUT_U = 8, KB_OUT_FPC = 9,
  KB_OUT_COUNT = 10
};

typedef struct {
  double ybest;      /* min(y) of the training responses, prediction.py:249,272 */
  double limit;      /* KrigInfo["limit"], prediction.py:275-285 */
  int limit_sign;    /* +1 for '>' '>=', -1 for '<' '<=' */
  double sigmalcb;   /* KrigInfo["sigmalcb"], prediction.py:245 */
} kb_acq_params;

/* Reduce-only results of a sweep over a point set (akmcs.py:208-238, np.argmin
 * over the acquisition in acquifunc_opt.py:74-76).  Indices are GLOBAL point
 * indices (idx_offset + local); ties resolve to the lowest index (np.argmin). */
typedef struct {
  double best_val;          /* min of the selected output */
  long long best_idx;
  double min_u;             /* min |pred|/s */
  long long min_u_idx;
  long long n_fail;         /* #{pred <= 0}              akmcs.py:209 */
  long long n_fail_minus;   /* #{pred - 1.96 s <= 0}     akmcs.py:230,232 */
  long long n_fail_plus;    /* #{pred + 1.96 s <= 0}     akmcs.py:231,233 */
  long long n_points;
} KbSweepResult;

int kb_version(void);
const char* kb_last_error(void);
int kb_device_count(void);

/* Create a model for one design.  Replaces the state set up by
 * Kriging.__init__/standardize (kriging_model.py:54-167) and KPLS.standardize
 * (kpls_model.py:93-108): Xnorm (n x d), y (n), F (n x nind) are the
 * standardised design, responses and regression matrix; kernel_ids[nkernel] the
 * composite kernel list; for KB_TYPE_KPLS plscoeff is (d x h). */
kb_status kb_model_create(kb_model** out, int device, int n, int d, int nind, const double* Xnorm,
                          const double* y, const double* F, int nkernel, const int* kernel_ids,
                          int model_type, int h, const double* plscoeff);
void kb_model_destroy(kb_model* m);
kb_status kb_model_set_stream(kb_model* m, void* cuda_stream);
/* Back to the stream the handle created for itself (the default after kb_model_create). */
kb_status kb_model_use_own_stream(kb_model* m);
kb_status kb_model_sync(kb_model* m);

/* Population-batched negative ln-likelihood.  Replaces `likelihood(x, KrigInfo,
 * 'default', trainvar)` (likelihood_func.py:6-118, maincalc :168-213) called once
 * per candidate by the optimisers (kriging_model.py:420-446, uncGA.py:46-136).
 * xpop is (B x nbhyp) log10 values laid out as nuggetset decodes them
 * (likelihood_func.py:121-165): [theta | sigma2? | nugget? | weights?].
 * info_out[b] = 0, or k>0 when pivot k of the Cholesky was not positive (then
 * nll_out[b] = +inf; the reference raises LinAlgError at likelihood_func.py:175). */
kb_status kb_nll_batch(kb_model* m, int B, int nbhyp, const double* xpop, int trainvar,
                       double fixed_nugget_log10, double* nll_out, int* info_out);
kb_status kb_nll_batch_dev(kb_model* m, int B, int nbhyp, const double* xpop_dev, int trainvar,
                           double fixed_nugget_log10, double* nll_out_dev, int* info_out_dev);
/* Optional per-candidate by-products of the last kb_nll_batch call (host copies). */
kb_status kb_nll_batch_extras(kb_model* m, int B, double* beta_out /*B x nind*/, double* sigma2_out /*B*/);
/* Size the population workspace once (avoids allocation inside a timed region). */
kb_status kb_reserve_population(kb_model* m, int B);

/* `likelihood(x, KrigInfo, mode='all')` (likelihood_func.py:107-116): factorise
 * at one hyper-parameter vector and keep the predictor state on the device.
 * Any of U (n x n upper, = L^T), Psi (n x n, WITHOUT nugget, :108), BE (nind),
 * sigma2, nll, theta_out (d-vector of length scales actually used), may be NULL. */
kb_status kb_fit_state(kb_model* m, const double* x, int nbhyp, int trainvar, double fixed_nugget_log10,
                       double* U, double* Psi, double* BE, double* sigma2, double* nll, double* nugget_out,
                       double* wgkf_out /*nkernel*/);

/* Trend multi-index for prediction sites (trendfunction.py:7-41,70-107):
 * idx (nind x d) ints, Legendre basis on [-1,1].  Needed when nind > 1. */
kb_status kb_model_set_trend(kb_model* m, const int* idx, int nind, int d);

/* Input normalisation of prediction sites: KB_NORM_DEFAULT p0=lb,p1=ub
 * (samplingplan.py:84-88); KB_NORM_STD p0=X_mean,p1=X_std (prediction.py:160). */
kb_status kb_model_set_norm(kb_model* m, int norm_type, const double* p0, const double* p1);

/* Accuracy of the sweeps on ILL-CONDITIONED fits.  prediction.py:222-224 obtains v = L^-1 psi by a triangular solve; the
 * sweeps here multiply by the explicit inverse, which is one GEMM but, once the pivots of the factor reach down to the
 * nugget (dense low-dimensional designs, smooth kernels), 5-30x further from an exact evaluation than the solve.
 * kb_fit_state measures max / min of diag L and flags such fits (spread > 300); their sweeps then use the corrected
 * paths (blocked forward substitution in the small-model kernel, one residual step per strip otherwise: 3x the tensor
 * work of the strip kernel, 0-20 % on the small-model kernel).
 *   KB_REFINE_AUTO (default)  corrected paths for flagged fits
 *   KB_REFINE_OFF             always the plain product
 *   KB_REFINE_ON              always the corrected paths
 * Takes effect at the next kb_fit_state.  kb_model_predict_refined reports what the last fit chose (0 / 1). */
enum { KB_REFINE_AUTO = -1, KB_REFINE_OFF = 0, KB_REFINE_ON = 1 };
kb_status kb_model_set_predict_refine(kb_model* m, int mode);
int kb_model_predict_refined(const kb_model* m);

/* Output map of the predicted mean when the model was trained on normalised responses (KrigInfo["norm_y"],
 * prediction.py:213-217 -> stdtoreal :300-313): applied to f BEFORE every acquisition function, U and the AK-MCS
 * counters, exactly where the reference applies it; variance and s stay in the model's units, as in the reference.
 *   KB_YMAP_NONE                                  f
 *   KB_YMAP_DEFAULT  a = max(y) - min(y), b = min(y)   (f / 2 + 0.5) * a + b
 *   KB_YMAP_STD      a = y_std, b = y_mean             b + a * f                                      */
enum { KB_YMAP_NONE = 0, KB_YMAP_DEFAULT = 1, KB_YMAP_STD = 2 };
kb_status kb_model_set_ymap(kb_model* m, int mode, double a, double b);

/* Batched predictor.  Replaces `prediction(x, KrigInfo, predtypes)`
 * (prediction.py:67-298): x is (m x d) in RAW units; outs[k] (k = KB_OUT_*) is a
 * buffer of m doubles or NULL.  result (may be NULL) receives the reductions
 * with `acq` (a KB_OUT_* id) selecting the arg-min output. */
kb_status kb_predict(kb_model* m, long long npts, const double* x, const kb_acq_params* acq_params,
                     double* const* outs, int acq, KbSweepResult* result);
kb_status kb_predict_dev(kb_model* m, long long npts, const double* x_dev, long long idx_offset,
                         const kb_acq_params* acq_params, double* const* outs_dev, int acq,
                         KbSweepResult* result_dev);

/* Device-resident point sets.  AK-MCS sweeps the SAME Monte-Carlo population after every model
 * update (akmcs.py:66-206: `init_samp` is fixed, only the surrogate changes), so the population is
 * uploaded once and every sweep is a single launch with a 64-byte result.  A point set belongs to a
 * device, not to a model: it survives the re-creation of the model handle when the design grows. */
typedef struct kb_points kb_points;
kb_status kb_points_create(kb_points** out, int device, long long npts, int d, const double* x /*npts x d raw*/);
void kb_points_destroy(kb_points* p);
/* Monte-Carlo population generated ON the device (mcpopgen, akmcs.py:241-260) with the counter-based
 * Philox4x32-10 generator: the two uniforms of site i, dimension pair j are Philox(counter = (i, j),
 * key = seed), so any subset can be reproduced on the host (oracle/philox.py) and shards of a
 * multi-GPU population are generated independently (first_index = global index of site 0).
 * dist: KB_DIST_NORMAL (p0 = mean, p1 = stddev), KB_DIST_LOGNORMAL (mean, stddev of the variable),
 * KB_DIST_GUMBEL (mean, stddev), KB_DIST_UNIFORM (per-dimension lb, ub). */
enum { KB_DIST_NORMAL = 0, KB_DIST_LOGNORMAL = 1, KB_DIST_GUMBEL = 2, KB_DIST_UNIFORM = 3, KB_DIST_HALTON = 4,
       KB_DIST_SOBOL = 5 };
kb_status kb_points_generate(kb_points** out, int device, long long npts, int d, int dist, double p0, double p1,
                             const double* lb, const double* ub, unsigned long long seed, long long first_index);
/* Low-discrepancy designs generated on the device, bit-identical to the reference's host samplers
 * (`sampling('halton' | 'sobolnew', ...)`, misc/sampling/haltonsampling.py:30-40, sobol_new.py:5-98) and
 * mapped to [lb, ub] as `realval` does (lb = ub = NULL: unit cube).  kind = KB_DIST_HALTON (dirs unused)
 * or KB_DIST_SOBOL with dirs = the 32 direction numbers v_1..v_32 (as 32-bit integers, v_k = m_k << (32-k))
 * of each of the d dimensions, d x 32 row-major; site i is the sequence's point first_index + i
 * (point 0 is the origin). */
kb_status kb_points_generate_lowdisc(kb_points** out, int device, long long npts, int d, int kind,
                                     const unsigned int* dirs, const double* lb, const double* ub,
                                     long long first_index);
/* Copy rows [row0, row0 + nrows) of a point set to the host (e.g. the site picked by the arg-min). */
kb_status kb_points_get_rows(const kb_points* p, long long row0, long long nrows, double* x_out);
/* Same outputs/reductions as kb_predict over a resident point set; outs[k] are HOST buffers of
 * npts doubles or NULL (reduce-only when all are NULL). */
kb_status kb_predict_points(kb_model* m, const kb_points* pts, const kb_acq_params* acq_params,
                            double* const* outs, int acq, KbSweepResult* result);

/* 2-objective expected hypervolume improvement (multi-objective acquisition).
 * kb_exi2d is `exi2d(P, r, mu, s)` (optim_tools/ehvi/exi2d.py:6-83) over npts candidates at once:
 * front is the k x 2 row-major set of non-dominated objective values P, ref the reference point r,
 * mu and s are (npts x 2) row-major HOST arrays, out[npts] the improvement of each candidate.
 * kb_ehvi2d is the batched `ehvicalc` (optim_tools/ehvi/EHVIcomputation.py:10-42): both fitted models
 * predict the npts raw candidates x (npts x d, HOST) and the improvements are computed from the
 * predictions without leaving the device.  spread = 0 hands the predictive VARIANCE to exi2d as its
 * `s` argument, which is what ehvicalc does (EHVIcomputation.py:30-35 fills `s` from 'SSqr');
 * spread = 1 hands the standard deviation.  out (npts, HOST) may be NULL; best_idx/best_val receive
 * the candidate with the LARGEST improvement (ties: lowest index).  The two models must live on the
 * same device and share d. */
kb_status kb_exi2d(int device, long long npts, const double* mu, const double* s, int k, const double* front,
                   const double* ref, double* out);
kb_status kb_ehvi2d(kb_model* m1, kb_model* m2, long long npts, const double* x, int k, const double* front,
                    const double* ref, int spread, double* out, long long* best_idx, double* best_val);
/* kb_ehvi2d over a device-resident point set (no host->device copy of the candidates). */
kb_status kb_ehvi2d_points(kb_model* m1, kb_model* m2, const kb_points* pts, int k, const double* front,
                           const double* ref, int spread, double* out, long long* best_idx, double* best_val);

/* Hyper-parameter search by CMA-ES with the whole generation loop on the device (replaces the host
 * loop around `cma.fmin2(likelihood, xhyp, sigmacmaes, {'bounds': ..., 'scaling_of_variables': ...})`,
 * kriging_model.py:420-424).  Per generation the device ranks the population, updates mean / evolution
 * paths / covariance / step size, eigen-decomposes the covariance (parallel Jacobi), tests the stop
 * criteria and samples the next population from the counter-based Philox stream; the likelihood pipeline
 * then evaluates it.  The host enqueues `check_every` generations at a time and reads back a 40-byte
 * status.  x0, lb, ub, scaling: nbhyp values each (lb = ub = NULL: unbounded; scaling NULL: ones).
 * Zero-valued options take the defaults of optim_tools/cmaes.py (popsize 4 + floor(3 ln N), maxiter
 * 100 + 150 (N+3)^2 / sqrt(popsize), tolfun = tolx = 1e-11, check_every 4).  Limits: nbhyp <= 64,
 * popsize <= 256.  hist (optional, maxiter doubles) receives the best penalised objective of every
 * generation; stop_reason: 1 maxiter, 2 tolx, 3 tolfun (flat history), 4 ill-conditioned covariance. */
typedef struct {
  int popsize, maxiter, check_every;
  double sigma0, tolfun, tolx;
  unsigned long long seed;
} kb_cmaes_opts;
kb_status kb_train_cmaes(kb_model* m, int nbhyp, int trainvar, double fixed_nugget_log10, const double* x0,
                         const double* lb, const double* ub, const double* scaling, const kb_cmaes_opts* opts,
                         double* xbest, double* fbest, int* generations, int* stop_reason, double* hist);

/* Leave-one-out predictions at the fitted hyper-parameters, pred_out[n].  Replaces the n
 * refits + n single-site predictions of loocv2 (krigloocv2.py:7-33) with the closed form
 * from one factorisation; same numbers up to rounding. */
kb_status kb_loocv(kb_model* m, double* pred_out);

/* `calckernel(XN, XM, theta, nvar, type=, plscoeff=)` (kernelfunc.py:4-34):
 * out is (n x m).  For KPLS pass theta of length h and plscoeff (d x h). */
kb_status kb_corr_matrix(int device, int n, int m, int d, const double* XN, const double* XM,
                         const double* theta, int kernel_id, int h, const double* plscoeff, double* out);

/* Roofline denominators measured on this device: which = 0 DMMA (mma.sync f64),
 * 1 scalar DFMA, 2 device-to-device copy.  Returns TFLOP/s (0,1) or GB/s (2). */
kb_status kb_measure_peak(int device, int which, double* value);

/* Per-stage device timing, measured with CUDA events on the handle's stream:
 * ms5 = {decode, correlation build, Cholesky, GLS reductions, predict}; `back` = 0 is the most
 * recent population call / predict launch, up to 31 calls back (-1 where no such call).
 * Enable with kb_set_profiling(m, 1): it adds event records only, no synchronisation. */
kb_status kb_set_profiling(kb_model* m, int on);
kb_status kb_get_stage_ms(kb_model* m, int back, float* ms5);

/* Kernel launches issued by this library since process start (bench "gpu_launches"). */
long long kb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* KRIGB200_H */
