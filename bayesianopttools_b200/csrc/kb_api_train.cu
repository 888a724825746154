// krig-b200: C-ABI entry points around training: the device-resident CMA-ES loop, closed-form LOOCV, the plain
// correlation matrix and the on-device peak measurements.
#include "kb_api_internal.h"

extern "C" {
// ---- device-resident CMA-ES --------------------------------------------------------------------------
kb_status kb_train_cmaes(kb_model* m, int nbhyp, int trainvar, double fixed_nugget_log10, const double* x0,
                         const double* lb, const double* ub, const double* scaling, const kb_cmaes_opts* opts,
                         double* xbest, double* fbest, int* generations, int* stop_reason, double* hist) {
  if (!m || !x0 || !opts || !xbest || nbhyp < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if ((lb == nullptr) != (ub == nullptr)) KB_FAIL(KB_ERR_INVALID, "give both bounds or neither");
  if (!(opts->sigma0 > 0.0)) KB_FAIL(KB_ERR_INVALID, "sigma0 must be positive");
  const int N = nbhyp;
  if (N > 64) KB_FAIL(KB_ERR_INVALID, "the device CMA-ES handles up to 64 hyper-parameters (got %d)", N);
  KbCmaParams P{};
  P.N = N; P.nbhyp = nbhyp;
  P.lam = opts->popsize > 0 ? opts->popsize : 4 + (int)(3.0 * std::log((double)N));
  if (P.lam < 4 || P.lam > 256) KB_FAIL(KB_ERR_INVALID, "population size %d outside [4, 256]", P.lam);
  P.mu = P.lam / 2;
  std::vector<double> w((size_t)P.mu);
  double wsum = 0.0, w2 = 0.0;
  for (int i = 0; i < P.mu; ++i) { w[i] = std::log(P.mu + 0.5) - std::log((double)(i + 1)); wsum += w[i]; }
  for (int i = 0; i < P.mu; ++i) { w[i] /= wsum; w2 += w[i] * w[i]; }
  P.mueff = 1.0 / w2;
  P.cc = (4.0 + P.mueff / N) / (N + 4.0 + 2.0 * P.mueff / N);
  P.cs = (P.mueff + 2.0) / (N + P.mueff + 5.0);
  P.c1 = 2.0 / ((N + 1.3) * (N + 1.3) + P.mueff);
  P.cmu = std::min(1.0 - P.c1, 2.0 * (P.mueff - 2.0 + 1.0 / P.mueff) / ((N + 2.0) * (N + 2.0) + P.mueff));
  P.damps = 1.0 + 2.0 * std::max(0.0, std::sqrt((P.mueff - 1.0) / (N + 1.0)) - 1.0) + P.cs;
  P.chiN = std::sqrt((double)N) * (1.0 - 1.0 / (4.0 * N) + 1.0 / (21.0 * N * N));
  P.maxiter = opts->maxiter > 0 ? opts->maxiter : 100 + 150 * (N + 3) * (N + 3) / (int)std::sqrt((double)P.lam);
  P.tolfun = opts->tolfun > 0.0 ? opts->tolfun : 1e-11;
  P.tolx = opts->tolx > 0.0 ? opts->tolx : 1e-11;
  P.window = 10 + (int)(30.0 * N / P.lam);
  P.has_bounds = lb ? 1 : 0;
  P.seed = opts->seed;
  const int check_every = opts->check_every > 0 ? opts->check_every : 4;

  CK(cudaSetDevice(m->device));
  kb_status s = ensure_ws(m, m->pop, P.lam, false);
  if (s != KB_OK) return s;
  const size_t nstate = kb_cmaes_state_doubles(N, P.lam, P.mu, P.maxiter);
  std::vector<double> init(nstate, 0.0);
  KbCmaView V;
  kb_cma_view(init.data(), N, P.lam, P.mu, P.maxiter, &V);
  V.scal[0] = opts->sigma0;
  V.scal[1] = INFINITY;
  for (int a = 0; a < N; ++a) {
    V.mean[a] = x0[a]; V.best_x[a] = x0[a]; V.D[a] = 1.0;
    V.scale[a] = scaling ? scaling[a] : 1.0;
    V.lb[a] = lb ? lb[a] : 0.0; V.ub[a] = ub ? ub[a] : 0.0;
    V.C[(size_t)a * N + a] = 1.0; V.A[(size_t)a * N + a] = 1.0;
  }
  for (int i = 0; i < P.mu; ++i) V.w[i] = w[i];
  double *state = nullptr, *xpop = nullptr, *fdev = nullptr;
  cudaError_t e = kb_dev_malloc(&state, sizeof(double) * nstate);
  if (e == cudaSuccess) e = kb_dev_malloc(&xpop, sizeof(double) * (size_t)P.lam * nbhyp);
  if (e == cudaSuccess) e = kb_dev_malloc(&fdev, sizeof(double) * (size_t)P.lam);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(state, init.data(), sizeof(double) * nstate, cudaMemcpyHostToDevice, m->stream);
  if (e == cudaSuccess) e = kb_launch_cmaes_step(m->stream, P, state, fdev, xpop, 1);
  g_launches += 1;
  double scal[8] = {0};
  while (e == cudaSuccess && s == KB_OK) {
    for (int it = 0; it < check_every && e == cudaSuccess && s == KB_OK; ++it) {
      s = run_pipeline(m, m->pop, P.lam, nbhyp, xpop, trainvar, fixed_nugget_log10, fdev, nullptr);
      if (s == KB_OK) e = kb_launch_cmaes_step(m->stream, P, state, fdev, xpop, 0);
      g_launches += 1;
    }
    if (e == cudaSuccess && s == KB_OK)
      e = cudaMemcpyAsync(scal, state, sizeof(scal), cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (scal[3] != 0.0) break;
  }
  if (e == cudaSuccess && s == KB_OK) {
    std::vector<double> fin(nstate);
    e = cudaMemcpy(fin.data(), state, sizeof(double) * nstate, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
      kb_cma_view(fin.data(), N, P.lam, P.mu, P.maxiter, &V);
      for (int a = 0; a < N; ++a) xbest[a] = V.best_x[a];
      if (fbest) *fbest = V.scal[1];
      if (generations) *generations = (int)V.scal[2];
      if (stop_reason) *stop_reason = (int)V.scal[3];
      if (hist) for (int g = 0; g < (int)V.scal[2]; ++g) hist[g] = V.hist[g];
    }
  }
  kb_dev_free(state); kb_dev_free(xpop); kb_dev_free(fdev);
  if (s != KB_OK) return s;
  if (e != cudaSuccess) KB_FAIL(KB_ERR_CUDA, "cmaes: %s", cudaGetErrorString(e));
  return KB_OK;
}

kb_status kb_loocv(kb_model* m, double* pred_out) {
  if (!m || !pred_out) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (!m->fitted) KB_FAIL(KB_ERR_STATE, "loocv called before kb_fit_state");
  CK(cudaSetDevice(m->device));
  if (!m->dense_tmp) CK(kb_dev_malloc(&m->dense_tmp, sizeof(double) * (size_t)m->n * m->n));
  CK(kb_launch_loocv(m->stream, m->n, m->n_pad, m->nind, m->AT, m->ldat, m->fit.Mbuf, m->y, m->dense_tmp));
  g_launches += 1;
  CK(cudaMemcpyAsync(pred_out, m->dense_tmp, sizeof(double) * m->n, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return KB_OK;
}

kb_status kb_corr_matrix(int device, int n, int mm, int d, const double* XN, const double* XM,
                         const double* theta, int kernel_id, int h, const double* plscoeff, double* out) {
  if (!XN || !XM || !theta || !out || n < 1 || mm < 1 || d < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (kernel_id < 0 || kernel_id > 3) KB_FAIL(KB_ERR_INVALID, "Kernel option not available!");
  if (d > KB_MAX_D) KB_FAIL(KB_ERR_INVALID, "d=%d exceeds KB_MAX_D=%d", d, KB_MAX_D);
  if (kb_device_count() == 0) KB_FAIL(KB_ERR_CUDA, "no CUDA device visible: no CPU fallback");
  CK(cudaSetDevice(device));
  const int ldn = round_up(n, KB_TILE), ldm = round_up(mm, KB_TILE);
  std::vector<double> gw(d), ith(d), wk(4, 0.0);
  wk[kernel_id] = 1.0;
  for (int i = 0; i < d; ++i) {
    if (plscoeff && h > 0) {
      double s = 0.0;
      for (int k = 0; k < h; ++k) s += plscoeff[i * h + k] * plscoeff[i * h + k] / (theta[k] * theta[k]);
      gw[i] = 0.5 * s; ith[i] = 0.0;
    } else {
      gw[i] = 0.5 / (theta[i] * theta[i]); ith[i] = 1.0 / theta[i];
    }
  }
  double *dXN = nullptr, *dXM = nullptr, *dXNT = nullptr, *dXMT = nullptr, *dpar = nullptr, *dout = nullptr;
  cudaStream_t st = nullptr;
  kb_status rc = KB_OK;
  auto cleanup = [&]() {
    kb_dev_free(dXN); kb_dev_free(dXM); kb_dev_free(dXNT); kb_dev_free(dXMT); kb_dev_free(dpar); kb_dev_free(dout);
  };
#define CKC(call)                                                                          \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      char buf_[512];                                                                      \
      snprintf(buf_, sizeof(buf_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      g_err = buf_; cleanup(); return KB_ERR_CUDA;                                         \
    }                                                                                      \
  } while (0)
  CKC(kb_dev_malloc(&dXN, sizeof(double) * (size_t)n * d));
  CKC(kb_dev_malloc(&dXM, sizeof(double) * (size_t)mm * d));
  CKC(kb_dev_malloc(&dXNT, sizeof(double) * (size_t)ldn * d));
  CKC(kb_dev_malloc(&dXMT, sizeof(double) * (size_t)ldm * d));
  CKC(kb_dev_malloc(&dpar, sizeof(double) * (2 * d + 4)));
  CKC(kb_dev_malloc(&dout, sizeof(double) * (size_t)n * mm));
  CKC(cudaMemcpy(dXN, XN, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dXM, XM, sizeof(double) * (size_t)mm * d, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dpar, gw.data(), sizeof(double) * d, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dpar + d, ith.data(), sizeof(double) * d, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dpar + 2 * d, wk.data(), sizeof(double) * 4, cudaMemcpyHostToDevice));
  CKC(kb_launch_transpose_pad(st, dXN, n, d, dXNT, ldn));
  CKC(kb_launch_transpose_pad(st, dXM, mm, d, dXMT, ldm));
  CKC(kb_launch_corr_general(st, n, mm, d, dXNT, ldn, dXMT, ldm, dpar, dpar + d, dpar + 2 * d,
                             1 << kernel_id, dout, mm));
  g_launches += 3;
  CKC(cudaMemcpy(out, dout, sizeof(double) * (size_t)n * mm, cudaMemcpyDeviceToHost));
#undef CKC
  cleanup();
  return rc;
}

kb_status kb_measure_peak(int device, int which, double* value) {
  if (!value || which < 0 || which > 2) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (kb_device_count() == 0) KB_FAIL(KB_ERR_CUDA, "no CUDA device visible");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms = 0.f;
  if (which < 2) {
    double* sink = nullptr;
    CK(kb_dev_malloc(&sink, sizeof(double)));
    const int grid = prop.multiProcessorCount * 4, iters = 20000;
    CK(kb_launch_peak(nullptr, which, 2000, grid, sink));
    CK(cudaDeviceSynchronize());
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      CK(kb_launch_peak(nullptr, which, iters, grid, sink));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double per_thread_block = (which == 0) ? (16.0 * 8.0 * 512.0) : (512.0 * 8.0 * 2.0);
      const double tf = per_thread_block * iters * grid / (ms * 1e-3) / 1e12;
      if (tf > best) best = tf;
    }
    g_launches += 4;
    kb_dev_free(sink);
    *value = best;
  } else {
    const size_t bytes = (size_t)1 << 30;
    double *a = nullptr, *b = nullptr;
    CK(kb_dev_malloc(&a, bytes));
    CK(kb_dev_malloc(&b, bytes));
    CK(cudaMemset(a, 0, bytes));
    CK(cudaMemcpy(b, a, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaDeviceSynchronize());
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, nullptr));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double gbs = 2.0 * bytes / (ms * 1e-3) / 1e9;
      if (gbs > best) best = gbs;
    }
    kb_dev_free(a); kb_dev_free(b);
    *value = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return KB_OK;
}

}  // extern "C"
