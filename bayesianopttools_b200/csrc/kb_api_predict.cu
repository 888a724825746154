// krig-b200: C-ABI entry points of the predictor: fused predict / acquisition sweeps over host or device point
// sets (prediction.py:67-298, akmcs.py:208-238), device-resident point sets and the 2-objective EHVI sweep.
#include "kb_api_internal.h"

extern "C" {
static kb_status ensure_predict_scratch(kb_model* m) {
  const int qx = 1 + m->nind;
  // Ill-conditioned fits (kb_fit_state): designs the warp-autonomous kernel serves (n <= 128, ordinary Kriging: the
  // AK-MCS population sweeps) take its blocked-substitution builds, everything else the strip kernel with the
  // residual-corrected product (never the few-sites or the resident CTA kernels, which multiply by the explicit inverse).
  const bool warp_ok = kb_predict_can_warp(m->d, m->n, qx, m->idx_dev ? 1 : 0);
  const int refine = m->pred_refine ? 1 : 0;
  const int resident = warp_ok ? 2 : (refine ? 0 : (kb_predict_can_reside(m->d, m->n_pad, qx, m->idx_dev ? 1 : 0) ? 1 : 0));
  if (m->ex_scratch && resident == m->pred_resident && m->pred_refine_scratch == refine) return KB_OK;
  CK(cudaStreamSynchronize(m->stream));
  m->pred_refine_scratch = refine;
  kb_dev_free(m->psi_scratch); kb_dev_free(m->ex_scratch); kb_dev_free(m->partials);
  m->psi_scratch = m->ex_scratch = nullptr; m->partials = nullptr;
  m->pred_resident = resident;
  m->pred_stage_x = 1;
  if (resident == 2) {
    m->pred_grid = kb_predict_warp_ctas_per_sm(m->d, m->n, qx) * m->num_sms;
  } else {
    // prefetching the next strip's raw sites costs 2 x 64 x d doubles of shared memory: keep it only
    // while it does not cost a resident CTA
    const int g1 = kb_predict_max_grid(m->kmask, m->d, m->num_sms, resident, m->n_pad, qx, 1);
    const int g0 = kb_predict_max_grid(m->kmask, m->d, m->num_sms, resident, m->n_pad, qx, 0);
    m->pred_stage_x = (g1 >= g0 && g1 > 0) ? 1 : 0;
    static const char* sx_env = getenv("KB_PRED_STAGE_X");            // debug aid: force 0 / 1 when it fits
    if (sx_env && atoi(sx_env) == 1 && g1 > 0) m->pred_stage_x = 1;
    if (sx_env && atoi(sx_env) == 0 && g0 > 0) m->pred_stage_x = 0;
    m->pred_grid = m->pred_stage_x ? g1 : g0;
  }
  if (m->pred_grid < 1)
    KB_FAIL(KB_ERR_CUDA, "the predictor keeps one 64-site strip and a 16-site training tile per dimension in shared "
                         "memory: d = %d does not fit (limit about 180)", m->d);
  const size_t strip = kb_predict_strip(resident != 0);
  // (refine: a second strip per CTA holds v0)
  if (!resident)
    CK(kb_dev_malloc(&m->psi_scratch, sizeof(double) * (size_t)m->pred_grid * m->n_pad * strip * (refine ? 2 : 1)));
  CK(kb_dev_malloc(&m->ex_scratch, sizeof(double) * (size_t)m->pred_grid * qx * strip));
  CK(kb_dev_malloc(&m->partials, sizeof(KbSweepResult) * m->pred_grid));
  return KB_OK;
}

kb_status kb_predict_dev(kb_model* m, long long npts, const double* x_dev, long long idx_offset,
                         const kb_acq_params* ap, double* const* outs_dev, int acq, KbSweepResult* result_dev) {
  if (!m || !x_dev || npts < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (!m->fitted) KB_FAIL(KB_ERR_STATE, "predict called before kb_fit_state");
  if (m->nind > 1 && !m->idx_dev) KB_FAIL(KB_ERR_STATE, "nind > 1 needs kb_model_set_trend");
  if (acq < 0 || acq >= KB_OUT_COUNT) KB_FAIL(KB_ERR_INVALID, "bad acquisition id");
  CK(cudaSetDevice(m->device));
  kb_status s = ensure_predict_scratch(m);
  if (s != KB_OK) return s;
  KbPredictArgs a{};
  a.n = m->n; a.d = m->d; a.nind = m->nind; a.n_pad = m->n_pad; a.kmask = m->kmask;
  a.XT = m->XT; a.gw = m->fit.hp.gw; a.ith = m->fit.hp.ith; a.wk = m->fit.hp.wk;
  a.AT = m->AT; a.ldat = m->ldat; a.nrows = m->nrows;
  a.beta = m->fit.beta; a.LM = m->fit.Mbuf; a.idx = m->idx_dev; a.maxorder = m->maxorder;
  a.sigma2 = m->sigma2;
  a.norm_type = m->norm_type; a.p0 = m->p0; a.p1 = m->p1;
  a.ymode = m->ymode; a.ya = m->ya; a.yb = m->yb;
  a.m = npts; a.x = x_dev; a.idx_offset = idx_offset;
  a.ybest = ap ? ap->ybest : 0.0;
  a.limit = ap ? ap->limit : 0.0;
  a.limit_sign = ap ? (ap->limit_sign >= 0 ? 1 : -1) : 1;
  a.sigmalcb = ap ? ap->sigmalcb : 0.0;
  a.acq = acq;
  for (int k = 0; k < KB_OUT_COUNT; ++k) a.out[k] = outs_dev ? outs_dev[k] : nullptr;
  a.psi_scratch = m->psi_scratch; a.ex_scratch = m->ex_scratch; a.qx = 1 + m->nind; a.partials = m->partials;
  a.resident = m->pred_resident;
  a.stage_x = m->pred_stage_x;
  a.refine = m->pred_refine_scratch;        // (decided in ensure_predict_scratch)
  a.LT = m->LT;
  // a handful of sites on a large model: spread the single strip over n_pad / 64 CTAs
  static const char* few_env = getenv("KB_PREDICT_FEW");            // debug aid: 0 disables
  const int few_ny = npts <= (long long)KB_FEW_MAX * KB_FEW_MAX_SLICES ? kb_predict_few_slices(npts) : 0;
  // the scratch grows with the slice count; cap it at 256 MB (many regressors on a large design)
  const bool few_fits = few_ny > 0 &&
                        kb_predict_few_scratch_doubles(m->n_pad, a.qx, few_ny) * sizeof(double) <= (256u << 20);
  if (few_fits && m->pred_resident == 0 && !a.refine && !(few_env && atoi(few_env) == 0)) {
    if (few_ny > m->few_slices) {
      CK(cudaStreamSynchronize(m->stream));
      kb_dev_free(m->few_scratch);
      m->few_scratch = nullptr;
      int cap = m->few_slices > 0 ? m->few_slices : 1;
      while (cap < few_ny) cap *= 2;
      while (cap > few_ny && kb_predict_few_scratch_doubles(m->n_pad, a.qx, cap) * sizeof(double) > (256u << 20)) --cap;
      const size_t nd = kb_predict_few_scratch_doubles(m->n_pad, a.qx, cap);
      CK(kb_dev_malloc(&m->few_scratch, sizeof(double) * nd));
      CK(cudaMemsetAsync(m->few_scratch, 0, sizeof(double) * nd, m->stream));
      m->few_slices = cap;
    }
    if (m->profile) CK(cudaEventRecord(m->evp[m->pred_calls % KB_PROF_RING][0], m->stream));
    CK(kb_launch_predict_few(m->stream, a, m->few_scratch, m->few_slices, result_dev));
    if (m->profile) { CK(cudaEventRecord(m->evp[m->pred_calls % KB_PROF_RING][1], m->stream)); m->pred_calls += 1; }
    g_launches += (result_dev && few_ny > 1) ? 3 : 2;
    return KB_OK;
  }
  const long long strip = (m->pred_resident == 2) ? 256 : kb_predict_strip(m->pred_resident != 0);
  const long long nstrips = (npts + strip - 1) / strip;
  const int grid = (int)(nstrips < m->pred_grid ? nstrips : m->pred_grid);
  if (m->profile) CK(cudaEventRecord(m->evp[m->pred_calls % KB_PROF_RING][0], m->stream));
  CK(kb_launch_predict(m->stream, a, grid, result_dev));
  if (m->profile) { CK(cudaEventRecord(m->evp[m->pred_calls % KB_PROF_RING][1], m->stream)); m->pred_calls += 1; }
  g_launches += result_dev ? 2 : 1;
  return KB_OK;
}

static void merge_result(KbSweepResult& r, const KbSweepResult& q, bool first) {
  if (first) { r = q; return; }
  if (q.best_val < r.best_val || (q.best_val == r.best_val && q.best_idx < r.best_idx)) {
    r.best_val = q.best_val; r.best_idx = q.best_idx;
  }
  if (q.min_u < r.min_u || (q.min_u == r.min_u && q.min_u_idx < r.min_u_idx)) {
    r.min_u = q.min_u; r.min_u_idx = q.min_u_idx;
  }
  r.n_fail += q.n_fail; r.n_fail_minus += q.n_fail_minus; r.n_fail_plus += q.n_fail_plus;
  r.n_points += q.n_points;
}

// Host-buffer sweep: 2^20-site chunks through two staging buffers; the host->device copy of chunk
// i+1 runs on a second stream while chunk i is swept, per-chunk reductions stay on the device and
// come back in one copy at the end.  (With pageable host memory the CUDA runtime stages the copies
// itself and the overlap shrinks; pinned buffers give the full overlap.)
kb_status kb_predict(kb_model* m, long long npts, const double* x, const kb_acq_params* ap,
                     double* const* outs, int acq, KbSweepResult* result) {
  if (!m || !x || npts < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(m->device));
  int nout = 0;
  int slot[KB_OUT_COUNT];
  for (int k = 0; k < KB_OUT_COUNT; ++k) slot[k] = (outs && outs[k]) ? nout++ : -1;
  const long long chunk = npts < (1LL << 20) ? npts : (1LL << 20);
  const long long nchunks = (npts + chunk - 1) / chunk;
  if (!m->copy_stream) {
    CK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&m->ev_h2d[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&m->ev_swept[i], cudaEventDisableTiming));
    }
  }
  if (chunk > m->stage_pts || nout > m->stage_nout) {
    CK(cudaStreamSynchronize(m->stream));
    CK(cudaStreamSynchronize(m->copy_stream));
    const long long cp = chunk > m->stage_pts ? chunk : m->stage_pts;
    const int no = nout > m->stage_nout ? nout : m->stage_nout;
    for (int i = 0; i < 2; ++i) {
      kb_dev_free(m->stage_x[i]); kb_dev_free(m->stage_out[i]);
      m->stage_x[i] = m->stage_out[i] = nullptr;
      CK(kb_dev_malloc(&m->stage_x[i], sizeof(double) * (size_t)cp * m->d));
      if (no > 0) CK(kb_dev_malloc(&m->stage_out[i], sizeof(double) * (size_t)cp * no));
    }
    m->stage_pts = cp; m->stage_nout = no;
  }
  if (nchunks > m->results_cap) {
    CK(cudaStreamSynchronize(m->stream));
    kb_dev_free(m->results_dev);
    m->results_dev = nullptr;
    CK(kb_dev_malloc(&m->results_dev, sizeof(KbSweepResult) * (size_t)nchunks));
    m->results_cap = nchunks;
  }
  if (nchunks == 1) {
    // one chunk: nothing to overlap, so no second stream and no events (a local acquisition optimiser
    // calls this path one design at a time -- every API call saved is latency).  Small calls go through
    // pinned staging so that the copies back do not each serialise on pageable memory.
    const bool pinned = npts <= 2048;
    double* hx = nullptr;
    if (pinned) {
      const size_t need = sizeof(double) * ((size_t)npts * m->d + (size_t)nout * npts) + sizeof(KbSweepResult);
      if (need > m->pin_small_bytes) {
        CK(cudaStreamSynchronize(m->stream));
        g_pinned.put(m->pin_small, m->pin_small_bytes);
        m->pin_small = nullptr; m->pin_small_bytes = 0;
        size_t cap = 1 << 16;
        while (cap < need) cap *= 2;
        m->pin_small = static_cast<double*>(g_pinned.get(cap));
        if (!m->pin_small) KB_FAIL(KB_ERR_CUDA, "cannot allocate %zu bytes of pinned staging", cap);
        m->pin_small_bytes = cap;
      }
      hx = m->pin_small;
      std::memcpy(hx, x, sizeof(double) * (size_t)npts * m->d);
    }
    double* hout = pinned ? hx + (size_t)npts * m->d : nullptr;
    KbSweepResult* hres = pinned ? reinterpret_cast<KbSweepResult*>(hout + (size_t)nout * npts) : result;
    CK(cudaMemcpyAsync(m->stage_x[0], pinned ? hx : x, sizeof(double) * (size_t)npts * m->d, cudaMemcpyHostToDevice,
                       m->stream));
    double* od[KB_OUT_COUNT];
    for (int k = 0; k < KB_OUT_COUNT; ++k)
      od[k] = (slot[k] >= 0) ? m->stage_out[0] + (size_t)slot[k] * m->stage_pts : nullptr;
    kb_status s1 = kb_predict_dev(m, npts, m->stage_x[0], 0, ap, od, acq, result ? m->results_dev : nullptr);
    if (s1 != KB_OK) return s1;
    for (int k = 0; k < KB_OUT_COUNT; ++k)
      if (slot[k] >= 0)
        CK(cudaMemcpyAsync(pinned ? hout + (size_t)slot[k] * npts : outs[k], od[k], sizeof(double) * (size_t)npts,
                           cudaMemcpyDeviceToHost, m->stream));
    if (result)
      CK(cudaMemcpyAsync(hres, m->results_dev, sizeof(KbSweepResult), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (pinned) {
      for (int k = 0; k < KB_OUT_COUNT; ++k)
        if (slot[k] >= 0) std::memcpy(outs[k], hout + (size_t)slot[k] * npts, sizeof(double) * (size_t)npts);
      if (result) *result = *hres;
    }
    return KB_OK;
  }
  // the caller's stream may still be using the staging buffers of a previous call
  CK(cudaEventRecord(m->ev_swept[0], m->stream));
  CK(cudaEventRecord(m->ev_swept[1], m->stream));
  for (long long i = 0; i < nchunks; ++i) {
    const int buf = (int)(i & 1);
    const long long off = i * chunk;
    const long long cur = (npts - off < chunk) ? (npts - off) : chunk;
    CK(cudaStreamWaitEvent(m->copy_stream, m->ev_swept[buf], 0));
    CK(cudaMemcpyAsync(m->stage_x[buf], x + (size_t)off * m->d, sizeof(double) * (size_t)cur * m->d,
                       cudaMemcpyHostToDevice, m->copy_stream));
    CK(cudaEventRecord(m->ev_h2d[buf], m->copy_stream));
    CK(cudaStreamWaitEvent(m->stream, m->ev_h2d[buf], 0));
    double* od[KB_OUT_COUNT];
    for (int k = 0; k < KB_OUT_COUNT; ++k)
      od[k] = (slot[k] >= 0) ? m->stage_out[buf] + (size_t)slot[k] * m->stage_pts : nullptr;
    kb_status s = kb_predict_dev(m, cur, m->stage_x[buf], off, ap, od, acq, m->results_dev + i);
    if (s != KB_OK) return s;
    CK(cudaEventRecord(m->ev_swept[buf], m->stream));
    for (int k = 0; k < KB_OUT_COUNT; ++k)
      if (slot[k] >= 0)
        CK(cudaMemcpyAsync(outs[k] + off, od[k], sizeof(double) * (size_t)cur, cudaMemcpyDeviceToHost, m->stream));
  }
  std::vector<KbSweepResult> parts((size_t)nchunks);
  CK(cudaMemcpyAsync(parts.data(), m->results_dev, sizeof(KbSweepResult) * (size_t)nchunks, cudaMemcpyDeviceToHost,
                     m->stream));
  CK(cudaStreamSynchronize(m->stream));
  KbSweepResult total{};
  for (long long i = 0; i < nchunks; ++i) merge_result(total, parts[(size_t)i], i == 0);
  if (result) *result = total;
  return KB_OK;
}

struct kb_points {
  int device = 0, d = 0;
  long long npts = 0;
  double* x = nullptr;
  double* out = nullptr;       // lazily sized output staging [nout][chunk]
  long long out_cap = 0;
};

kb_status kb_points_create(kb_points** out, int device, long long npts, int d, const double* x) {
  if (!out || !x || npts < 1 || d < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (kb_device_count() == 0) KB_FAIL(KB_ERR_CUDA, "no CUDA device visible: no CPU fallback");
  CK(cudaSetDevice(device));
  kb_points* p = new kb_points();
  p->device = device; p->d = d; p->npts = npts;
  if (kb_dev_malloc(&p->x, sizeof(double) * (size_t)npts * d) != cudaSuccess) {
    delete p;
    KB_FAIL(KB_ERR_CUDA, "cannot allocate %lld x %d points on device %d", npts, d, device);
  }
  // chunked so that pageable host memory does not need one huge pinned staging area
  const long long chunk = 1LL << 22;
  for (long long off = 0; off < npts; off += chunk) {
    const long long cur = (npts - off < chunk) ? (npts - off) : chunk;
    CK(cudaMemcpy(p->x + (size_t)off * d, x + (size_t)off * d, sizeof(double) * (size_t)cur * d, cudaMemcpyHostToDevice));
  }
  *out = p;
  return KB_OK;
}

kb_status kb_points_generate(kb_points** out, int device, long long npts, int d, int dist, double p0, double p1,
                             const double* lb, const double* ub, unsigned long long seed, long long first_index) {
  if (!out || npts < 1 || d < 1 || dist < 0 || dist > 3) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (dist == KB_DIST_UNIFORM && (!lb || !ub))
    KB_FAIL(KB_ERR_INVALID, "type 'random' is selected, please input lower bound and upper bound value");
  if (kb_device_count() == 0) KB_FAIL(KB_ERR_CUDA, "no CUDA device visible: no CPU fallback");
  CK(cudaSetDevice(device));
  kb_points* p = new kb_points();
  p->device = device; p->d = d; p->npts = npts;
  if (kb_dev_malloc(&p->x, sizeof(double) * (size_t)npts * d) != cudaSuccess) {
    delete p;
    KB_FAIL(KB_ERR_CUDA, "cannot allocate %lld x %d points on device %d", npts, d, device);
  }
  double* bounds = nullptr;
  if (dist == KB_DIST_UNIFORM) {
    CK(kb_dev_malloc(&bounds, sizeof(double) * 2 * d));
    CK(cudaMemcpy(bounds, lb, sizeof(double) * d, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(bounds + d, ub, sizeof(double) * d, cudaMemcpyHostToDevice));
  }
  cudaError_t e = kb_launch_points_generate(nullptr, p->x, npts, d, dist, p0, p1, bounds, bounds ? bounds + d : nullptr,
                                            seed, first_index);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  kb_dev_free(bounds);
  g_launches += 1;
  if (e != cudaSuccess) {
    kb_dev_free(p->x);
    delete p;
    KB_FAIL(KB_ERR_CUDA, "population generator: %s", cudaGetErrorString(e));
  }
  *out = p;
  return KB_OK;
}

kb_status kb_points_generate_lowdisc(kb_points** out, int device, long long npts, int d, int kind,
                                     const unsigned int* dirs, const double* lb, const double* ub,
                                     long long first_index) {
  if (!out || npts < 1 || d < 1 || first_index < 0 || (kind != KB_DIST_HALTON && kind != KB_DIST_SOBOL))
    KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (kind == KB_DIST_SOBOL && !dirs) KB_FAIL(KB_ERR_INVALID, "the Sobol generator needs the direction numbers");
  if ((lb == nullptr) != (ub == nullptr)) KB_FAIL(KB_ERR_INVALID, "give both bounds or neither");
  if (kind == KB_DIST_SOBOL && first_index + npts > (1LL << 32))
    KB_FAIL(KB_ERR_INVALID, "32-bit direction numbers cover 2^32 Sobol points");
  if (kb_device_count() == 0) KB_FAIL(KB_ERR_CUDA, "no CUDA device visible: no CPU fallback");
  CK(cudaSetDevice(device));
  std::vector<int> primes;
  if (kind == KB_DIST_HALTON)
    for (int c = 2; (int)primes.size() < d; ++c) {
      bool is_prime = true;
      for (int p : primes) if (c % p == 0) { is_prime = false; break; }
      if (is_prime) primes.push_back(c);
    }
  kb_points* p = new kb_points();
  p->device = device; p->d = d; p->npts = npts;
  unsigned int* dirs_dev = nullptr;
  int* primes_dev = nullptr;
  double* bounds = nullptr;
  cudaError_t e = kb_dev_malloc(&p->x, sizeof(double) * (size_t)npts * d);
  if (e == cudaSuccess && kind == KB_DIST_SOBOL) {
    e = kb_dev_malloc(&dirs_dev, sizeof(unsigned int) * 32 * d);
    if (e == cudaSuccess) e = cudaMemcpy(dirs_dev, dirs, sizeof(unsigned int) * 32 * d, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && kind == KB_DIST_HALTON) {
    e = kb_dev_malloc(&primes_dev, sizeof(int) * d);
    if (e == cudaSuccess) e = cudaMemcpy(primes_dev, primes.data(), sizeof(int) * d, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && lb) {
    e = kb_dev_malloc(&bounds, sizeof(double) * 2 * d);
    if (e == cudaSuccess) e = cudaMemcpy(bounds, lb, sizeof(double) * d, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(bounds + d, ub, sizeof(double) * d, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess)
    e = kb_launch_points_lowdisc(nullptr, p->x, npts, d, kind, dirs_dev, primes_dev, bounds, bounds ? bounds + d : nullptr,
                                 first_index);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  kb_dev_free(dirs_dev); kb_dev_free(primes_dev); kb_dev_free(bounds);
  g_launches += 1;
  if (e != cudaSuccess) {
    kb_dev_free(p->x);
    delete p;
    KB_FAIL(KB_ERR_CUDA, "low-discrepancy generator: %s", cudaGetErrorString(e));
  }
  *out = p;
  return KB_OK;
}

kb_status kb_points_get_rows(const kb_points* p, long long row0, long long nrows, double* x_out) {
  if (!p || !x_out || row0 < 0 || nrows < 1 || row0 + nrows > p->npts) KB_FAIL(KB_ERR_INVALID, "bad row range");
  CK(cudaSetDevice(p->device));
  CK(cudaMemcpy(x_out, p->x + (size_t)row0 * p->d, sizeof(double) * (size_t)nrows * p->d, cudaMemcpyDeviceToHost));
  return KB_OK;
}

void kb_points_destroy(kb_points* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  kb_dev_free(p->x); kb_dev_free(p->out);
  delete p;
}

kb_status kb_predict_points(kb_model* m, const kb_points* pts, const kb_acq_params* ap, double* const* outs,
                            int acq, KbSweepResult* result) {
  if (!m || !pts) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (pts->device != m->device || pts->d != m->d)
    KB_FAIL(KB_ERR_INVALID, "point set (device %d, d=%d) does not match the model (device %d, d=%d)", pts->device,
            pts->d, m->device, m->d);
  CK(cudaSetDevice(m->device));
  int nout = 0;
  int slot[KB_OUT_COUNT];
  for (int k = 0; k < KB_OUT_COUNT; ++k) slot[k] = (outs && outs[k]) ? nout++ : -1;
  if (nout == 0) {
    // reduce-only: one launch over the whole resident population
    kb_status s = kb_predict_dev(m, pts->npts, pts->x, 0, ap, nullptr, acq, m->result_dev);
    if (s != KB_OK) return s;
    KbSweepResult r;
    CK(cudaMemcpyAsync(&r, m->result_dev, sizeof(KbSweepResult), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (result) *result = r;
    return KB_OK;
  }
  // with per-site outputs: chunks of 2^22 sites through a device staging area
  kb_points* pw = const_cast<kb_points*>(pts);
  const long long chunk = pts->npts < (1LL << 22) ? pts->npts : (1LL << 22);
  if ((long long)nout * chunk > pw->out_cap) {
    CK(cudaStreamSynchronize(m->stream));
    kb_dev_free(pw->out);
    pw->out = nullptr;
    CK(kb_dev_malloc(&pw->out, sizeof(double) * (size_t)nout * chunk));
    pw->out_cap = (long long)nout * chunk;
  }
  KbSweepResult total{};
  bool first = true;
  for (long long off = 0; off < pts->npts; off += chunk) {
    const long long cur = (pts->npts - off < chunk) ? (pts->npts - off) : chunk;
    double* od[KB_OUT_COUNT];
    for (int k = 0; k < KB_OUT_COUNT; ++k) od[k] = (slot[k] >= 0) ? pw->out + (size_t)slot[k] * chunk : nullptr;
    kb_status s = kb_predict_dev(m, cur, pts->x + (size_t)off * pts->d, off, ap, od, acq, m->result_dev);
    if (s != KB_OK) return s;
    for (int k = 0; k < KB_OUT_COUNT; ++k)
      if (slot[k] >= 0)
        CK(cudaMemcpyAsync(outs[k] + off, od[k], sizeof(double) * (size_t)cur, cudaMemcpyDeviceToHost, m->stream));
    KbSweepResult part;
    CK(cudaMemcpyAsync(&part, m->result_dev, sizeof(KbSweepResult), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    merge_result(total, part, first);
    first = false;
  }
  if (result) *result = total;
  return KB_OK;
}

// ---- 2-objective expected hypervolume improvement -------------------------------------------------
struct KbEhviBestHost { double val; long long idx; };

static kb_status ehvi_table_to_device(int k, const double* front, const double* ref, double** table_dev) {
  if (k < 1 || !front || !ref) KB_FAIL(KB_ERR_INVALID, "bad Pareto front");
  std::vector<double> table(kb_ehvi_table_doubles(k));
  kb_ehvi_build_table(k, front, ref, table.data());
  CK(kb_dev_malloc(table_dev, sizeof(double) * table.size()));
  CK(cudaMemcpy(*table_dev, table.data(), sizeof(double) * table.size(), cudaMemcpyHostToDevice));
  return KB_OK;
}

kb_status kb_exi2d(int device, long long npts, const double* mu, const double* s, int k, const double* front,
                   const double* ref, double* out) {
  if (!mu || !s || !out || npts < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (kb_device_count() == 0) KB_FAIL(KB_ERR_CUDA, "no CUDA device visible: no CPU fallback");
  CK(cudaSetDevice(device));
  int num_sms = 0;
  CK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
  double* table_dev = nullptr;
  kb_status st = ehvi_table_to_device(k, front, ref, &table_dev);
  if (st != KB_OK) return st;
  // de-interleave on the host: the kernel reads one array per quantity
  std::vector<double> cols((size_t)4 * npts);
  for (long long i = 0; i < npts; ++i) {
    cols[(size_t)i] = mu[2 * i]; cols[(size_t)(npts + i)] = s[2 * i];
    cols[(size_t)(2 * npts + i)] = mu[2 * i + 1]; cols[(size_t)(3 * npts + i)] = s[2 * i + 1];
  }
  double* buf = nullptr;
  void* red = nullptr;
  const int grid = kb_ehvi_grid(npts, num_sms);
  cudaError_t e = kb_dev_malloc(&buf, sizeof(double) * (size_t)5 * npts);
  if (e == cudaSuccess) e = kb_dev_malloc(&red, sizeof(KbEhviBestHost) * (size_t)(grid + 1));
  if (e == cudaSuccess) e = cudaMemcpy(buf, cols.data(), sizeof(double) * cols.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = kb_launch_ehvi2d(nullptr, npts, buf, buf + npts, buf + 2 * npts, buf + 3 * npts, table_dev, k, 0, buf + 4 * npts,
                         (KbEhviBestHost*)red + 1, red, 0, num_sms);
  if (e == cudaSuccess) e = cudaMemcpy(out, buf + 4 * npts, sizeof(double) * (size_t)npts, cudaMemcpyDeviceToHost);
  g_launches += 2;
  kb_dev_free(buf); kb_dev_free(red); kb_dev_free(table_dev);
  if (e != cudaSuccess) KB_FAIL(KB_ERR_CUDA, "exi2d: %s", cudaGetErrorString(e));
  return KB_OK;
}

static kb_status ehvi2d_core(kb_model* m1, kb_model* m2, long long npts, const double* x, bool x_on_device, int k,
                             const double* front, const double* ref, int spread, double* out, long long* best_idx,
                             double* best_val) {
  if (!m1 || !m2 || !x || npts < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (m1->device != m2->device || m1->d != m2->d)
    KB_FAIL(KB_ERR_INVALID, "the two objective models must share the device and the input dimension");
  if (!m1->fitted || !m2->fitted) KB_FAIL(KB_ERR_STATE, "ehvi called before kb_fit_state");
  CK(cudaSetDevice(m1->device));
  if (k < 1 || !front || !ref) KB_FAIL(KB_ERR_INVALID, "bad Pareto front");
  auto& W = m1->ehvi;
  const long long chunk = npts < (1LL << 20) ? npts : (1LL << 20);
  const int d = m1->d;
  const int grid = kb_ehvi_grid(chunk, m1->num_sms);
  // persistent workspace: the acquisition is evaluated in a loop by the optimisers
  if (!W.ev_x) {
    CK(cudaEventCreateWithFlags(&W.ev_x, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&W.ev_p2, cudaEventDisableTiming));
  }
  if (chunk > W.cap_pts || grid > W.cap_grid) {
    CK(cudaStreamSynchronize(m1->stream));
    CK(cudaStreamSynchronize(m2->stream));
    kb_dev_free(W.xd); kb_dev_free(W.buf); kb_dev_free(W.red);
    W.xd = W.buf = nullptr; W.red = nullptr;
    const long long cp = chunk > W.cap_pts ? chunk : W.cap_pts;
    const int cg = kb_ehvi_grid(cp, m1->num_sms);
    CK(kb_dev_malloc(&W.xd, sizeof(double) * (size_t)cp * d));
    CK(kb_dev_malloc(&W.buf, sizeof(double) * (size_t)5 * cp));
    CK(kb_dev_malloc(&W.red, sizeof(KbEhviBestHost) * (size_t)(cg + 1)));
    W.cap_pts = cp; W.cap_grid = cg;
  }
  // the cell table depends on the front and the reference point only: rebuilt when they change
  {
    std::vector<double> key((size_t)2 * k + 2);
    std::memcpy(key.data(), front, sizeof(double) * 2 * k);
    key[2 * k] = ref[0]; key[2 * k + 1] = ref[1];
    if (key != W.key) {
      std::vector<double> table(kb_ehvi_table_doubles(k));
      kb_ehvi_build_table(k, front, ref, table.data());
      CK(cudaStreamSynchronize(m1->stream));
      if (table.size() > W.cap_table) {
        kb_dev_free(W.table);
        W.table = nullptr;
        CK(kb_dev_malloc(&W.table, sizeof(double) * table.size()));
        W.cap_table = table.size();
      }
      CK(h2d_sync(m1->stream, W.table, table.data(), sizeof(double) * table.size()));
      W.key.swap(key);
    }
  }
  double* const table_dev = W.table;
  double* const xd = W.xd;
  double* const buf = W.buf;
  void* const red = W.red;
  const cudaEvent_t ev_x = W.ev_x, ev_p2 = W.ev_p2;
  const long long bstride = W.cap_pts;          // the five result arrays are cap_pts apart
  kb_status st = KB_OK;
  cudaError_t e = cudaSuccess;
  const int sid = spread ? KB_OUT_S : KB_OUT_SSQR;
  for (long long off = 0; off < npts && e == cudaSuccess && st == KB_OK; off += chunk) {
    const long long cur = (npts - off < chunk) ? (npts - off) : chunk;
    // the second model's sweep of the previous chunk must be done with xd before it is overwritten
    e = cudaStreamWaitEvent(m1->stream, ev_p2, 0);
    const double* xc = x + (size_t)off * d;
    if (!x_on_device) {
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(xd, xc, sizeof(double) * (size_t)cur * d, cudaMemcpyHostToDevice, m1->stream);
      xc = xd;
    }
    if (e == cudaSuccess) e = cudaEventRecord(ev_x, m1->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(m2->stream, ev_x, 0);
    if (e != cudaSuccess) break;
    double* o1[KB_OUT_COUNT] = {nullptr};
    double* o2[KB_OUT_COUNT] = {nullptr};
    o1[KB_OUT_PRED] = buf; o1[sid] = buf + bstride;
    o2[KB_OUT_PRED] = buf + 2 * bstride; o2[sid] = buf + 3 * bstride;
    // the two objective models sweep the chunk concurrently on their own streams
    st = kb_predict_dev(m1, cur, xc, off, nullptr, o1, KB_OUT_PRED, nullptr);
    if (st == KB_OK) st = kb_predict_dev(m2, cur, xc, off, nullptr, o2, KB_OUT_PRED, nullptr);
    if (st != KB_OK) break;
    e = cudaEventRecord(ev_p2, m2->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(m1->stream, ev_p2, 0);
    if (e == cudaSuccess)
      e = kb_launch_ehvi2d(m1->stream, cur, buf, buf + bstride, buf + 2 * bstride, buf + 3 * bstride, table_dev, k, off,
                           buf + 4 * bstride, (KbEhviBestHost*)red + 1, red, off > 0 ? 1 : 0, m1->num_sms);
    g_launches += 2;
    if (e == cudaSuccess && out)
      e = cudaMemcpyAsync(out + off, buf + 4 * bstride, sizeof(double) * (size_t)cur, cudaMemcpyDeviceToHost, m1->stream);
  }
  KbEhviBestHost best{0.0, -1};
  if (e == cudaSuccess && st == KB_OK)
    e = cudaMemcpyAsync(&best, red, sizeof(best), cudaMemcpyDeviceToHost, m1->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(m1->stream);
  cudaStreamSynchronize(m2->stream);
  if (st != KB_OK) return st;
  if (e != cudaSuccess) KB_FAIL(KB_ERR_CUDA, "ehvi2d: %s", cudaGetErrorString(e));
  if (best_idx) *best_idx = best.idx;
  if (best_val) *best_val = best.val;
  return KB_OK;
}

kb_status kb_ehvi2d(kb_model* m1, kb_model* m2, long long npts, const double* x, int k, const double* front,
                    const double* ref, int spread, double* out, long long* best_idx, double* best_val) {
  return ehvi2d_core(m1, m2, npts, x, false, k, front, ref, spread, out, best_idx, best_val);
}

kb_status kb_ehvi2d_points(kb_model* m1, kb_model* m2, const kb_points* pts, int k, const double* front,
                           const double* ref, int spread, double* out, long long* best_idx, double* best_val) {
  if (!m1 || !pts) KB_FAIL(KB_ERR_INVALID, "bad argument");
  if (pts->device != m1->device || pts->d != m1->d)
    KB_FAIL(KB_ERR_INVALID, "point set (device %d, d=%d) does not match the model (device %d, d=%d)", pts->device,
            pts->d, m1->device, m1->d);
  return ehvi2d_core(m1, m2, pts->npts, pts->x, true, k, front, ref, spread, out, best_idx, best_val);
}

}  // extern "C"
