// krig-b200: C-ABI (include/krigb200.h) over the sm_100a kernels -- model life cycle, workspaces, the likelihood
// pipeline and the fit state.  Host-side orchestration only: workspace ownership, task-queue construction,
// stream-ordered launches.  No arithmetic of the hot path happens on the host.
#include "kb_api_internal.h"

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
KbAugGeom make_geom(int n, int nind, bool fit_mode, bool allow_seated) {
  KbAugGeom g{};
  g.n = n;
  g.n_pad = round_up(n, KB_TILE);
  g.nb = g.n_pad / KB_TILE;
  g.nind = nind;
  g.q = 1 + nind;
  g.sb = round_up(g.q, KB_TILE) / KB_TILE;
  static const char* seated_env = getenv("KB_CHOL_SEATED");         // debug aid: 0 disables
  g.seated = (!fit_mode && allow_seated && g.q <= 16 && g.nb >= 3 && (n % KB_TILE) != 0 &&
              (n % KB_TILE) + g.q <= KB_TILE && !(seated_env && atoi(seated_env) == 0)) ? 1 : 0;
  if (g.seated) g.sb = 0;
  g.ib = fit_mode ? g.nb : 0;
  g.NT = g.nb + g.sb + g.ib;
  g.ld = (g.nb + g.sb) * KB_TILE;
  g.stride = (size_t)g.NT * KB_TILE * g.ld;
  return g;
}

void drop_nll_graph(KbNllGraph& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  g_pinned.put(g.h_x, g.h_bytes);
  g = KbNllGraph{};
}
void drop_nll_graphs(KbNllGraphCache& c) {
  for (KbNllGraph& g : c.g) drop_nll_graph(g);
  for (auto& sn : c.seen) sn = KbNllGraphCache::Seen{};
  c.miss_B = 0;
}

void free_ws(KbPopWs& w) {
  kb_dev_free(w.slab); kb_dev_free(w.xpop);
  w = KbPopWs{};
}

// One-tile designs (n <= 64) take the single-launch kernel of kb_small.cu and never run the task queue.
bool use_small_kernel(const kb_model* m, const KbPopWs& w) {
  static const char* small_env = getenv("KB_SMALL_NLL");            // debug aid: 0 disables
  return (&w == &m->pop) && kb_small_applies(w.g, m->d, m->model_type == KB_TYPE_KPLS) &&
         !(small_env && atoi(small_env) == 0);
}

static std::atomic<unsigned long long> g_ws_gen{0};

// Copy from caller-owned (possibly pageable) host memory on the handle's stream and wait: the blocking
// cudaMemcpy targets the legacy stream, which the handle's non-blocking stream does not order against.
cudaError_t h2d_sync(cudaStream_t st, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(st);
}

kb_status ensure_ws(kb_model* m, KbPopWs& w, int B, bool fit_mode) {
  if (B <= w.B_cap) {
    // (one-tile designs never run the task queue: kb_small.cu)
    const bool queue_unused = use_small_kernel(m, w);
    if (w.tasks_B != B && !queue_unused) {
      std::vector<int4> tasks;
      kb_chol_build_tasks(w.g, B, 2 * m->num_sms, tasks);
      if (tasks.size() > w.tasks_cap)
        KB_FAIL(KB_ERR_STATE, "task queue of %zu entries exceeds the workspace carve of %zu (B=%d, B_cap=%d)",
                tasks.size(), w.tasks_cap, B, w.B_cap);
      CK(h2d_sync(m->stream, w.tasks, tasks.data(), tasks.size() * sizeof(int4)));
      w.ntasks = (int)tasks.size();
      w.tasks_B = B;
    }
    return KB_OK;
  }
  CK(cudaStreamSynchronize(m->stream));
  // graphs captured on the old workspace carry its slab offsets: they die with it (the pooled allocator
  // may hand the very same block back, so pointer identity proves nothing)
  if (&w == &m->pop) drop_nll_graphs(m->nll_graphs);
  free_ws(w);
  w.g = make_geom(m->n, m->nind, fit_mode, true);
  const KbAugGeom& g = w.g;
  const int d = m->d, nind = m->nind;
  std::vector<int4> tasks;
  kb_chol_build_tasks(g, B, 2 * m->num_sms, tasks);
  // The task count is NOT monotone in B (two tile rows share a task only while that leaves every CTA a task
  // per column), so the carve is sized for the unpaired queue at B_cap, which bounds every B <= B_cap.
  size_t tasks_cap = tasks.size();
  {
    std::vector<int4> worst;
    kb_chol_build_tasks(g, B, 0x7fffffff, worst);
    if (worst.size() > tasks_cap) tasks_cap = worst.size();
  }
  // one slab, 256-byte aligned pieces
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_aug = carve(sizeof(double) * g.stride * B);
  const size_t o_dinv = carve(sizeof(double) * (size_t)B * g.nb * KB_TILE * KB_TILE);
  const size_t o_sync = carve(sizeof(int) * kb_chol_sync_words(g, B));
  const size_t o_tasks = carve(tasks_cap * sizeof(int4));
  const size_t o_info = carve(sizeof(int) * B);
  const size_t o_gw = carve(sizeof(double) * (size_t)B * d);
  const size_t o_ith = carve(sizeof(double) * (size_t)B * d);
  const size_t o_wk = carve(sizeof(double) * B * 4);
  const size_t o_eps = carve(sizeof(double) * B);
  const size_t o_s2h = carve(sizeof(double) * B);
  const size_t o_nug = carve(sizeof(double) * B);
  const size_t o_gk = carve(sizeof(double) * B * KB_KPLS_MAX_H);
  const size_t o_M = carve(sizeof(double) * (size_t)B * nind * nind);
  const size_t o_beta = carve(sizeof(double) * (size_t)B * nind);
  const size_t o_s2 = carve(sizeof(double) * B);
  const size_t o_nll = carve(sizeof(double) * B);
  const size_t o_lndet = carve(sizeof(double) * B);
  CK(kb_dev_malloc(&w.slab, off));
  char* base = static_cast<char*>(w.slab);
  w.aug = reinterpret_cast<double*>(base + o_aug);
  w.dinv = reinterpret_cast<double*>(base + o_dinv);
  w.sync_words = reinterpret_cast<int*>(base + o_sync);
  w.tasks = reinterpret_cast<int4*>(base + o_tasks);
  w.info = reinterpret_cast<int*>(base + o_info);
  w.hp.gw = reinterpret_cast<double*>(base + o_gw);
  w.hp.ith = reinterpret_cast<double*>(base + o_ith);
  w.hp.wk = reinterpret_cast<double*>(base + o_wk);
  w.hp.eps = reinterpret_cast<double*>(base + o_eps);
  w.hp.sigma2 = reinterpret_cast<double*>(base + o_s2h);
  w.hp.nugget = reinterpret_cast<double*>(base + o_nug);
  w.hp.gk = reinterpret_cast<double*>(base + o_gk);
  w.Mbuf = reinterpret_cast<double*>(base + o_M);
  w.beta = reinterpret_cast<double*>(base + o_beta);
  w.sigma2 = reinterpret_cast<double*>(base + o_s2);
  w.nll = reinterpret_cast<double*>(base + o_nll);
  w.lndet = reinterpret_cast<double*>(base + o_lndet);
  CK(h2d_sync(m->stream, w.tasks, tasks.data(), tasks.size() * sizeof(int4)));
  w.ntasks = (int)tasks.size();
  w.tasks_B = B;
  w.tasks_cap = tasks_cap;
  w.B_cap = B;
  w.gen = ++g_ws_gen;
  return KB_OK;
}

kb_status ensure_xpop(KbPopWs& w, int count) {
  if (count <= w.xpop_cap) return KB_OK;
  kb_dev_free(w.xpop);
  w.xpop = nullptr;
  CK(kb_dev_malloc(&w.xpop, sizeof(double) * count));
  w.xpop_cap = count;
  w.gen = ++g_ws_gen;
  return KB_OK;
}

// decode -> build -> factorise -> reduce, all stream ordered
kb_status run_pipeline(kb_model* m, KbPopWs& w, int B, int nbhyp, const double* xpop_dev, int trainvar,
                              double fixed_nugget, double* nll_dev, double* rt_out) {
  const int nth = (m->model_type == KB_TYPE_KPLS) ? m->h : m->d;
  const int rest = nbhyp - nth - (trainvar ? 1 : 0);
  int has_nugget = 0, has_weights = 0;
  if (rest == 0) { }
  else if (rest == 1) has_nugget = 1;
  else if (rest == m->nkernel) has_weights = 1;
  else if (rest == m->nkernel + 1) { has_nugget = 1; has_weights = 1; }
  else KB_FAIL(KB_ERR_INVALID, "hyper-parameter vector of length %d does not match nvar=%d trainvar=%d nkernel=%d "
               "(likelihood_func.py:121-165)", nbhyp, nth, trainvar, m->nkernel);
  if (!has_weights && m->nkernel > 1)
    KB_FAIL(KB_ERR_INVALID, "composite kernel needs %d weights in the hyper-parameter vector", m->nkernel);
  KbModelDev md{m->n, m->d, m->nind, m->n_pad, m->nb, m->kmask, m->XT, m->y, m->F};
  const bool prof = m->profile && (&w == &m->pop);
  // One-tile designs (n <= 64): a single launch decodes, builds, factors and solves (kb_small.cu)
  const bool small = use_small_kernel(m, w);
  if (small) {
    if (prof) {
      CK(cudaEventRecord(m->ev[m->pop_calls % KB_PROF_RING][0], m->stream));
      CK(cudaEventRecord(m->ev[m->pop_calls % KB_PROF_RING][1], m->stream));
      CK(cudaEventRecord(m->ev[m->pop_calls % KB_PROF_RING][2], m->stream));
    }
    CK(kb_launch_nll_small(m->stream, md, w.g, B, nbhyp, xpop_dev, nth, trainvar, has_nugget, has_weights,
                           fixed_nugget, m->nkernel, m->kernel_ids_dev, w.hp, w.aug, w.info));
    if (prof) CK(cudaEventRecord(m->ev[m->pop_calls % KB_PROF_RING][3], m->stream));
    g_launches += 2;
  } else {
    if (prof) CK(cudaEventRecord(m->ev[m->pop_calls % KB_PROF_RING][0], m->stream));
    CK(kb_launch_decode(m->stream, B, nbhyp, xpop_dev, m->d, nth, trainvar, has_nugget, has_weights,
                        fixed_nugget, m->nkernel, m->kernel_ids_dev,
                        (m->model_type == KB_TYPE_KPLS) ? m->pls : nullptr, w.hp));
    if (prof) CK(cudaEventRecord(m->ev[m->pop_calls % KB_PROF_RING][1], m->stream));
    if (m->kpls_dtiles) CK(kb_launch_build_aug_kpls(m->stream, md, w.g, B, m->h, w.hp, m->kpls_dtiles, w.aug));
    else CK(kb_launch_build_aug(m->stream, md, w.g, B, w.hp, w.aug, 0));
    CK(cudaMemsetAsync(w.info, 0, sizeof(int) * B, m->stream));
    if (prof) CK(cudaEventRecord(m->ev[m->pop_calls % KB_PROF_RING][2], m->stream));
    CK(kb_launch_chol(m->stream, w.g, B, w.aug, w.dinv, w.sync_words, w.tasks, w.ntasks, w.info, m->num_sms));
    if (prof) CK(cudaEventRecord(m->ev[m->pop_calls % KB_PROF_RING][3], m->stream));
  }
  CK(kb_launch_gls(m->stream, w.g, B, w.aug, trainvar, w.hp.sigma2, w.Mbuf, w.beta, w.sigma2, nll_dev,
                   w.info, rt_out, w.lndet));
  if (prof) CK(cudaEventRecord(m->ev[m->pop_calls % KB_PROF_RING][4], m->stream));
  if (prof) m->pop_calls += 1;
  if (!small) g_launches += m->kpls_dtiles ? 5 : 4;
  return KB_OK;
}

// small export kernels live here (pure data movement)
static __global__ void kb_export_upper_kernel(const double* __restrict__ L, int ld, int n, double* __restrict__ U) {
  const long long total = (long long)n * n;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / n), c = (int)(e % n);
    U[e] = (c >= r) ? L[(size_t)c * ld + r] : 0.0;
  }
}

// ---------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------
extern "C" {

int kb_version(void) { return 100; }
const char* kb_last_error(void) { return g_err.c_str(); }
long long kb_launch_count(void) { return g_launches.load(); }

int kb_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
  return c;
}

kb_status kb_model_create(kb_model** out, int device, int n, int d, int nind, const double* Xnorm,
                          const double* y, const double* F, int nkernel, const int* kernel_ids,
                          int model_type, int h, const double* plscoeff) {
  if (!out || !Xnorm || !y || !F || !kernel_ids) KB_FAIL(KB_ERR_INVALID, "null argument");
  if (n < 1 || d < 1 || nind < 1 || nkernel < 1 || nkernel > 8) KB_FAIL(KB_ERR_INVALID, "bad sizes");
  if (d > KB_MAX_D) KB_FAIL(KB_ERR_INVALID, "d=%d exceeds KB_MAX_D=%d", d, KB_MAX_D);
  if (model_type == KB_TYPE_KPLS && (!plscoeff || h < 1)) KB_FAIL(KB_ERR_INVALID, "KPLS needs plscoeff and h");
  int ndev = kb_device_count();
  if (ndev == 0) KB_FAIL(KB_ERR_CUDA, "no CUDA device visible: the krig-b200 path has no CPU fallback");
  if (device < 0 || device >= ndev) KB_FAIL(KB_ERR_INVALID, "device %d out of range (%d)", device, ndev);
  CK(cudaSetDevice(device));
  // (three attribute queries instead of cudaGetDeviceProperties: that call alone costs 1-2 ms, and the callers
  //  create a handle per design update)
  int cc_major = 0, cc_minor = 0, sm_count = 0;
  CK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  CK(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
  CK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
  if (cc_major != 10)
    KB_FAIL(KB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, cc_major, cc_minor);
  kb_model* m = new kb_model();
  struct Guard {               // any failure below releases the handle, its stream and its device memory
    kb_model* m;
    ~Guard() { if (m) kb_model_destroy(m); }
  } guard{m};
  m->device = device;
  m->num_sms = sm_count;
  m->n = n; m->d = d; m->nind = nind; m->nkernel = nkernel; m->model_type = model_type; m->h = h;
  m->n_pad = round_up(n, KB_TILE);
  m->nb = m->n_pad / KB_TILE;
  m->kernel_ids.assign(kernel_ids, kernel_ids + nkernel);
  for (int k = 0; k < nkernel; ++k) {
    if (kernel_ids[k] < 0 || kernel_ids[k] > 3) KB_FAIL(KB_ERR_INVALID, "Kernel option not available!");
    if (model_type == KB_TYPE_KPLS && kernel_ids[k] != KB_GAUSSIAN)
      KB_FAIL(KB_ERR_INVALID, "KPLS supports the gaussian kernel only (kernelfunc.py:43-49)");
    m->kmask |= 1 << kernel_ids[k];
  }
  CK(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
  m->stream = m->own_stream;
  double* Xtmp = nullptr;
  CK(kb_dev_malloc(&Xtmp, sizeof(double) * (size_t)n * d));
  CK(h2d_sync(m->stream, Xtmp, Xnorm, sizeof(double) * (size_t)n * d));
  CK(kb_dev_malloc(&m->XT, sizeof(double) * (size_t)d * m->n_pad));
  CK(kb_launch_transpose_pad(m->stream, Xtmp, n, d, m->XT, m->n_pad));
  g_launches += 1;
  CK(cudaStreamSynchronize(m->stream));
  kb_dev_free(Xtmp);
  CK(kb_dev_malloc(&m->y, sizeof(double) * n));
  CK(h2d_sync(m->stream, m->y, y, sizeof(double) * n));
  CK(kb_dev_malloc(&m->F, sizeof(double) * (size_t)n * nind));
  CK(h2d_sync(m->stream, m->F, F, sizeof(double) * (size_t)n * nind));
  CK(kb_dev_malloc(&m->kernel_ids_dev, sizeof(int) * nkernel));
  CK(h2d_sync(m->stream, m->kernel_ids_dev, kernel_ids, sizeof(int) * nkernel));
  if (model_type == KB_TYPE_KPLS) {
    CK(kb_dev_malloc(&m->pls, sizeof(double) * (size_t)d * h));
    CK(h2d_sync(m->stream, m->pls, plscoeff, sizeof(double) * (size_t)d * h));
  }
  if (model_type == KB_TYPE_KPLS && h <= KB_KPLS_MAX_H) {
    // kernelfunc.py:43-49: the per-component distance sums do not depend on theta -> once per design
    const size_t ntile = (size_t)m->nb * (m->nb + 1) / 2;
    CK(kb_dev_malloc(&m->kpls_dtiles, sizeof(double) * ntile * h * KB_TILE * KB_TILE));
    KbModelDev md{m->n, m->d, m->nind, m->n_pad, m->nb, m->kmask, m->XT, m->y, m->F};
    CK(kb_launch_kpls_dist(m->stream, md, h, m->pls, m->kpls_dtiles));
    g_launches += 1;
    CK(cudaStreamSynchronize(m->stream));
  }
  CK(kb_dev_malloc(&m->result_dev, sizeof(KbSweepResult)));
  guard.m = nullptr;
  *out = m;
  return KB_OK;
}

void kb_model_destroy(kb_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaStreamSynchronize(m->stream);
  drop_nll_graphs(m->nll_graphs);
  free_ws(m->pop);
  free_ws(m->fit);
  kb_dev_free(m->XT); kb_dev_free(m->y); kb_dev_free(m->F); kb_dev_free(m->pls); kb_dev_free(m->kpls_dtiles); kb_dev_free(m->kernel_ids_dev);
  kb_dev_free(m->AT); kb_dev_free(m->rt); kb_dev_free(m->dense_tmp); kb_dev_free(m->idx_dev); kb_dev_free(m->p0);
  kb_dev_free(m->LT); kb_dev_free(m->spread_dev);
  kb_dev_free(m->p1); kb_dev_free(m->psi_scratch); kb_dev_free(m->ex_scratch); kb_dev_free(m->partials);
  kb_dev_free(m->result_dev); kb_dev_free(m->results_dev); kb_dev_free(m->few_scratch);
  g_pinned.put(m->pin_small, m->pin_small_bytes);
  kb_dev_free(m->ehvi.xd); kb_dev_free(m->ehvi.buf); kb_dev_free(m->ehvi.table); kb_dev_free(m->ehvi.red);
  if (m->ehvi.ev_x) cudaEventDestroy(m->ehvi.ev_x);
  if (m->ehvi.ev_p2) cudaEventDestroy(m->ehvi.ev_p2);
  for (int i = 0; i < 2; ++i) {
    kb_dev_free(m->stage_x[i]); kb_dev_free(m->stage_out[i]);
    if (m->ev_h2d[i]) cudaEventDestroy(m->ev_h2d[i]);
    if (m->ev_swept[i]) cudaEventDestroy(m->ev_swept[i]);
  }
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  cudaStreamDestroy(m->own_stream);
  delete m;
}

kb_status kb_model_set_stream(kb_model* m, void* cuda_stream) {
  if (!m) KB_FAIL(KB_ERR_INVALID, "null model");
  CK(cudaSetDevice(m->device));
  CK(cudaStreamSynchronize(m->stream));
  // caller-owned stream (e.g. torch's current stream); NULL is the legacy default stream
  drop_nll_graphs(m->nll_graphs);
  m->stream = (cudaStream_t)cuda_stream;
  return KB_OK;
}

kb_status kb_model_use_own_stream(kb_model* m) {
  if (!m) KB_FAIL(KB_ERR_INVALID, "null model");
  CK(cudaSetDevice(m->device));
  CK(cudaStreamSynchronize(m->stream));
  drop_nll_graphs(m->nll_graphs);
  m->stream = m->own_stream;
  return KB_OK;
}

kb_status kb_model_sync(kb_model* m) {
  if (!m) KB_FAIL(KB_ERR_INVALID, "null model");
  CK(cudaSetDevice(m->device));
  CK(cudaStreamSynchronize(m->stream));
  return KB_OK;
}

kb_status kb_set_profiling(kb_model* m, int on) {
  if (!m) KB_FAIL(KB_ERR_INVALID, "null model");
  CK(cudaSetDevice(m->device));
  if (on && !m->ev[0][0]) {
    for (int r = 0; r < KB_PROF_RING; ++r) {
      for (int i = 0; i < 5; ++i) CK(cudaEventCreate(&m->ev[r][i]));
      for (int i = 0; i < 2; ++i) CK(cudaEventCreate(&m->evp[r][i]));
    }
  }
  m->profile = on ? 1 : 0;
  m->pop_calls = m->pred_calls = 0;
  return KB_OK;
}

kb_status kb_get_stage_ms(kb_model* m, int back, float* ms5) {
  if (!m || !ms5 || !m->ev[0][0]) KB_FAIL(KB_ERR_INVALID, "profiling is not enabled");
  if (back < 0 || back >= KB_PROF_RING) KB_FAIL(KB_ERR_INVALID, "history depth is %d", KB_PROF_RING);
  CK(cudaSetDevice(m->device));
  for (int i = 0; i < 5; ++i) ms5[i] = -1.f;
  if (m->pop_calls > back) {
    cudaEvent_t* e = m->ev[(m->pop_calls - 1 - back) % KB_PROF_RING];
    CK(cudaEventSynchronize(e[4]));
    for (int i = 0; i < 4; ++i) CK(cudaEventElapsedTime(&ms5[i], e[i], e[i + 1]));
  }
  if (m->pred_calls > back) {
    cudaEvent_t* e = m->evp[(m->pred_calls - 1 - back) % KB_PROF_RING];
    CK(cudaEventSynchronize(e[1]));
    CK(cudaEventElapsedTime(&ms5[4], e[0], e[1]));
  }
  return KB_OK;
}

kb_status kb_reserve_population(kb_model* m, int B) {
  if (!m || B < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(m->device));
  return ensure_ws(m, m->pop, B, false);
}

kb_status kb_nll_batch_dev(kb_model* m, int B, int nbhyp, const double* xpop_dev, int trainvar,
                           double fixed_nugget_log10, double* nll_out_dev, int* info_out_dev) {
  if (!m || !xpop_dev || !nll_out_dev || B < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(m->device));
  kb_status s = ensure_ws(m, m->pop, B, false);
  if (s != KB_OK) return s;
  s = run_pipeline(m, m->pop, B, nbhyp, xpop_dev, trainvar, fixed_nugget_log10, nll_out_dev, nullptr);
  if (s != KB_OK) return s;
  if (info_out_dev)
    CK(cudaMemcpyAsync(info_out_dev, m->pop.info, sizeof(int) * B, cudaMemcpyDeviceToDevice, m->stream));
  return KB_OK;
}

kb_status kb_nll_batch(kb_model* m, int B, int nbhyp, const double* xpop, int trainvar,
                       double fixed_nugget_log10, double* nll_out, int* info_out) {
  if (!m || !xpop || !nll_out || B < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(m->device));
  kb_status s = ensure_ws(m, m->pop, B, false);
  if (s != KB_OK) return s;
  s = ensure_xpop(m->pop, B * nbhyp);
  if (s != KB_OK) return s;
  // An optimiser repeats the same call shape hundreds of times: replay a captured graph of the whole call
  // (most of the gain is at launch-bound sizes, a few tile columns and small populations).
  static const char* graph_env = getenv("KB_NLL_GRAPH");            // debug aid: 0 disables
  KbNllGraphCache& GC = m->nll_graphs;
  if (GC.ws_gen != m->pop.gen) {      // the workspace (or its candidate buffer) was re-allocated
    drop_nll_graphs(GC);
    GC.ws_gen = m->pop.gen;
  }
  // (measured: 82 -> 66 us at n=40, 216 -> 198 us at n=500 B=16, 1337 -> 1326 us at n=1000 B=64)
  const bool want_graph = !m->profile && m->stream == m->own_stream && !GC.disabled &&
                          !(graph_env && atoi(graph_env) == 0);
  if (want_graph) {
    auto matches = [&](const KbNllGraph& g) {
      return g.exec && g.B == B && g.nbhyp == nbhyp && g.trainvar == trainvar && g.nugget == fixed_nugget_log10;
    };
    KbNllGraph* G = nullptr;
    for (KbNllGraph& g : GC.g)
      if (matches(g)) G = &g;
    if (!G) {
      KbNllGraphCache::Seen* sn = nullptr;
      for (auto& e : GC.seen)
        if (e.count > 0 && e.B == B && e.nbhyp == nbhyp && e.trainvar == trainvar && e.nugget == fixed_nugget_log10) sn = &e;
      if (!sn) {
        sn = &GC.seen[0];
        for (auto& e : GC.seen)
          if (e.stamp < sn->stamp) sn = &e;
        *sn = KbNllGraphCache::Seen{};
        sn->B = B; sn->nbhyp = nbhyp; sn->trainvar = trainvar; sn->nugget = fixed_nugget_log10;
      }
      sn->count += 1;
      sn->stamp = ++GC.clock;
      GC.miss_B = B;
      const bool again = sn->count >= 3;
      if (again) {
        sn->count = 0;
        KbNllGraph* slot = &GC.g[0];
        for (KbNllGraph& g : GC.g)
          if (g.stamp < slot->stamp) slot = &g;                 // empty slots have stamp 0
        drop_nll_graph(*slot);
        CK(cudaStreamSynchronize(m->stream));
        cudaGraph_t graph = nullptr;
        slot->h_bytes = sizeof(double) * ((size_t)B * nbhyp + B) + sizeof(int) * (size_t)B;
        slot->h_x = static_cast<double*>(g_pinned.get(slot->h_bytes));
        bool ok = slot->h_x != nullptr;
        if (ok) {
          slot->h_nll = slot->h_x + (size_t)B * nbhyp;
          slot->h_info = reinterpret_cast<int*>(slot->h_nll + B);
        }
        if (ok) ok = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          ok = cudaMemcpyAsync(m->pop.xpop, slot->h_x, sizeof(double) * B * nbhyp, cudaMemcpyHostToDevice, m->stream) == cudaSuccess;
          if (ok) ok = run_pipeline(m, m->pop, B, nbhyp, m->pop.xpop, trainvar, fixed_nugget_log10, m->pop.nll, nullptr) == KB_OK;
          if (ok) ok = cudaMemcpyAsync(slot->h_nll, m->pop.nll, sizeof(double) * B, cudaMemcpyDeviceToHost, m->stream) == cudaSuccess;
          if (ok) ok = cudaMemcpyAsync(slot->h_info, m->pop.info, sizeof(int) * B, cudaMemcpyDeviceToHost, m->stream) == cudaSuccess;
          const cudaError_t ec = cudaStreamEndCapture(m->stream, &graph);
          ok = ok && ec == cudaSuccess && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&slot->exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
          cudaGetLastError();
          GC.disabled = 1;                    // plain launches from now on (same kernels)
          drop_nll_graphs(GC);
        } else {
          slot->B = B; slot->nbhyp = nbhyp; slot->trainvar = trainvar; slot->nugget = fixed_nugget_log10;
          G = slot;
        }
      }
    }
    if (G) {
      G->stamp = ++GC.clock;
      std::memcpy(G->h_x, xpop, sizeof(double) * B * nbhyp);
      CK(cudaGraphLaunch(G->exec, m->stream));
      CK(cudaStreamSynchronize(m->stream));
      std::memcpy(nll_out, G->h_nll, sizeof(double) * B);
      if (info_out) std::memcpy(info_out, G->h_info, sizeof(int) * B);
      g_launches += m->kpls_dtiles ? 5 : 4;
      return KB_OK;
    }
  }
  CK(cudaMemcpyAsync(m->pop.xpop, xpop, sizeof(double) * B * nbhyp, cudaMemcpyHostToDevice, m->stream));
  s = run_pipeline(m, m->pop, B, nbhyp, m->pop.xpop, trainvar, fixed_nugget_log10, m->pop.nll, nullptr);
  if (s != KB_OK) return s;
  CK(cudaMemcpyAsync(nll_out, m->pop.nll, sizeof(double) * B, cudaMemcpyDeviceToHost, m->stream));
  if (info_out)
    CK(cudaMemcpyAsync(info_out, m->pop.info, sizeof(int) * B, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return KB_OK;
}

kb_status kb_nll_batch_extras(kb_model* m, int B, double* beta_out, double* sigma2_out) {
  if (!m || B < 1 || B > m->pop.B_cap) KB_FAIL(KB_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(m->device));
  CK(cudaStreamSynchronize(m->stream));
  if (beta_out) CK(cudaMemcpy(beta_out, m->pop.beta, sizeof(double) * (size_t)B * m->nind, cudaMemcpyDeviceToHost));
  if (sigma2_out) CK(cudaMemcpy(sigma2_out, m->pop.sigma2, sizeof(double) * B, cudaMemcpyDeviceToHost));
  return KB_OK;
}

kb_status kb_fit_state(kb_model* m, const double* x, int nbhyp, int trainvar, double fixed_nugget_log10,
                       double* U, double* Psi, double* BE, double* sigma2, double* nll, double* nugget_out,
                       double* wgkf_out) {
  if (!m || !x || nbhyp < 1) KB_FAIL(KB_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(m->device));
  kb_status s = ensure_ws(m, m->fit, 1, true);
  if (s != KB_OK) return s;
  s = ensure_xpop(m->fit, nbhyp);
  if (s != KB_OK) return s;
  const KbAugGeom& g = m->fit.g;
  const int n = m->n;
  if (!m->rt) {
    CK(kb_dev_malloc(&m->rt, sizeof(double) * g.n_pad));
    CK(cudaMemsetAsync(m->rt, 0, sizeof(double) * g.n_pad, m->stream));
  }
  if (!m->AT) {
    m->nrows = round_up(g.n_pad + g.q, 128);
    m->ldat = m->nrows;
    CK(kb_dev_malloc(&m->AT, sizeof(double) * (size_t)g.n_pad * m->ldat));
  }
  m->fitted = false;
  CK(cudaMemcpyAsync(m->fit.xpop, x, sizeof(double) * nbhyp, cudaMemcpyHostToDevice, m->stream));
  s = run_pipeline(m, m->fit, 1, nbhyp, m->fit.xpop, trainvar, fixed_nugget_log10, m->fit.nll, m->rt);
  if (s != KB_OK) return s;
  CK(kb_launch_fit_extract(m->stream, g, m->fit.aug, m->rt, m->AT, m->ldat));
  if (!m->spread_dev) CK(kb_dev_malloc(&m->spread_dev, sizeof(double) * 2));
  CK(kb_launch_diag_spread(m->stream, g, m->fit.aug, m->spread_dev));
  g_launches += 2;
  double h_nll = 0.0, h_s2 = 0.0, h_nug = 0.0, h_wk[4], h_spread[2] = {1.0, 1.0};
  int h_info = 0;
  CK(cudaMemcpyAsync(h_spread, m->spread_dev, sizeof(double) * 2, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaMemcpyAsync(&h_nll, m->fit.nll, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaMemcpyAsync(&h_s2, m->fit.sigma2, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaMemcpyAsync(&h_nug, m->fit.hp.nugget, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaMemcpyAsync(h_wk, m->fit.hp.wk, sizeof(double) * 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaMemcpyAsync(&h_info, m->fit.info, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  if (BE) CK(cudaMemcpyAsync(BE, m->fit.beta, sizeof(double) * m->nind, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  if (h_info != 0)
    KB_FAIL(KB_ERR_NOT_PD, "Matrix is not positive definite (pivot %d); the reference raises LinAlgError at "
            "likelihood_func.py:175", h_info);
  // Pivots reaching down to the nugget (diag L spanning more than KB_PRED_REFINE_SPREAD: dense low-dimensional designs
  // with smooth kernels) make the predictor's product with the EXPLICIT inverse 5-30x less accurate than the
  // reference's triangular solve (profiles/r02_emulate_predictor.txt, tools/emulate_predictor.py; below that spread the
  // two agree to rounding).  Such fits keep a k-major copy of the factor and the sweeps correct the product with one
  // residual step, v = v0 + W (psi - L v0), at three times the tile stream -- on the strip kernel, whatever the size.
  // KB_PRED_REFINE = 0 / 1 forces it off / on (debug aid).
  {
    static const char* ref_env = getenv("KB_PRED_REFINE");
    const bool ill = h_spread[0] > 0.0 && h_spread[1] > KB_PRED_REFINE_SPREAD * h_spread[0];
    const int refine = ref_env ? (atoi(ref_env) > 0 ? 1 : 0)
                               : (m->pred_refine_mode >= 0 ? (m->pred_refine_mode > 0 ? 1 : 0) : (ill ? 1 : 0));
    if (refine) {
      if (!m->LT) {
        CK(kb_dev_malloc(&m->LT, sizeof(double) * (size_t)g.n_pad * m->ldat));
        // (the strip kernel reads 128-column blocks: the columns past n_pad must hold zeros, not whatever was there)
        CK(cudaMemsetAsync(m->LT, 0, sizeof(double) * (size_t)g.n_pad * m->ldat, m->stream));
      }
      CK(kb_launch_fit_extract_factor(m->stream, g, m->fit.aug, m->LT, m->ldat));
      // ... and alpha, R^-1 F (formed with the explicit inverse) get one residual step against the factor
      double* res = nullptr;
      CK(kb_dev_malloc(&res, sizeof(double) * (size_t)g.q * g.n_pad));
      CK(kb_launch_fit_refine_extra(m->stream, g, m->fit.aug, m->rt, m->AT, m->LT, m->ldat, res));
      CK(cudaStreamSynchronize(m->stream));
      kb_dev_free(res);
      g_launches += 3;
    }
    m->pred_refine = refine;
  }
  m->sigma2 = h_s2;
  if (sigma2) *sigma2 = h_s2;
  if (nll) *nll = h_nll;
  if (nugget_out) *nugget_out = h_nug;
  if (wgkf_out) {
    // normalised weight of each listed kernel (likelihood_func.py:139,161).  Kernels listed
    // twice share the folded weight of their id; report an even split.
    for (int k = 0; k < m->nkernel; ++k) {
      int cnt = 0;
      for (int j = 0; j < m->nkernel; ++j) cnt += (m->kernel_ids[j] == m->kernel_ids[k]);
      wgkf_out[k] = h_wk[m->kernel_ids[k]] / cnt;
    }
  }
  if (U || Psi) {
    if (!m->dense_tmp) CK(kb_dev_malloc(&m->dense_tmp, sizeof(double) * (size_t)n * n));
    if (U) {
      kb_export_upper_kernel<<<1184, 256, 0, m->stream>>>(m->fit.aug, g.ld, n, m->dense_tmp);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(U, m->dense_tmp, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, m->stream));
      CK(cudaStreamSynchronize(m->stream));
    }
    if (Psi) {
      CK(kb_launch_corr_general(m->stream, n, n, m->d, m->XT, m->n_pad, m->XT, m->n_pad, m->fit.hp.gw,
                                m->fit.hp.ith, m->fit.hp.wk, m->kmask, m->dense_tmp, n));
      CK(cudaMemcpyAsync(Psi, m->dense_tmp, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, m->stream));
      CK(cudaStreamSynchronize(m->stream));
    }
    g_launches += (U ? 1 : 0) + (Psi ? 1 : 0);
  }
  m->fitted = true;
  return KB_OK;
}

kb_status kb_model_set_predict_refine(kb_model* m, int mode) {
  if (!m || mode < KB_REFINE_AUTO || mode > KB_REFINE_ON) KB_FAIL(KB_ERR_INVALID, "bad argument");
  m->pred_refine_mode = mode;
  return KB_OK;
}

int kb_model_predict_refined(const kb_model* m) { return (m && m->fitted) ? m->pred_refine : 0; }

kb_status kb_model_set_trend(kb_model* m, const int* idx, int nind, int d) {
  if (!m || !idx || nind != m->nind || d != m->d) KB_FAIL(KB_ERR_INVALID, "trend index shape mismatch");
  CK(cudaSetDevice(m->device));
  kb_dev_free(m->idx_dev);
  m->idx_dev = nullptr;
  int maxo = 0;
  for (int i = 0; i < nind * d; ++i) maxo = idx[i] > maxo ? idx[i] : maxo;
  m->maxorder = maxo;
  if (nind == 1 && maxo == 0) return KB_OK;   // ordinary Kriging: constant basis, fast path
  CK(kb_dev_malloc(&m->idx_dev, sizeof(int) * (size_t)nind * d));
  CK(h2d_sync(m->stream, m->idx_dev, idx, sizeof(int) * (size_t)nind * d));
  return KB_OK;
}

kb_status kb_model_set_norm(kb_model* m, int norm_type, const double* p0, const double* p1) {
  if (!m) KB_FAIL(KB_ERR_INVALID, "null model");
  if (norm_type != KB_NORM_NONE && (!p0 || !p1)) KB_FAIL(KB_ERR_INVALID, "normalisation needs two vectors");
  CK(cudaSetDevice(m->device));
  m->norm_type = norm_type;
  if (norm_type == KB_NORM_NONE) return KB_OK;
  if (!m->p0) CK(kb_dev_malloc(&m->p0, sizeof(double) * m->d));
  if (!m->p1) CK(kb_dev_malloc(&m->p1, sizeof(double) * m->d));
  CK(cudaStreamSynchronize(m->stream));          // a sweep in flight may still read the old vectors
  CK(h2d_sync(m->stream, m->p0, p0, sizeof(double) * m->d));
  CK(h2d_sync(m->stream, m->p1, p1, sizeof(double) * m->d));
  return KB_OK;
}

kb_status kb_model_set_ymap(kb_model* m, int mode, double a, double b) {
  if (!m || mode < KB_YMAP_NONE || mode > KB_YMAP_STD) KB_FAIL(KB_ERR_INVALID, "bad argument");
  m->ymode = mode; m->ya = a; m->yb = b;
  return KB_OK;
}

}  // extern "C"
