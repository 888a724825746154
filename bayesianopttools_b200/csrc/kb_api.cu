// krig-b200: C-ABI (include/krigb200.h) over the sm_100a kernels.
// Host-side orchestration only: workspace ownership, task-queue construction,
// stream-ordered launches.  No arithmetic of the hot path happens on the host.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <cmath>
#include <vector>

#include "kb_internal.h"

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

#define KB_FAIL(code, ...)                           \
  do {                                               \
    char buf_[512];                                  \
    snprintf(buf_, sizeof(buf_), __VA_ARGS__);       \
    g_err = buf_;                                    \
    return (code);                                   \
  } while (0)

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      KB_FAIL(KB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

#define KB_PROF_RING 32
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct KbPopWs {       // one augmented-matrix workspace (population or fit geometry)
  KbAugGeom g{};
  int B_cap = 0;
  double* aug = nullptr;
  double* dinv = nullptr;
  int* sync_words = nullptr;
  int4* tasks = nullptr;
  int ntasks = 0, tasks_B = 0;
  size_t tasks_cap = 0;   // int4 entries the `tasks` carve holds (the worst case over every B <= B_cap)
  unsigned long long gen = 0;   // changes whenever any device buffer of this workspace is re-allocated
  int* info = nullptr;
  KbHyper hp{};
  double *Mbuf = nullptr, *beta = nullptr, *sigma2 = nullptr, *nll = nullptr, *lndet = nullptr;
  double* xpop = nullptr;
  int xpop_cap = 0;
  void* slab = nullptr;   // ONE device allocation behind every pointer above except xpop (cudaMalloc / cudaFree cost
                          // 0.25-0.3 ms each, and the callers re-create the model at every design update)
};

// Pinned host memory comes from a process-wide pool of power-of-two blocks that are never returned to the
// driver: cudaMallocHost / cudaFreeHost cost 0.3-0.8 ms / 0.13 ms each (tools/ubench/host_costs.cu), and
// SOBO / AK-MCS / MOBO re-create the model handle at every design update.
#include <mutex>
#include <unordered_map>
namespace {
struct KbPinnedPool {
  std::mutex mu;
  std::vector<void*> free_blocks[40];
  static int cls(size_t bytes) {
    int c = 12;                                   // smallest block 4 KB
    while (((size_t)1 << c) < bytes) ++c;
    return c;
  }
  void* get(size_t bytes) {
    const int c = cls(bytes);
    {
      std::lock_guard<std::mutex> lk(mu);
      if (!free_blocks[c].empty()) {
        void* p = free_blocks[c].back();
        free_blocks[c].pop_back();
        return p;
      }
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, (size_t)1 << c) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
  }
  void put(void* p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu);
    free_blocks[cls(bytes)].push_back(p);
  }
};
KbPinnedPool g_pinned;

// Device memory likewise: blocks up to 32 MB are recycled through per-device power-of-two free lists
// (cudaMalloc / cudaFree cost 0.25-0.3 ms each and cudaFree synchronises the device; a model handle makes about
// forty allocations in its life and SOBO / AK-MCS / MOBO create one per design update).  Larger blocks -- the
// augmented-matrix slab of a big population -- go straight to the driver.  A recycled block is only handed out
// again after a device synchronisation at release time, which is what cudaFree would have done.
struct KbDevicePool {
  static constexpr size_t kMaxPooled = (size_t)32 << 20;
  static constexpr size_t kMaxCached = (size_t)1 << 30;     // per process
  std::mutex mu;
  std::unordered_map<void*, std::pair<int, int>> live;       // pointer -> (device, size class), pooled blocks only
  std::unordered_map<long long, std::vector<void*>> free_blocks;   // (device << 8 | class) -> blocks
  size_t cached = 0;
  static int cls(size_t bytes) {
    int c = 9;                                    // smallest block 512 B
    while (((size_t)1 << c) < bytes) ++c;
    return c;
  }
  cudaError_t get(void** out, size_t bytes) {
    *out = nullptr;
    if (bytes == 0) bytes = 1;
    if (bytes > kMaxPooled) return cudaMalloc(out, bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    const int c = cls(bytes);
    {
      std::lock_guard<std::mutex> lk(mu);
      auto& fl = free_blocks[((long long)dev << 8) | c];
      if (!fl.empty()) {
        *out = fl.back();
        fl.pop_back();
        cached -= (size_t)1 << c;
        live[*out] = {dev, c};
        return cudaSuccess;
      }
    }
    const cudaError_t e = cudaMalloc(out, (size_t)1 << c);
    if (e == cudaSuccess) {
      std::lock_guard<std::mutex> lk(mu);
      live[*out] = {dev, c};
    }
    return e;
  }
  void put(void* p) {
    if (!p) return;
    std::pair<int, int> dc;
    {
      std::unique_lock<std::mutex> lk(mu);
      auto it = live.find(p);
      if (it == live.end()) {                     // a large block: straight back to the driver
        lk.unlock();
        cudaFree(p);
        return;
      }
      dc = it->second;
      live.erase(it);
    }
    cudaDeviceSynchronize();                      // nothing in flight may still use the block
    bool keep;
    {
      std::lock_guard<std::mutex> lk(mu);
      keep = cached + ((size_t)1 << dc.second) <= kMaxCached;
      if (keep) {
        free_blocks[((long long)dc.first << 8) | dc.second].push_back(p);
        cached += (size_t)1 << dc.second;
      }
    }
    if (!keep) cudaFree(p);
  }
};
KbDevicePool g_devpool;
}  // namespace
template <typename T>
static inline cudaError_t kb_dev_malloc(T** out, size_t bytes) { return g_devpool.get(reinterpret_cast<void**>(out), bytes); }
static inline void kb_dev_free(void* p) { g_devpool.put(p); }

// Replayable capture of one host-buffer likelihood call (copy in -> decode -> build -> factor -> GLS ->
// copy out) for a fixed (B, nbhyp, trainvar, nugget): at KrigDemo / AK-MCS sizes the call is launch-bound
// (six launches and three copies for ~50 us of device work), and an optimiser repeats the same shape
// hundreds of times.
struct KbNllGraph {
  cudaGraphExec_t exec = nullptr;
  int B = 0, nbhyp = 0, trainvar = 0;
  double nugget = 0.0;
  double* h_x = nullptr;      // pinned staging: one pooled block  x | nll | info
  double* h_nll = nullptr;
  int* h_info = nullptr;
  size_t h_bytes = 0;
  unsigned long long stamp = 0;   // last use (LRU)
};
// A few shapes alternate in practice (a GA's population / offspring / mutants, an optimiser's value and
// gradient calls): keep four graphs, and capture a shape only when it is requested twice in a row -- a
// capture costs several plain calls.
struct KbNllGraphCache {
  KbNllGraph g[4];
  unsigned long long clock = 0;
  int disabled = 0;               // capture failed once: plain launches from then on
  // shapes requested without a graph, and how often (an optimiser alternates a few shapes: value and
  // gradient calls, a GA's population / offspring / mutants): a shape is captured on its third request
  struct Seen { int B = 0, nbhyp = 0, trainvar = 0, count = 0; double nugget = 0.0; unsigned long long stamp = 0; } seen[4];
  int miss_B = 0;                 // (non-zero once anything has been seen; cleared with the cache)
  unsigned long long ws_gen = 0;  // generation of the workspace the cached graphs were captured on
};

struct kb_model {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int num_sms = 148;
  int n = 0, d = 0, nind = 0, nkernel = 0, model_type = 0, h = 0, kmask = 0, n_pad = 0, nb = 0;
  std::vector<int> kernel_ids;
  double *XT = nullptr, *y = nullptr, *F = nullptr, *pls = nullptr;
  double* kpls_dtiles = nullptr;   // KPLS: theta-independent per-component squared distances, tile-major
  int* kernel_ids_dev = nullptr;
  KbPopWs pop, fit;
  KbNllGraphCache nll_graphs;
  // predictor state
  bool fitted = false;
  double* AT = nullptr;
  int ldat = 0, nrows = 0;
  double* rt = nullptr;
  double* dense_tmp = nullptr;   // n x n staging for U / Psi export
  double sigma2 = 0.0;
  int* idx_dev = nullptr;
  int maxorder = 0;
  int norm_type = KB_NORM_NONE;
  double *p0 = nullptr, *p1 = nullptr;
  int ymode = KB_YMAP_NONE;
  double ya = 1.0, yb = 0.0;
  // predict scratch
  double* psi_scratch = nullptr;
  double* ex_scratch = nullptr;
  KbSweepResult* partials = nullptr;
  KbSweepResult* result_dev = nullptr;
  int pred_grid = 0;
  int pred_resident = 0;
  int pred_stage_x = 1;
  // 2-objective EHVI workspace (kept on the first objective's handle: the acquisition is called in a loop)
  struct {
    double *xd = nullptr, *buf = nullptr, *table = nullptr;
    void* red = nullptr;
    long long cap_pts = 0;
    int cap_grid = 0;
    size_t cap_table = 0;
    cudaEvent_t ev_x = nullptr, ev_p2 = nullptr;
    std::vector<double> key;       // front | ref of the table currently on the device
  } ehvi;
  double* pin_small = nullptr;     // pinned staging for small predict calls: x | outputs | result
  size_t pin_small_bytes = 0;
  double* few_scratch = nullptr;   // few-sites predictor (kb_launch_predict_few)
  int few_slices = 0;              // slices of 16 sites the scratch is sized for
  // host-API staging
  double* stage_x[2] = {nullptr, nullptr};
  double* stage_out[2] = {nullptr, nullptr};
  long long stage_pts = 0;
  int stage_nout = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {}, ev_swept[2] = {};
  KbSweepResult* results_dev = nullptr;   // one reduction record per chunk of a host sweep
  long long results_cap = 0;
  // optional per-stage timing (bench roofline): decode|build|chol|gls boundaries, predict
  int profile = 0;
  cudaEvent_t ev[KB_PROF_RING][5] = {};
  cudaEvent_t evp[KB_PROF_RING][2] = {};
  long long pop_calls = 0, pred_calls = 0;
};


// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
static KbAugGeom make_geom(int n, int nind, bool fit_mode, bool allow_seated) {
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

static void drop_nll_graph(KbNllGraph& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  g_pinned.put(g.h_x, g.h_bytes);
  g = KbNllGraph{};
}
static void drop_nll_graphs(KbNllGraphCache& c) {
  for (KbNllGraph& g : c.g) drop_nll_graph(g);
  for (auto& sn : c.seen) sn = KbNllGraphCache::Seen{};
  c.miss_B = 0;
}

static void free_ws(KbPopWs& w) {
  kb_dev_free(w.slab); kb_dev_free(w.xpop);
  w = KbPopWs{};
}

// One-tile designs (n <= 64) take the single-launch kernel of kb_small.cu and never run the task queue.
static bool use_small_kernel(const kb_model* m, const KbPopWs& w) {
  static const char* small_env = getenv("KB_SMALL_NLL");            // debug aid: 0 disables
  return (&w == &m->pop) && kb_small_applies(w.g, m->d, m->model_type == KB_TYPE_KPLS) &&
         !(small_env && atoi(small_env) == 0);
}

static std::atomic<unsigned long long> g_ws_gen{0};

// Copy from caller-owned (possibly pageable) host memory on the handle's stream and wait: the blocking
// cudaMemcpy targets the legacy stream, which the handle's non-blocking stream does not order against.
static cudaError_t h2d_sync(cudaStream_t st, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(st);
}

static kb_status ensure_ws(kb_model* m, KbPopWs& w, int B, bool fit_mode) {
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

static kb_status ensure_xpop(KbPopWs& w, int count) {
  if (count <= w.xpop_cap) return KB_OK;
  kb_dev_free(w.xpop);
  w.xpop = nullptr;
  CK(kb_dev_malloc(&w.xpop, sizeof(double) * count));
  w.xpop_cap = count;
  w.gen = ++g_ws_gen;
  return KB_OK;
}

// decode -> build -> factorise -> reduce, all stream ordered
static kb_status run_pipeline(kb_model* m, KbPopWs& w, int B, int nbhyp, const double* xpop_dev, int trainvar,
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
  g_launches += 1;
  double h_nll = 0.0, h_s2 = 0.0, h_nug = 0.0, h_wk[4];
  int h_info = 0;
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

static kb_status ensure_predict_scratch(kb_model* m) {
  const int qx = 1 + m->nind;
  const int resident = kb_predict_can_warp(m->d, m->n, qx, m->idx_dev ? 1 : 0)
                           ? 2
                           : (kb_predict_can_reside(m->d, m->n_pad, qx, m->idx_dev ? 1 : 0) ? 1 : 0);
  if (m->ex_scratch && resident == m->pred_resident) return KB_OK;
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
  if (!resident) CK(kb_dev_malloc(&m->psi_scratch, sizeof(double) * (size_t)m->pred_grid * m->n_pad * strip));
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
  // a handful of sites on a large model: spread the single strip over n_pad / 64 CTAs
  static const char* few_env = getenv("KB_PREDICT_FEW");            // debug aid: 0 disables
  const int few_ny = npts <= (long long)KB_FEW_MAX * KB_FEW_MAX_SLICES ? kb_predict_few_slices(npts) : 0;
  // the scratch grows with the slice count; cap it at 256 MB (many regressors on a large design)
  const bool few_fits = few_ny > 0 &&
                        kb_predict_few_scratch_doubles(m->n_pad, a.qx, few_ny) * sizeof(double) <= (256u << 20);
  if (few_fits && m->pred_resident == 0 && !(few_env && atoi(few_env) == 0)) {
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
