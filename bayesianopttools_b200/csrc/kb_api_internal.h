// krig-b200: shared declarations of the C-ABI translation units (host side only)
//   kb_pools.cu        process-wide pinned / device block pools, last-error string, launch counter
//   kb_api.cu          model life cycle, workspaces, the likelihood pipeline, fit state
//   kb_api_predict.cu  predictor / sweep entry points, device point sets, 2-objective EHVI
//   kb_api_train.cu    device CMA-ES loop, LOOCV, correlation matrix, peak measurement
#pragma once
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

// defined in kb_pools.cu
extern thread_local std::string g_err;
extern std::atomic<long long> g_launches;

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
extern KbPinnedPool g_pinned;

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
extern KbDevicePool g_devpool;
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
  double* LT = nullptr;            // k-major copy of the factor, kept only for ill-conditioned fits (pred_refine)
  double* spread_dev = nullptr;    // [2] min / max of diag L of the last fit
  int pred_refine = 0;             // the predictor corrects its explicit-inverse product with one residual step
  int pred_refine_mode = -1;       // KB_REFINE_AUTO / OFF / ON (kb_model_set_predict_refine)
  int pred_refine_scratch = 0;     // ... and the scratch was sized for it
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



// ---- kb_api.cu: workspace and pipeline helpers shared by the entry points ----
KbAugGeom make_geom(int n, int nind, bool fit_mode, bool allow_seated);
void drop_nll_graph(KbNllGraph& g);
void drop_nll_graphs(KbNllGraphCache& c);
void free_ws(KbPopWs& w);
bool use_small_kernel(const kb_model* m, const KbPopWs& w);
cudaError_t h2d_sync(cudaStream_t st, void* dst, const void* src, size_t bytes);
kb_status ensure_ws(kb_model* m, KbPopWs& w, int B, bool fit_mode);
kb_status ensure_xpop(KbPopWs& w, int count);
kb_status run_pipeline(kb_model* m, KbPopWs& w, int B, int nbhyp, const double* xpop_dev, int trainvar,
                              double fixed_nugget, double* nll_dev, double* rt_out);
