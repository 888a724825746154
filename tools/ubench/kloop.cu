// K-loop structure test for the Cholesky tile tasks (kb_chol.cu::kb_kloop) WITHOUT dependencies or epilogues:
// the real left-looking access pattern (64 candidates, column sweep, operands streamed from HBM/L2 through a
// cp.async ring) for several CTA shapes.  Upper bound of what the bulk workers can reach with a given loop.
//   kloop <n_pad> <B>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(void* s, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// WR x WC warps, each MI x NI 8x8 blocks; CTA tile (WR*MI*8) x (WC*NI*8 == 64); ring of NS stages of KC columns.
// REFILL_AT: k-step after which the freed stage is refilled (0 = as the product kernel).
template <int WR, int WC, int MI, int NI, int KC, int NS, int MINB, int VAR>
__global__ void __launch_bounds__(WR * WC * 32, MINB)
kloop(const double* __restrict__ base, size_t mstride, int ld, const int4* __restrict__ tasks, int ntasks, int* head,
      double* sink, int nalias) {
  constexpr int T = WR * WC * 32, AROWS = WR * MI * 8, BROWS = WC * NI * 8, SROWS = AROWS + BROWS, LDK = KC + 4;
  extern __shared__ __align__(16) double ring[];
  __shared__ int s_task;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
  const int wr = warp / WC, wc = warp % WC;
  double tot = 0.0;
  while (true) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(head, 1);
    __syncthreads();
    const int tk = s_task;
    if (tk >= ntasks) break;
    const int4 t = tasks[tk];
    const double* A = base + (size_t)(t.x % nalias) * mstride + (size_t)t.y * 64 * ld;
    const double* Bm = base + (size_t)(t.x % nalias) * mstride + (size_t)t.z * 64 * ld;
    const int nchunks = t.z * 64 / KC;
    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    auto load = [&](int c, int stage, int part, int nparts) {
      double* as = ring + (size_t)stage * SROWS * LDK;
      double* bs = as + AROWS * LDK;
      const int col0 = c * KC;
      constexpr int UPR = KC / 2;     // 16-byte units per row
      if (VAR == 1) return;
#pragma unroll
      for (int j = 0; j < (AROWS * UPR + T - 1) / T; ++j) {
        if (j % nparts != part) continue;
        const int idx = tid + j * T;
        if ((AROWS * UPR) % T == 0 || idx < AROWS * UPR) {
          const int row = idx / UPR, seg = (idx % UPR) * 2;
          cp16(&as[row * LDK + seg], A + (size_t)row * ld + col0 + seg);
        }
      }
#pragma unroll
      for (int j = 0; j < (BROWS * UPR + T - 1) / T; ++j) {
        if ((j + (AROWS * UPR + T - 1) / T) % nparts != part) continue;
        const int idx = tid + j * T;
        if ((BROWS * UPR) % T == 0 || idx < BROWS * UPR) {
          const int row = idx / UPR, seg = (idx % UPR) * 2;
          cp16(&bs[row * LDK + seg], Bm + (size_t)row * ld + col0 + seg);
        }
      }
    };
#pragma unroll
    for (int s = 0; s < NS - 1; ++s) {
      if (s < nchunks) load(s, s, 0, 1);
      cp_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
      cp_wait<NS - 2>();
      __syncthreads();
      const int cn = c + NS - 1;
      const double* as = ring + (size_t)(c % NS) * SROWS * LDK;
      const double* bs = as + AROWS * LDK;
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        double af[MI], bf[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) af[mi] = as[(wr * MI * 8 + mi * 8 + gq) * LDK + ks * 4 + tq];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) bf[ni] = bs[(wc * NI * 8 + ni * 8 + gq) * LDK + ks * 4 + tq];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        if (VAR == 2) {
          if (cn < nchunks) load(cn, cn % NS, ks, KC / 4);
          if (ks == KC / 4 - 1) cp_commit();
        } else if (ks == 0) {
          if (cn < nchunks) load(cn, cn % NS, 0, 1);
          cp_commit();
        }
      }
    }
    cp_wait<0>();
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) tot += acc[mi][ni][0] + acc[mi][ni][1];
  }
  if (tot == 123.456) sink[tid] = tot;
}


// ---- variant 3: per-row cp.async.bulk copies issued by warp 0, full/empty mbarriers per stage, no CTA barrier in the loop
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile("{\n.reg .pred P1;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra WD;\nbra WL;\nWD:\n}" ::"r"(su32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* s, const void* g, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(s)), "l"(g), "r"(bytes), "r"(su32(b)) : "memory");
}
template <int WR, int WC, int MI, int NI, int KC, int NS, int MINB>
__global__ void __launch_bounds__(WR * WC * 32, MINB)
kloop_tma(const double* __restrict__ base, size_t mstride, int ld, const int4* __restrict__ tasks, int ntasks, int* head,
          double* sink, int nalias) {
  constexpr int NW = WR * WC, AROWS = WR * MI * 8, BROWS = WC * NI * 8, SROWS = AROWS + BROWS, LDK = KC + 4;
  extern __shared__ __align__(16) double ring[];
  __shared__ int s_task;
  __shared__ __align__(8) unsigned long long full[NS], empty[NS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
  const int wr = warp / WC, wc = warp % WC;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  __syncthreads();
  double tot = 0.0;
  unsigned gchunk = 0;          // chunks consumed so far by this CTA (all tasks): stage = gchunk % NS, phase = gchunk / NS
  while (true) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(head, 1);
    __syncthreads();
    const int tk = s_task;
    if (tk >= ntasks) break;
    const int4 t = tasks[tk];
    const double* A = base + (size_t)(t.x % nalias) * mstride + (size_t)t.y * 64 * ld;
    const double* Bm = base + (size_t)(t.x % nalias) * mstride + (size_t)t.z * 64 * ld;
    const int nchunks = t.z * 64 / KC;
    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    // producer (warp 0): chunk index g (global count) goes to stage g % NS once its previous occupant (g - NS) is consumed
    auto produce = [&](int c) {
      const unsigned g = gchunk + c;       // valid because warp 0 calls this with its own gchunk base of the task
      const int stage = g % NS;
      if (g >= NS) mbar_wait(&empty[stage], ((g / NS) - 1) & 1);
      double* as = ring + (size_t)stage * SROWS * LDK;
      if (lane == 0) mbar_expect(&full[stage], SROWS * KC * 8);
      __syncwarp();
      for (int r = lane; r < SROWS; r += 32) {
        const double* src = (r < AROWS ? A + (size_t)r * ld : Bm + (size_t)(r - AROWS) * ld) + c * KC;
        bulk_g2s(&as[r * LDK], src, KC * 8, &full[stage]);
      }
    };
    if (warp == 0)
      for (int s = 0; s < NS - 1 && s < nchunks; ++s) produce(s);
    for (int c = 0; c < nchunks; ++c) {
      const unsigned g = gchunk + c;
      const int stage = g % NS;
      mbar_wait(&full[stage], (g / NS) & 1);
      const double* as = ring + (size_t)stage * SROWS * LDK;
      const double* bs = as + AROWS * LDK;
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        double af[MI], bf[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) af[mi] = as[(wr * MI * 8 + mi * 8 + gq) * LDK + ks * 4 + tq];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) bf[ni] = bs[(wc * NI * 8 + ni * 8 + gq) * LDK + ks * 4 + tq];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        if (ks == 0 && warp == 0 && c + NS - 1 < nchunks) produce(c + NS - 1);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
    gchunk += nchunks;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) tot += acc[mi][ni][0] + acc[mi][ni][1];
  }
  if (tot == 123.456) sink[tid] = tot;
}

static double* g_base; static size_t g_mstride; static int g_ld, g_nb, g_B; static int g_alias; static int* g_head; static double* g_sink;

template <int WR, int WC, int MI, int NI, int KC, int NS, int MINB, int VAR = 0>
void run(const char* name) {
  constexpr int AROWS = WR * MI * 8, TR = AROWS / 64;
  std::vector<int4> tasks;
  double flop = 0;
  for (int k = 1; k < g_nb; ++k)
    for (int b = 0; b < g_B; ++b)
      for (int i = k + 1; i + TR <= g_nb; i += TR) { tasks.push_back(make_int4(b, i, k, 0)); flop += 2.0 * AROWS * 64 * k * 64; }
  int4* d_tasks; cudaMalloc(&d_tasks, tasks.size() * sizeof(int4));
  cudaMemcpy(d_tasks, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice);
  const size_t smem = sizeof(double) * NS * (AROWS + 64) * (KC + 4);
  auto kern = VAR == 3 ? kloop_tma<WR, WC, MI, NI, KC, NS, MINB> : kloop<WR, WC, MI, NI, KC, NS, MINB, VAR>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WR * WC * 32, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(g_head, 0, 4);
    cudaEventRecord(e0);
    kern<<<148 * occ, WR * WC * 32, smem>>>(g_base, g_mstride, g_ld, d_tasks, (int)tasks.size(), g_head, g_sink, g_alias);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-34s CTA %3dx64 thr %3d KC %2d NS %d smem %6zu occ %d: %8.3f ms  %6.2f TFLOP/s %s\n", name, AROWS, WR * WC * 32, KC, NS,
         smem, occ, best, flop / (best * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d_tasks);
}

// ---- variant 4: 2-D tensor-map TMA (one instruction per 64-row x 16-column box, 128-byte swizzle), dense stages,
// k-permuted fragment loads (k-step ks takes columns 2ks, 2ks+1, 8+2ks, 9+2ks: conflict-free under the swizzle)
#include <cuda.h>
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(su32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(su32(bar)) : "memory");
}
template <int WR, int WC, int MI, int NI, int NS, int MINB, bool USE_EMPTY>
__global__ void __launch_bounds__(WR * WC * 32, MINB)
kloop_tma2d(const __grid_constant__ CUtensorMap tmap, const int4* __restrict__ tasks, int ntasks, int* head, double* sink, int nalias) {
  constexpr int KC = 16, NW = WR * WC, AROWS = WR * MI * 8, BROWS = WC * NI * 8, SROWS = AROWS + BROWS;
  extern __shared__ __align__(1024) unsigned char ring_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ring_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ int s_task;
  __shared__ __align__(8) unsigned long long full[NS], empty[NS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
  const int wr = warp / WC, wc = warp % WC;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  __syncthreads();
  const int g3 = gq & 3, hi = ((tq >> 1) ^ (gq >> 2)) << 2, lo8 = (tq & 1) * 8;
  double tot = 0.0;
  unsigned gchunk = 0;
  while (true) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(head, 1);
    __syncthreads();
    const int tk = s_task;
    if (tk >= ntasks) break;
    const int4 t = tasks[tk];
    const int bsel = t.x % nalias;
    const int nchunks = t.z * 64 / KC;
    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    auto produce = [&](int c) {      // one thread
      const unsigned g = gchunk + c;
      const int stage = g % NS;
      if (USE_EMPTY && g >= NS) mbar_wait(&empty[stage], ((g / NS) - 1) & 1);
      unsigned char* as = ring + (size_t)stage * SROWS * KC * 8;
      mbar_expect(&full[stage], SROWS * KC * 8);
#pragma unroll
      for (int r = 0; r < AROWS / 64; ++r) tma_load_3d(as + r * 64 * KC * 8, &tmap, c * KC, t.y * 64 + r * 64, bsel, &full[stage]);
      tma_load_3d(as + AROWS * KC * 8, &tmap, c * KC, t.z * 64, bsel, &full[stage]);
    };
    if (tid == 0)
      for (int s = 0; s < NS - 1 && s < nchunks; ++s) produce(s);
    for (int c = 0; c < nchunks; ++c) {
      const unsigned g = gchunk + c;
      const int stage = g % NS;
      mbar_wait(&full[stage], (g / NS) & 1);
      if (!USE_EMPTY) __syncthreads();                 // every warp is done with chunk c - 1: its stage may be refilled
      if (tid == 0 && c + NS - 1 < nchunks) produce(c + NS - 1);
      const unsigned char* as = ring + (size_t)stage * SROWS * KC * 8;
      const unsigned char* bs = as + AROWS * KC * 8;
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        const int coff = (((ks ^ g3) | hi) << 4) + lo8;
        double af[MI], bf[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) af[mi] = *reinterpret_cast<const double*>(as + (wr * MI * 8 + mi * 8 + gq) * 128 + coff);
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) bf[ni] = *reinterpret_cast<const double*>(bs + (wc * NI * 8 + ni * 8 + gq) * 128 + coff);
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
      }
      if (USE_EMPTY) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
    }
    gchunk += nchunks;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) tot += acc[mi][ni][0] + acc[mi][ni][1];
  }
  if (tot == 123.456) sink[tid] = tot;
}

static CUtensorMap g_tmap;
static void make_tmap() {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  cuuint64_t gdim[3] = {(cuuint64_t)g_ld, (cuuint64_t)g_ld, (cuuint64_t)g_B};
  cuuint64_t gstr[2] = {(cuuint64_t)g_ld * 8, (cuuint64_t)g_mstride * 8};
  cuuint32_t box[3] = {16, 64, 1}, estr[3] = {1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&g_tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, g_base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
}
template <int WR, int WC, int MI, int NI, int NS, int MINB, bool USE_EMPTY>
void run_tma2d(const char* name) {
  constexpr int AROWS = WR * MI * 8, TR = AROWS / 64;
  std::vector<int4> tasks;
  double flop = 0;
  for (int k = 1; k < g_nb; ++k)
    for (int b = 0; b < g_B; ++b)
      for (int i = k + 1; i + TR <= g_nb; i += TR) { tasks.push_back(make_int4(b, i, k, 0)); flop += 2.0 * AROWS * 64 * k * 64; }
  int4* d_tasks; cudaMalloc(&d_tasks, tasks.size() * sizeof(int4));
  cudaMemcpy(d_tasks, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice);
  const size_t smem = (size_t)NS * (AROWS + 64) * 16 * 8 + 1024;
  auto kern = kloop_tma2d<WR, WC, MI, NI, NS, MINB, USE_EMPTY>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WR * WC * 32, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(g_head, 0, 4);
    cudaEventRecord(e0);
    kern<<<148 * occ, WR * WC * 32, smem>>>(g_tmap, d_tasks, (int)tasks.size(), g_head, g_sink, g_alias);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-34s CTA %3dx64 thr %3d KC 16 NS %d smem %6zu occ %d: %8.3f ms  %6.2f TFLOP/s %s\n", name, AROWS, WR * WC * 32, NS,
         smem, occ, best, flop / (best * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d_tasks);
}

int main(int argc, char** argv) {
  const int npad = argc > 1 ? atoi(argv[1]) : 2048;
  g_B = argc > 2 ? atoi(argv[2]) : 64;
  g_alias = argc > 3 ? atoi(argv[3]) : g_B; g_nb = npad / 64; g_ld = npad; g_mstride = (size_t)npad * npad;
  cudaMalloc(&g_base, g_mstride * g_B * sizeof(double));
  cudaMemset(g_base, 0, g_mstride * g_B * sizeof(double));
  cudaMalloc(&g_head, 4); cudaMalloc(&g_sink, 8192);
  printf("n_pad %d B %d distinct matrices %d (%.2f GB)\n", npad, g_B, g_alias, g_mstride * g_B * 8 / 1e9);
  make_tmap();
  //   WR WC MI NI KC NS MINB VAR
  run<4, 2, 4, 4, 32, 2, 2, 0>("product: 8 warps 32x32, KC32 NS2");
  run<4, 2, 4, 4, 32, 2, 2, 1>("  .. no loads (stale smem)");
  run<4, 2, 4, 4, 16, 3, 2, 0>("8 warps 32x32, KC16 NS3");
  run_tma2d<4, 2, 4, 4, 3, 2, false>("  TMA 2-D swizzled, NS3, syncthreads");
  run_tma2d<4, 2, 4, 4, 4, 2, false>("  TMA 2-D swizzled, NS4, syncthreads");
  run_tma2d<4, 2, 4, 4, 4, 2, true>("  TMA 2-D swizzled, NS4, empty mbarriers");
  run_tma2d<4, 2, 2, 4, 4, 2, false>("  64-row TMA 2-D, NS4, syncthreads");
  run_tma2d<4, 2, 2, 4, 6, 2, true>("  64-row TMA 2-D, NS6, empty mbarriers");
  run_tma2d<8, 2, 4, 4, 4, 1, true>("  16 warps 256x64 TMA 2-D NS4 empty");
  return 0;
}
