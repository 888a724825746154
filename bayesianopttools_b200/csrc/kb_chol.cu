// krig-b200 K2: population-batched blocked FP64 Cholesky on the FP64 tensor pipe.
//
// Replaces np.linalg.cholesky (likelihood_func.py:175) and the forward halves of
// the six triangular solves (likelihood_func.py:183-193) for every candidate of
// the hyper-parameter population at once.
//
// Design (B200-first, not LAPACK-shaped):
//  * Each candidate's matrix is the AUGMENTED lower-trapezoid [R ; (y F)^T ; I?]:
//    the right-hand sides ride along as extra tile ROWS, so the panel solve of
//    the factorisation also produces L^-1 y, L^-1 F (and L^-T in fit mode), and
//    the trailing Schur tile is -(L^-1[y F])^T (L^-1[y F]) = the GLS normal
//    equations.  No separate, latency-bound forward substitution exists.
//  * Left-looking TILE task graph: task (i,k) computes
//        C = A[i,k] - sum_{p<k} L[i,p] L[k,p]^T        (DMMA m8n8k4, K = 64k)
//    with the accumulator held in registers for the whole K loop (each tile is
//    read once and written once), then POTRF+inverse (diagonal) or a DMMA
//    multiply by the inverted diagonal block (off-diagonal).
//  * One persistent launch for the whole population: CTAs pull tasks from a
//    global queue ordered (with one-column look-ahead) so that every dependency
//    is EARLIER in the queue; tile completion is published with release/acquire
//    flags.  A CTA that claimed a task is running, so the oldest unfinished task
//    always has all inputs -> no deadlock, no grid-wide barrier, and all SMs
//    stay busy across candidates regardless of population size.
//  * Operand tiles stream L2 -> shared memory through a 3-stage cp.async ring;
//    shared tiles are padded (row stride == 4 mod 16 doubles) so every DMMA
//    fragment load is bank-conflict free.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kb_common.cuh"
#include "kb_internal.h"

#define KB_KC 32        // K chunk (doubles) per pipeline stage
#define KB_STAGES 3
#define KB_LDK 36       // padded row stride of a 64 x KC chunk
#define KB_LDC 68       // padded row stride of a 64 x 64 tile

struct KbCholSmem {
  double a[KB_STAGES][KB_TILE][KB_LDK];
  double b[KB_STAGES][KB_TILE][KB_LDK];
  int task;
  int ready;
};
static_assert(sizeof(double) * KB_STAGES * KB_TILE * KB_LDK >= sizeof(double) * KB_TILE * KB_LDC,
              "epilogue tiles alias the pipeline buffers");


__device__ __forceinline__ unsigned long long kb_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

#define KB_LDT 12       // row stride of the per-warp 16x8 scratch tile (12 == 12 mod 16: conflict free)
static_assert(sizeof(double) * (KB_TILE * KB_LDC + 8 * 16 * KB_LDT) <=
                  sizeof(double) * KB_STAGES * KB_TILE * KB_LDK, "diag scratch must fit behind Ds");

// 1/sqrt(p) to full FP64 accuracy with a 4-deep dependent chain: hardware approximation
// (relative error ~2^-23) + one third-order step  y (1 + e/2 + 3e^2/8),  e = 1 - p y^2.
__device__ __forceinline__ double kb_rsqrt(double p) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  const double py = p * y;
  const double e = fma(-py, y, 1.0);
  const double ye = y * e;
  const double q = fma(e, 0.375, 0.5);
  return fma(ye, q, y);
}

// 64x64 SPD tile in Cs (lower part used) -> Cs = L (lower), Ds = L^-1 (lower, zeros above).
// Right-looking with 16-wide panels.  The 16x16 diagonal blocks are factored AND inverted by
// warp 0 entirely in registers (lane r owns row r; columns travel by shuffle) - the only serial
// chain left is 64 x (rsqrt -> scale -> shuffle).  Panels, trailing updates and the off-diagonal
// blocks of the inverse are 8x8x4 DMMA products on padded shared tiles (conflict free).
__device__ __forceinline__ void kb_diag_factor_inverse(double (*Cs)[KB_LDC], double (*Ds)[KB_LDC],
                                                       double* Ts_base, int* info, int pivot_base) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const unsigned FULL = 0xffffffffu;
  for (int e = tid; e < KB_TILE * KB_TILE; e += KB_THREADS) Ds[e >> 6][e & 63] = 0.0;
  __syncthreads();
  for (int jb = 0; jb < 4; ++jb) {
    const int c0 = jb * 16;
    {
      // 16x16 diagonal block: factor AND invert with one element of each per thread
      // (thread (r,c) keeps A[r][c] and W[r][c] in registers).  Per column j the pivot column
      // and the finished inverse row travel through a double-buffered 32-double shared
      // exchange: ONE block barrier per column, and every thread redoes the 4-op rsqrt so the
      // serial chain is  barrier -> rsqrt -> scale -> fma.  (A single-warp register/shuffle
      // version was issue-bound at ~4 clk per instruction.)
      const int r = tid >> 4, c = tid & 15;
      double a = Cs[c0 + r][c0 + c];
      double w = (r == c) ? 1.0 : 0.0;
      double* xch = Ts_base;                       // [2][32]: column of A | row of W
#pragma unroll 1
      for (int j = 0; j < 16; ++j) {
        double* buf = xch + (j & 1) * 32;
        if (c == j) buf[r] = a;
        if (r == j) buf[16 + c] = w;
        __syncthreads();
        double piv = buf[j];
        if (!(piv > 0.0)) {          // not positive definite (or NaN): flag, keep the batch alive
          if (tid == 0) atomicCAS(info, 0, pivot_base + c0 + j + 1);
          piv = 1.0;
        }
        const double inv = kb_rsqrt(piv);
        const double lr = (r == j) ? piv * inv : buf[r] * inv;     // L[r][j]
        const double lc = buf[c] * inv;                            // L[c][j]
        const double xj = buf[16 + c] * inv;                       // X[j][c]
        if (c == j) a = lr;
        else if (c > j) a = fma(-lr, lc, a);
        if (r == j) w = xj;
        else if (r > j) w = fma(-lr, xj, w);
      }
      if (c <= r) Cs[c0 + r][c0 + c] = a;
      Ds[c0 + r][c0 + c] = (c <= r) ? w : 0.0;
    }
    __syncthreads();
    if (jb == 3) break;
    // panel: rows below the block, P = A * inv(L_D)^T ; one warp per 8 rows, both 8-column halves
    const int nrb = (KB_TILE - 16 - c0) / 8;
    if (warp < nrb) {
      const int r0 = c0 + 16 + 8 * warp;
      double af[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) af[ks] = Cs[r0 + gq][c0 + ks * 4 + tq];
      double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) kb_dmma(p[ni][0], p[ni][1], af[ks], Ds[c0 + ni * 8 + gq][c0 + ks * 4 + tq]);
      __syncwarp();
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
        *reinterpret_cast<double2*>(&Cs[r0 + gq][c0 + ni * 8 + 2 * tq]) = make_double2(p[ni][0], p[ni][1]);
    }
    __syncthreads();
    // trailing update of the lower 8x8 tiles: C -= P P^T
    const int mt = nrb, ntile = mt * (mt + 1) / 2;
    for (int t = warp; t < ntile; t += 8) {
      int ri = 0;
      while ((ri + 1) * (ri + 2) / 2 <= t) ++ri;
      const int ci = t - ri * (ri + 1) / 2;
      const int r0 = c0 + 16 + 8 * ri, cc0 = c0 + 16 + 8 * ci;
      double2 cv = *reinterpret_cast<double2*>(&Cs[r0 + gq][cc0 + 2 * tq]);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        kb_dmma(cv.x, cv.y, -Cs[r0 + gq][c0 + ks * 4 + tq], Cs[cc0 + gq][c0 + ks * 4 + tq]);
      *reinterpret_cast<double2*>(&Cs[r0 + gq][cc0 + 2 * tq]) = cv;
    }
    __syncthreads();
  }
  // off-diagonal 16x16 blocks of the inverse, by block distance:
  //   X_ij = -inv(L_ii) * sum_{k=j}^{i-1} L_ik X_kj
  double* Ts = Ts_base + warp * 16 * KB_LDT;
  for (int delta = 1; delta <= 3; ++delta) {
    const int nblk = 4 - delta;
    if (warp < 2 * nblk) {
      const int j = warp >> 1, h = warp & 1, i = j + delta;
      double t0[2] = {0.0, 0.0}, t1[2] = {0.0, 0.0};
      for (int kb = j; kb < i; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double bv = Ds[kb * 16 + ks * 4 + tq][j * 16 + h * 8 + gq];
          kb_dmma(t0[0], t0[1], Cs[i * 16 + gq][kb * 16 + ks * 4 + tq], bv);
          kb_dmma(t1[0], t1[1], Cs[i * 16 + 8 + gq][kb * 16 + ks * 4 + tq], bv);
        }
      *reinterpret_cast<double2*>(&Ts[gq * KB_LDT + 2 * tq]) = make_double2(t0[0], t0[1]);
      *reinterpret_cast<double2*>(&Ts[(8 + gq) * KB_LDT + 2 * tq]) = make_double2(t1[0], t1[1]);
      __syncwarp();
      double x0[2] = {0.0, 0.0}, x1[2] = {0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const double bv = Ts[(ks * 4 + tq) * KB_LDT + gq];
        kb_dmma(x0[0], x0[1], Ds[i * 16 + gq][i * 16 + ks * 4 + tq], bv);
        kb_dmma(x1[0], x1[1], Ds[i * 16 + 8 + gq][i * 16 + ks * 4 + tq], bv);
      }
      *reinterpret_cast<double2*>(&Ds[i * 16 + gq][j * 16 + h * 8 + 2 * tq]) = make_double2(-x0[0], -x0[1]);
      *reinterpret_cast<double2*>(&Ds[i * 16 + 8 + gq][j * 16 + h * 8 + 2 * tq]) = make_double2(-x1[0], -x1[1]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(KB_THREADS, 2)
kb_chol_kernel(KbAugGeom g, double* __restrict__ aug, double* __restrict__ dinv, int* __restrict__ flags,
               int epoch, const int4* __restrict__ tasks, int ntasks, int* __restrict__ counter,
               int* __restrict__ info, unsigned long long* __restrict__ trace) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  KbCholSmem& sm = *reinterpret_cast<KbCholSmem*>(smem_raw);
  double (*Cs)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(&sm.a[0][0][0]);
  double (*Ds)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(&sm.b[0][0][0]);
  double* Ts_base = &sm.b[0][0][0] + KB_TILE * KB_LDC;     // 8 warps x 16 x KB_LDT scratch after Ds

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int wr = warp >> 1, wc = warp & 1;
  const int NT = g.NT, ld = g.ld;

  while (true) {
    __syncthreads();
    if (tid == 0) sm.task = atomicAdd(counter, 1);
    __syncthreads();
    const int tk = sm.task;
    if (tk >= ntasks) break;
    const int4 T = tasks[tk];
    const int b = T.x, ti = T.y, tkc = T.z, type = T.w;
    unsigned long long tr0 = 0, tr1 = 0, tr2 = 0, tr3 = 0;
    if (trace && tid == 0) tr0 = kb_globaltimer();
    double* A = aug + (size_t)b * g.stride;
    int* fl = flags + (size_t)b * NT;
    const int kt = (tkc < g.nb) ? tkc : g.nb;

    // ---- input tiles: row-progress counters, polled only when the K loop catches up ----
    // prog[r] = (epoch << 10) + number of leading tile columns of row r that are final.
    const int ebase = epoch << 10;
    int ready_cols = 0;
    if (kt > 0) {
      if (tid == 0) {
        int ra, rb;
        while (true) {
          ra = kb_flag_acquire(&fl[ti]) - ebase;
          rb = kb_flag_acquire(&fl[tkc]) - ebase;
          if (ra >= 1 && rb >= 1) break;
          __nanosleep(64);
        }
        sm.ready = ra < rb ? ra : rb;
      }
      __syncthreads();
      ready_cols = sm.ready;
    }

    if (trace && tid == 0) tr1 = kb_globaltimer();
    // ---- C_acc = sum_p L[i,p] L[k,p]^T on the FP64 tensor pipe ------------
    double acc[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    const bool same = (ti == tkc);
    const double* Arow = A + (size_t)ti * KB_TILE * ld;
    const double* Brow = A + (size_t)tkc * KB_TILE * ld;
    const int nchunks = kt * (KB_TILE / KB_KC);

    auto load_chunk = [&](int c, int stage) {
      const int col0 = c * KB_KC;
#pragma unroll
      for (int j = 0; j < (KB_TILE * KB_KC / 2) / KB_THREADS; ++j) {
        const int idx = tid + j * KB_THREADS;
        const int row = idx >> 4, seg = (idx & 15) * 2;
        kb_cp_async16(&sm.a[stage][row][seg], Arow + (size_t)row * ld + col0 + seg);
        if (!same) kb_cp_async16(&sm.b[stage][row][seg], Brow + (size_t)row * ld + col0 + seg);
      }
    };

#pragma unroll
    for (int s = 0; s < KB_STAGES - 1; ++s) {
      if (s < nchunks) load_chunk(s, s);
      kb_cp_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
      const int cn = c + KB_STAGES - 1;
      const int pn = cn / (KB_TILE / KB_KC);          // tile column the next prefetch reads
      const bool poll = (cn < nchunks) && (pn >= ready_cols);
      if (poll && tid == 0) {
        int ra, rb;
        while (true) {
          ra = kb_flag_acquire(&fl[ti]) - ebase;
          rb = kb_flag_acquire(&fl[tkc]) - ebase;
          if (ra > pn && rb > pn) break;
          __nanosleep(64);
        }
        sm.ready = ra < rb ? ra : rb;
      }
      kb_cp_async_wait<KB_STAGES - 2>();
      __syncthreads();
      if (poll) ready_cols = sm.ready;
      {
        if (cn < nchunks) load_chunk(cn, cn % KB_STAGES);
        kb_cp_async_commit();
      }
      const int stg = c % KB_STAGES;
      const double (*as)[KB_LDK] = sm.a[stg];
      const double (*bs)[KB_LDK] = same ? sm.a[stg] : sm.b[stg];
#pragma unroll
      for (int ks = 0; ks < KB_KC / 4; ++ks) {
        double af[2], bf[4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) af[mi] = as[wr * 16 + mi * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ni] = bs[wc * 32 + ni * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) kb_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
      }
    }
    kb_cp_async_wait<0>();
    __syncthreads();

    if (trace && tid == 0) tr2 = kb_globaltimer();
    // ---- C = A[i,k] - C_acc ------------------------------------------------
    double* Atile = A + (size_t)ti * KB_TILE * ld + (size_t)tkc * KB_TILE;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
        const double2 v = __ldcg(reinterpret_cast<const double2*>(&Atile[(size_t)r * ld + c]));
        acc[mi][ni][0] = v.x - acc[mi][ni][0];
        acc[mi][ni][1] = v.y - acc[mi][ni][1];
      }

    if (type == KB_TASK_SCHUR) {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
          *reinterpret_cast<double2*>(&Atile[(size_t)r * ld + c]) =
              make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
    } else {
      // C -> shared (A-operand layout for the multiply / in-place POTRF)
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
          *reinterpret_cast<double2*>(&Cs[r][c]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
      if (type == KB_TASK_OFF) {
        // L[i,k] = C * inv(L[k,k])^T : wait for the diagonal task, then one 64^3 DMMA multiply
        if (tid == 0)
          while (kb_flag_acquire(&fl[tkc]) - ebase < tkc + 1) __nanosleep(64);
        __syncthreads();
        if (trace && tid == 0) tr3 = kb_globaltimer();
        const double* Dg = dinv + ((size_t)b * g.nb + tkc) * (KB_TILE * KB_TILE);
#pragma unroll
        for (int j = 0; j < (KB_TILE * KB_TILE / 2) / KB_THREADS; ++j) {
          const int idx = tid + j * KB_THREADS;
          const int row = idx >> 5, seg = (idx & 31) * 2;
          kb_cp_async16(&Ds[row][seg], Dg + row * KB_TILE + seg);
        }
        kb_cp_async_commit();
        kb_cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll 4
        for (int ks = 0; ks < KB_TILE / 4; ++ks) {
          double af[2], bf[4];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) af[mi] = Cs[wr * 16 + mi * 8 + gq][ks * 4 + tq];
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) bf[ni] = Ds[wc * 32 + ni * 8 + gq][ks * 4 + tq];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) kb_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
            *reinterpret_cast<double2*>(&Atile[(size_t)r * ld + c]) =
                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          }
      } else {
        // ---- diagonal tile: blocked (16) POTRF + inverse, DMMA for every block update ----
        kb_diag_factor_inverse(Cs, Ds, Ts_base, &info[b], tkc * KB_TILE);
        double* Dg = dinv + ((size_t)b * g.nb + tkc) * (KB_TILE * KB_TILE);
        for (int e = tid; e < KB_TILE * KB_TILE; e += KB_THREADS) {
          const int r = e >> 6, c = e & 63;
          Atile[(size_t)r * ld + c] = (c <= r) ? Cs[r][c] : 0.0;
          Dg[e] = Ds[r][c];
        }
      }
    }
    // ---- publish the tile ---------------------------------------------------
    __syncthreads();
    if (tid == 0 && type != KB_TASK_SCHUR) {
      __threadfence();
      kb_flag_release(&fl[ti], ebase + tkc + 1);
    }
    if (trace && tid == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned long long* t = trace + (size_t)tk * 8;
      t[0] = tr0; t[1] = tr1; t[2] = tr2; t[3] = tr3; t[4] = kb_globaltimer();
      t[5] = ((unsigned long long)b << 40) | ((unsigned long long)ti << 24) | ((unsigned long long)tkc << 8) | type;
      t[6] = smid; t[7] = blockIdx.x;
    }
  }
}

cudaError_t kb_launch_chol(cudaStream_t st, const KbAugGeom& g, int B, double* aug, double* dinv,
                           int* flags, int epoch, const int4* tasks, int ntasks, int* counter,
                           int* info, int num_sms) {
  (void)B;
  const size_t smem = sizeof(KbCholSmem);
  cudaError_t e = cudaFuncSetAttribute(kb_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(counter, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  int grid = num_sms * 2;
  if (grid > ntasks) grid = ntasks;
  if (grid < 1) grid = 1;
  // Debug aid: KB_CHOL_TRACE=<file> dumps per-task timestamps (ns) of this launch.
  static const char* trace_path = getenv("KB_CHOL_TRACE");
  unsigned long long* trace = nullptr;
  if (trace_path) {
    e = cudaMalloc(&trace, sizeof(unsigned long long) * 8 * (size_t)ntasks);
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(trace, 0, sizeof(unsigned long long) * 8 * (size_t)ntasks, st);
  }
  kb_chol_kernel<<<grid, KB_THREADS, smem, st>>>(g, aug, dinv, flags, epoch, tasks, ntasks, counter, info, trace);
  e = cudaGetLastError();
  if (trace_path && e == cudaSuccess) {
    std::vector<unsigned long long> h((size_t)ntasks * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    FILE* f = fopen(trace_path, "wb");
    if (f) { fwrite(h.data(), sizeof(unsigned long long), h.size(), f); fclose(f); }
    cudaFree(trace);
  }
  return e;
}
