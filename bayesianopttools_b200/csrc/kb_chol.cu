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
};
static_assert(sizeof(double) * KB_STAGES * KB_TILE * KB_LDK >= sizeof(double) * KB_TILE * KB_LDC,
              "epilogue tiles alias the pipeline buffers");

__global__ void __launch_bounds__(KB_THREADS, 2)
kb_chol_kernel(KbAugGeom g, double* __restrict__ aug, double* __restrict__ dinv, int* __restrict__ flags,
               int epoch, const int4* __restrict__ tasks, int ntasks, int* __restrict__ counter,
               int* __restrict__ info) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  KbCholSmem& sm = *reinterpret_cast<KbCholSmem*>(smem_raw);
  double (*Cs)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(&sm.a[0][0][0]);
  double (*Ds)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(&sm.b[0][0][0]);

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
    double* A = aug + (size_t)b * g.stride;
    int* fl = flags + (size_t)b * NT * NT;
    const int kt = (tkc < g.nb) ? tkc : g.nb;

    // ---- wait for the input tiles (all earlier in the queue) -------------
    for (int p = tid; p < 2 * kt; p += KB_THREADS) {
      const int* f = (p < kt) ? &fl[ti * NT + p] : &fl[tkc * NT + (p - kt)];
      while (kb_flag_acquire(f) != epoch) __nanosleep(64);
    }
    __syncthreads();

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
      kb_cp_async_wait<KB_STAGES - 2>();
      __syncthreads();
      {
        const int cn = c + KB_STAGES - 1;
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
          while (kb_flag_acquire(&fl[tkc * NT + tkc]) != epoch) __nanosleep(64);
        __syncthreads();
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
        // ---- diagonal tile: unblocked POTRF in shared memory ---------------
        for (int j = 0; j < KB_TILE; ++j) {
          __syncthreads();
          double dj = Cs[j][j];
          if (!(dj > 0.0)) {   // not positive definite (or NaN): flag, keep the batch alive
            if (tid == 0) atomicCAS(&info[b], 0, tkc * KB_TILE + j + 1);
            dj = 1.0;
          }
          const double sj = sqrt(dj), inv = 1.0 / sj;
          __syncthreads();
          if (tid < KB_TILE) {
            if (tid > j) Cs[tid][j] *= inv;
            else if (tid == j) Cs[j][j] = sj;
          }
          __syncthreads();
          const int nel = (KB_TILE - 1 - j) * KB_TILE;
          for (int e = tid; e < nel; e += KB_THREADS) {
            const int r = j + 1 + (e >> 6), c = e & 63;
            if (c > j && c <= r) Cs[r][c] -= Cs[r][j] * Cs[c][j];
          }
        }
        __syncthreads();
        // ---- inverse of the 64x64 lower factor (row-by-row elimination) ----
        for (int e = tid; e < KB_TILE * KB_TILE; e += KB_THREADS) {
          const int r = e >> 6, c = e & 63;
          Ds[r][c] = (r == c) ? 1.0 : 0.0;
        }
        for (int j = 0; j < KB_TILE; ++j) {
          __syncthreads();
          const double inv = 1.0 / Cs[j][j];
          if (tid <= j) Ds[j][tid] *= inv;
          __syncthreads();
          const int nel = (KB_TILE - 1 - j) * KB_TILE;
          for (int e = tid; e < nel; e += KB_THREADS) {
            const int r = j + 1 + (e >> 6), c = e & 63;
            if (c <= j) Ds[r][c] -= Cs[r][j] * Ds[j][c];
          }
        }
        __syncthreads();
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
    if (tid == 0) {
      __threadfence();
      kb_flag_release(&fl[ti * NT + tkc], epoch);
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
  kb_chol_kernel<<<grid, KB_THREADS, smem, st>>>(g, aug, dinv, flags, epoch, tasks, ntasks, counter, info);
  return cudaGetLastError();
}
