// krig-b200: (mu/mu_w, lambda)-CMA-ES whose whole generation loop stays on the device
// (replaces the host loop around `cma.fmin2(likelihood, ...)`,
// surrogate_models/kriging_model.py:420-424 -- pycma is a third-party dependency of the reference).
//
// One generation = kb_cmaes_step_kernel (rank the population that was just evaluated, update mean /
// paths / covariance / step size, eigen-decompose, test the stop criteria, sample the next population)
// followed by the likelihood pipeline on the sampled population.  The host only enqueues generations
// and reads a 4-byte stop flag every few of them: no per-generation round trip.
//
// Differences from a textbook implementation, chosen so that a host restatement (oracle/cmaes_oracle.py)
// reproduces the trajectory without depending on eigenvector signs or ordering:
//  * samples are y = C^(1/2) z with the SYMMETRIC square root A = B diag(D) B^T, hence the conjugate
//    evolution path uses C^(-1/2) y_w = z_w directly;
//  * z comes from the counter-based Philox stream: generation g, individual i is site g * lambda + i of
//    the standard normal population kb_points_generate(seed) would produce.
// Box handling as optim_tools/cmaes.py: candidates are clipped into the box for evaluation and a
// quadratic distance penalty is added; non-finite objective values (R not positive definite) rank last.
#include "kb_common.cuh"
#include "kb_internal.h"
#include "kb_philox.cuh"

#define KB_CMA_THREADS 256

// Parallel-order cyclic Jacobi on the symmetric S (N x N, leading dimension ld, shared memory):
// on return diag(S) = eigenvalues, V = eigenvectors (columns).  M/2 disjoint rotations per round.
// Run by ONE warp when N <= 32 (warp barriers only), else by the whole block; converged when the
// largest off-diagonal entry is at rounding level of the largest diagonal one (quadratic convergence:
// 6-8 sweeps for N ~ 10).
template <bool WARP>
__device__ void kb_cma_jacobi(double* S, double* V, int N, int ld, double* rot, int* rpq) {
  const int tid = threadIdx.x, nthr = WARP ? 32 : KB_CMA_THREADS;
  auto sync = [] { if (WARP) __syncwarp(); else __syncthreads(); };
  for (int e = tid; e < N * N; e += nthr) V[(e / N) * ld + (e % N)] = (e / N == e % N) ? 1.0 : 0.0;
  sync();
  if (N < 2) return;
  const int M = (N + 1) & ~1, half = M / 2;
  for (int sweep = 0; sweep < 40; ++sweep) {
    // every thread scans the whole matrix: N <= 64, and it keeps the exit test barrier-free
    double off = 0.0, dia = 0.0;
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < N; ++b) {
        const double v = fabs(S[a * ld + b]);
        if (a == b) dia = fmax(dia, v); else off = fmax(off, v);
      }
    if (off <= 2.3e-16 * dia || off < 1e-300) break;
    sync();
    for (int r = 0; r < M - 1; ++r) {
      // round-robin pairing: (M-1, r) and ((r+k) mod (M-1), (r-k) mod (M-1))
      if (tid < half) {
        int p, q;
        if (tid == 0) { p = M - 1; q = r; }
        else { p = (r + tid) % (M - 1); q = (r - tid + (M - 1)) % (M - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < N) {
          const double apq = S[p * ld + q];
          if (fabs(apq) > 1e-300) {
            const double tau = (S[q * ld + q] - S[p * ld + p]) / (2.0 * apq);
            const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        }
        rot[2 * tid] = c; rot[2 * tid + 1] = s;
        rpq[2 * tid] = p; rpq[2 * tid + 1] = q;
      }
      sync();
      for (int e = tid; e < half * N; e += nthr) {          // rows p, q of J^T S
        const int pr = e / N, j = e % N;
        const int p = rpq[2 * pr], q = rpq[2 * pr + 1];
        if (q >= N) continue;
        const double c = rot[2 * pr], s = rot[2 * pr + 1];
        const double a = S[p * ld + j], b = S[q * ld + j];
        S[p * ld + j] = c * a - s * b;
        S[q * ld + j] = s * a + c * b;
      }
      sync();
      for (int e = tid; e < half * N; e += nthr) {          // columns p, q of (.) J, and V J
        const int pr = e / N, i = e % N;
        const int p = rpq[2 * pr], q = rpq[2 * pr + 1];
        if (q >= N) continue;
        const double c = rot[2 * pr], s = rot[2 * pr + 1];
        double a = S[i * ld + p], b = S[i * ld + q];
        S[i * ld + p] = c * a - s * b;
        S[i * ld + q] = s * a + c * b;
        a = V[i * ld + p]; b = V[i * ld + q];
        V[i * ld + p] = c * a - s * b;
        V[i * ld + q] = s * a + c * b;
      }
      sync();
    }
  }
}

size_t kb_cmaes_state_doubles(int N, int lam, int mu, int maxiter) { return kb_cma_view(nullptr, N, lam, mu, maxiter, nullptr); }

// tell (skipped when first != 0) + stop test + ask.  f: objective of the population sampled by the
// previous call; xpop: the clipped population to evaluate next (lam x nbhyp).
__global__ void __launch_bounds__(KB_CMA_THREADS) kb_cmaes_step_kernel(KbCmaParams P, double* __restrict__ state,
                                                                       const double* __restrict__ f_in,
                                                                       double* __restrict__ xpop, int first) {
  extern __shared__ double sh[];
  const int tid = threadIdx.x, N = P.N, lam = P.lam, mu = P.mu, ld = N + 1;
  KbCmaView S;
  kb_cma_view(state, N, lam, mu, P.maxiter, &S);
  double* sC = sh;                       // N x ld
  double* sV = sC + N * ld;              // N x ld
  double* fpen = sV + N * ld;            // lam
  double* yw = fpen + lam;               // N
  double* zw = yw + N;                   // N
  double* rot = zw + N;                  // (c, s) of up to 64 simultaneous rotations
  double* red = rot + 2 * 64;            // 8
  int* order = (int*)(red + 8);          // lam
  int* rpq = order + lam;                // (p, q) of the rotations
  __shared__ double s_misc[4];
  if (S.scal[3] != 0.0) return;          // stopped: the state is frozen (uniform branch)

  if (!first) {
    // ---- tell -------------------------------------------------------------------------------------
    // worst finite value, best finite individual (lowest index on ties)
    if (tid == 0) {
      double worst = -INFINITY, bestf = INFINITY;
      int ib = -1, any = 0;
      for (int i = 0; i < lam; ++i) {
        const double v = f_in[i];
        if (isfinite(v)) {
          any = 1;
          worst = fmax(worst, v);
          if (v < bestf) { bestf = v; ib = i; }
        }
      }
      if (!any) worst = 0.0;
      s_misc[0] = worst; s_misc[1] = bestf; s_misc[2] = (double)ib;
    }
    __syncthreads();
    const double worst = s_misc[0];
    const int ib = (int)s_misc[2];
    if (ib >= 0 && s_misc[1] < S.scal[1]) {
      // uniform condition; every thread sees the same S.scal[1] until the barrier below
      for (int a = tid; a < N; a += KB_CMA_THREADS) {
        double xv = S.x[ib * N + a];
        if (P.has_bounds) xv = fmin(fmax(xv, S.lb[a]), S.ub[a]);
        S.best_x[a] = xv;
      }
    }
    __syncthreads();
    if (tid == 0 && ib >= 0 && s_misc[1] < S.scal[1]) S.scal[1] = s_misc[1];
    for (int i = tid; i < lam; i += KB_CMA_THREADS) {
      double v = f_in[i];
      if (!isfinite(v)) v = worst + 1e3 * (1.0 + fabs(worst));
      if (P.has_bounds) {
        double pen = 0.0;
        for (int a = 0; a < N; ++a) {
          const double xv = S.x[i * N + a];
          const double xf = fmin(fmax(xv, S.lb[a]), S.ub[a]);
          const double r = (xv - xf) / fmax(S.ub[a] - S.lb[a], 1e-300);
          pen += r * r;
        }
        v += (1.0 + fabs(worst)) * pen;
      }
      fpen[i] = v;
    }
    __syncthreads();
    // stable ranking by counting
    for (int i = tid; i < lam; i += KB_CMA_THREADS) {
      const double v = fpen[i];
      int rk = 0;
      for (int j = 0; j < lam; ++j) rk += (fpen[j] < v || (fpen[j] == v && j < i)) ? 1 : 0;
      order[rk] = i;
    }
    __syncthreads();
    for (int a = tid; a < N; a += KB_CMA_THREADS) {
      double sy = 0.0, sz = 0.0;
      for (int r = 0; r < mu; ++r) {
        sy += S.w[r] * S.y[order[r] * N + a];
        sz += S.w[r] * S.z[order[r] * N + a];
      }
      yw[a] = sy; zw[a] = sz;
    }
    __syncthreads();
    const double sigma = S.scal[0];
    const int gen = (int)S.scal[2] + 1;
    for (int a = tid; a < N; a += KB_CMA_THREADS) {
      S.mean[a] = S.mean[a] + sigma * yw[a] * S.scale[a];
      S.ps[a] = (1.0 - P.cs) * S.ps[a] + sqrt(P.cs * (2.0 - P.cs) * P.mueff) * zw[a];
    }
    __syncthreads();
    if (tid == 0) {
      double n2 = 0.0;
      for (int a = 0; a < N; ++a) n2 += S.ps[a] * S.ps[a];
      const double nps = sqrt(n2);
      const double hs = (nps / sqrt(1.0 - pow(1.0 - P.cs, 2.0 * gen)) / P.chiN) < (1.4 + 2.0 / (N + 1)) ? 1.0 : 0.0;
      s_misc[0] = nps; s_misc[1] = hs;
    }
    __syncthreads();
    const double nps = s_misc[0], hsig = s_misc[1];
    for (int a = tid; a < N; a += KB_CMA_THREADS)
      S.pc[a] = (1.0 - P.cc) * S.pc[a] + hsig * sqrt(P.cc * (2.0 - P.cc) * P.mueff) * yw[a];
    __syncthreads();
    // covariance update on the upper triangle, mirrored
    for (int e = tid; e < N * N; e += KB_CMA_THREADS) {
      const int a = e / N, b = e % N;
      if (b < a) continue;
      const double cab = S.C[a * N + b];
      double rk = 0.0;
      for (int r = 0; r < mu; ++r) rk += S.w[r] * S.y[order[r] * N + a] * S.y[order[r] * N + b];
      const double v = (1.0 - P.c1 - P.cmu) * cab +
                       P.c1 * (S.pc[a] * S.pc[b] + (1.0 - hsig) * P.cc * (2.0 - P.cc) * cab) + P.cmu * rk;
      sC[a * ld + b] = v; sC[b * ld + a] = v;
    }
    __syncthreads();
    for (int e = tid; e < N * N; e += KB_CMA_THREADS) S.C[e] = sC[(e / N) * ld + (e % N)];
    __syncthreads();
    if (N <= 32) { if (tid < 32) kb_cma_jacobi<true>(sC, sV, N, ld, rot, rpq); }
    else kb_cma_jacobi<false>(sC, sV, N, ld, rot, rpq);
    __syncthreads();
    for (int a = tid; a < N; a += KB_CMA_THREADS) S.D[a] = sqrt(fmax(sC[a * ld + a], 1e-300));
    __syncthreads();
    for (int e = tid; e < N * N; e += KB_CMA_THREADS) {
      const int a = e / N, b = e % N;
      double v = 0.0;
      for (int k = 0; k < N; ++k) v += sV[a * ld + k] * S.D[k] * sV[b * ld + k];
      S.A[e] = v;
    }
    if (tid == 0) {
      const double sig_new = sigma * exp((P.cs / P.damps) * (nps / P.chiN - 1.0));
      double fmin_pen = INFINITY;
      for (int i = 0; i < lam; ++i) fmin_pen = fmin(fmin_pen, fpen[i]);
      S.hist[gen - 1] = fmin_pen;
      // stop criteria (optim_tools/cmaes.py:83-94)
      double dmax = 0.0, dmin = INFINITY, smax = 0.0;
      for (int a = 0; a < N; ++a) {
        dmax = fmax(dmax, S.D[a]); dmin = fmin(dmin, S.D[a]); smax = fmax(smax, S.scale[a]);
      }
      int stop = 0;
      if (gen >= P.maxiter) stop = 1;
      else if (sig_new * dmax * smax < P.tolx) stop = 2;
      else if (dmax > 1e7 * dmin) stop = 4;
      else if (gen > P.window) {
        double hi = -INFINITY, lo = INFINITY;
        for (int k = gen - P.window; k < gen; ++k) { hi = fmax(hi, S.hist[k]); lo = fmin(lo, S.hist[k]); }
        if (hi - lo < P.tolfun) stop = 3;
      }
      S.scal[0] = sig_new;
      S.scal[2] = (double)gen;
      S.scal[3] = (double)stop;
      s_misc[3] = (double)stop;
    }
    __syncthreads();
    if (s_misc[3] != 0.0) return;
  }
  __threadfence_block();
  __syncthreads();
  // ---- ask: the next population --------------------------------------------------------------------
  const double sigma = S.scal[0];
  const unsigned long long gbase = (unsigned long long)S.scal[2] * (unsigned long long)lam;
  const int npair = (N + 1) / 2;
  for (int e = tid; e < lam * npair; e += KB_CMA_THREADS) {
    const int i = e / npair, j = e % npair;
    double v0, v1;
    kb_philox_normal_pair(gbase + i, j, P.seed, v0, v1);
    S.z[i * N + 2 * j] = v0;
    if (2 * j + 1 < N) S.z[i * N + 2 * j + 1] = v1;
  }
  __syncthreads();
  for (int e = tid; e < lam * N; e += KB_CMA_THREADS) {
    const int i = e / N, a = e % N;
    double v = 0.0;
    for (int k = 0; k < N; ++k) v += S.A[a * N + k] * S.z[i * N + k];
    S.y[e] = v;
    const double xv = S.mean[a] + sigma * v * S.scale[a];
    S.x[e] = xv;
    xpop[i * P.nbhyp + a] = P.has_bounds ? fmin(fmax(xv, S.lb[a]), S.ub[a]) : xv;
  }
  if (tid == 0) S.scal[4] += (double)lam;
}

size_t kb_cmaes_step_smem(int N, int lam) {
  const int ld = N + 1;
  return sizeof(double) * ((size_t)2 * N * ld + lam + 2 * N + 2 * 64 + 8) + sizeof(int) * (lam + 2 * 64) + 16;
}

cudaError_t kb_launch_cmaes_step(cudaStream_t st, const KbCmaParams& P, double* state, const double* f_in, double* xpop,
                                 int first) {
  const size_t smem = kb_cmaes_step_smem(P.N, P.lam);
  cudaError_t e = cudaFuncSetAttribute(kb_cmaes_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kb_cmaes_step_kernel<<<1, KB_CMA_THREADS, smem, st>>>(P, state, f_in, xpop, first);
  return cudaGetLastError();
}
