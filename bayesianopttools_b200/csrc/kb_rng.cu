// krig-b200: Monte-Carlo population generator on the device (mcpopgen, reliability_analysis/akmcs.py:241-260).
// Counter-based Philox4x32-10 (Salmon et al. 2011): no state, site i / dimension pair j -> one block of
// four 32-bit words -> two 53-bit uniforms -> two variates; reproducible on the host (oracle/philox.py).
#include "kb_common.cuh"
#include "kb_internal.h"
#include "kb_philox.cuh"

__global__ void kb_points_generate_kernel(double* __restrict__ x, long long npts, int d, int dist, double p0, double p1,
                                          const double* __restrict__ lb, const double* __restrict__ ub,
                                          unsigned long long seed, long long first_index) {
  const int npair = (d + 1) / 2;
  const long long total = npts * npair;
  double lmu = 0.0, lsig = 0.0, gscale = 0.0;
  if (dist == KB_DIST_LOGNORMAL) {          // akmcs.py:246-250
    const double var = p1 * p1;
    lsig = sqrt(log(var / (p0 * p0) + 1.0));
    lmu = log((p0 * p0) / sqrt(var + p0 * p0));
  } else if (dist == KB_DIST_GUMBEL) {      // akmcs.py:251-253
    gscale = (p1 / 3.14159265358979323846) * sqrt(6.0);
  }
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / npair;
    const int j = (int)(e - i * npair);
    const unsigned long long gi = (unsigned long long)(i + first_index);
    uint32_t r[4];
    kb_philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)j, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u1 = kb_u53(r[0], r[1]), u2 = kb_u53(r[2], r[3]);
    double v0, v1;
    if (dist == KB_DIST_UNIFORM) {
      v0 = u1; v1 = u2;
    } else if (dist == KB_DIST_GUMBEL) {
      v0 = p0 - gscale * log(-log(u1));
      v1 = p0 - gscale * log(-log(u2));
    } else {                                // Box-Muller
      const double rad = sqrt(-2.0 * log(u1));
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      v0 = rad * cs; v1 = rad * sn;
      if (dist == KB_DIST_NORMAL) { v0 = p1 * v0 + p0; v1 = p1 * v1 + p0; }
      else { v0 = exp(lsig * v0 + lmu); v1 = exp(lsig * v1 + lmu); }
    }
    const int d0 = 2 * j, d1 = 2 * j + 1;
    if (dist == KB_DIST_UNIFORM) {
      v0 = lb[d0] + (ub[d0] - lb[d0]) * v0;
      if (d1 < d) v1 = lb[d1] + (ub[d1] - lb[d1]) * v1;
    }
    x[(size_t)i * d + d0] = v0;
    if (d1 < d) x[(size_t)i * d + d1] = v1;
  }
}

cudaError_t kb_launch_points_generate(cudaStream_t st, double* x, long long npts, int d, int dist, double p0, double p1,
                                      const double* lb_dev, const double* ub_dev, unsigned long long seed,
                                      long long first_index) {
  kb_points_generate_kernel<<<148 * 8, 256, 0, st>>>(x, npts, d, dist, p0, p1, lb_dev, ub_dev, seed, first_index);
  return cudaGetLastError();
}


// ---- low-discrepancy designs on the device (misc/sampling: halton, sobolnew) -------------------------
// Site i of either sequence is a pure function of i, so shards are generated independently
// (first_index) and any subset is reproducible on the host bit for bit:
//  * Halton (haltonsampling.py:30-40): radical inverse of i in the k-th prime base, accumulated in the
//    reference's order (f /= base; r += f * digit) with IEEE division;
//  * Sobol (sobol_new.py:5-98, Joe & Kuo direction numbers): x_i = XOR of v[b + 1] over the set bits b of
//    the Gray code i ^ (i >> 1), the closed form of the reference's cumulative XOR; the 32 direction
//    numbers per dimension come from the host (they depend on the Joe-Kuo table, which the library does
//    not embed).
// Both are mapped to [lb, ub] as realval does: u * (ub - lb) + lb with separate roundings (no FMA).
__global__ void kb_points_lowdisc_kernel(double* __restrict__ x, long long npts, int d, int kind,
                                         const unsigned int* __restrict__ dirs, const int* __restrict__ primes,
                                         const double* __restrict__ lb, const double* __restrict__ ub,
                                         long long first_index) {
  const long long total = npts * d;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / d;
    const int j = (int)(e - i * d);
    const unsigned long long gi = (unsigned long long)(i + first_index);
    double u;
    if (kind == KB_DIST_HALTON) {
      const unsigned long long base = (unsigned long long)primes[j];
      const double fb = (double)base;
      double f = 1.0, r = 0.0;
      unsigned long long k = gi;
      while (k > 0) {
        const unsigned long long q = k / base, rem = k - q * base;
        k = q;
        f = __ddiv_rn(f, fb);
        r = __dadd_rn(r, __dmul_rn(f, (double)rem));
      }
      u = r;
    } else {
      unsigned long long g = gi ^ (gi >> 1);
      unsigned int acc = 0;
      const unsigned int* v = dirs + (size_t)j * 32;
      while (g) {
        const int b = __ffsll((long long)g) - 1;
        acc ^= v[b];
        g &= g - 1;
      }
      u = (double)acc * 2.3283064365386963e-10;          // / 2^32, exact
    }
    x[e] = lb ? __dadd_rn(__dmul_rn(u, __dsub_rn(ub[j], lb[j])), lb[j]) : u;
  }
}

cudaError_t kb_launch_points_lowdisc(cudaStream_t st, double* x, long long npts, int d, int kind,
                                     const unsigned int* dirs_dev, const int* primes_dev, const double* lb_dev,
                                     const double* ub_dev, long long first_index) {
  kb_points_lowdisc_kernel<<<148 * 8, 256, 0, st>>>(x, npts, d, kind, dirs_dev, primes_dev, lb_dev, ub_dev, first_index);
  return cudaGetLastError();
}
