// 2-objective expected hypervolume improvement over a batch of candidates
// (reference optim_tools/ehvi/exi2d.py:6-83, exipsi.py:4-14, gaussfcn.py:5-31, hvolume2d.py:3-13).
//
// The reference walks the (k+1)(k+2)/2 staircase cells of the Pareto front once per candidate and
// recomputes, inside the walk, each cell's bounds and the hypervolume S+ of the front points the cell
// dominates.  None of that depends on the candidate: the host builds it once per front
// (kb_ehvi_build_table) and the kernel keeps it in shared memory.  One thread owns one candidate and
// walks the cells in the reference's (i, j) order, carrying (Phi, phi) of a cell's upper f1 bound over
// as the next cell's lower bound, so a cell costs one erf and one exp.
#include "kb_common.cuh"
#include "kb_internal.h"

#include <algorithm>
#include <vector>

// table layout (doubles): c1[k] | c2[k] | r[2] | splus[(k+1)*(k+1)]
size_t kb_ehvi_table_doubles(int k) { return (size_t)2 * k + 2 + (size_t)(k + 1) * (k + 1); }

static double kb_hvolume2d(std::vector<std::pair<double, double>>& pts, double x0, double x1) {
  // hvolume2d.py:3-13 (stable sort on the first objective)
  std::stable_sort(pts.begin(), pts.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  double h = (x0 - pts[0].first) * (x1 - pts[0].second);
  for (size_t i = 1; i < pts.size(); ++i) h += (x0 - pts[i].first) * (pts[i - 1].second - pts[i].second);
  return h;
}

void kb_ehvi_build_table(int k, const double* front, const double* ref, double* table) {
  double* c1 = table;
  double* c2 = table + k;
  double* r = table + 2 * k;
  double* sp = table + 2 * k + 2;
  for (int i = 0; i < k; ++i) { c1[i] = front[2 * i]; c2[i] = front[2 * i + 1]; }
  std::sort(c1, c1 + k);
  std::sort(c2, c2 + k);
  r[0] = ref[0]; r[1] = ref[1];
  std::vector<std::pair<double, double>> sm;
  for (int i = -1; i < k; ++i) {
    const int ii = (i == -1) ? 0 : i + 1;
    for (int j = -1; j < k; ++j) {
      double v = 0.0;
      if (j < k - ii) {
        const double fmax2 = (j == -1) ? r[1] : c2[k - j - 1];
        const double fmax1 = (i == -1) ? r[0] : c1[k - i - 1];
        const double cU1 = (j == k - 1) ? r[0] : c1[j + 1];
        const double cU2 = (i == k - 1) ? r[1] : c2[i + 1];
        sm.clear();
        for (int m = 0; m < k; ++m)
          if (cU1 <= front[2 * m] && cU2 <= front[2 * m + 1]) sm.emplace_back(front[2 * m], front[2 * m + 1]);
        if (!sm.empty()) v = kb_hvolume2d(sm, fmax1, fmax2);
      }
      sp[(size_t)(i + 1) * (k + 1) + (j + 1)] = v;
    }
  }
}

struct KbEhviBest { double val; long long idx; };

__device__ __forceinline__ void kb_gauss(double z, double& Phi, double& phi) {
  Phi = 0.5 * (1.0 + erf(z * 0.70710678118654752440));
  phi = 0.39894228040143267794 * exp(-0.5 * z * z);
}

__device__ __forceinline__ void kb_ehvi_merge(KbEhviBest& a, double v, long long i) {
  if (v > a.val || (v == a.val && i < a.idx)) { a.val = v; a.idx = i; }
}

template <bool TABLE_IN_SMEM>
__global__ void __launch_bounds__(256) kb_ehvi2d_kernel(long long m, const double* __restrict__ mu1, const double* __restrict__ s1,
                                                         const double* __restrict__ mu2, const double* __restrict__ s2,
                                                         const double* __restrict__ table, int k, long long idx_offset,
                                                         double* __restrict__ out, KbEhviBest* __restrict__ partials) {
  extern __shared__ double sh[];
  const double* T = table;
  if (TABLE_IN_SMEM) {
    const int nt = 2 * k + 2 + (k + 1) * (k + 1);
    for (int i = threadIdx.x; i < nt; i += blockDim.x) sh[i] = table[i];
    __syncthreads();
    T = sh;
  }
  const double* c1 = T;
  const double* c2 = T + k;
  const double r0 = T[2 * k], r1 = T[2 * k + 1];
  const double* sp = T + 2 * k + 2;
  KbEhviBest best{-INFINITY, 0x7fffffffffffffffLL};
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (long long)gridDim.x * blockDim.x) {
    const double m1 = mu1[p], m2 = mu2[p], sd1 = s1[p], sd2 = s2[p];
    double total = 0.0;
    for (int i = -1; i < k; ++i) {
      const int ii = (i == -1) ? 0 : i + 1;
      const double fmax1 = (i == -1) ? r0 : c1[k - i - 1];
      double PhiL2 = 0.0, phiL2 = 0.0, PhiU2, phiU2;
      if (i >= 0) kb_gauss((c2[i] - m2) / sd2, PhiL2, phiL2);
      kb_gauss((((i == k - 1) ? r1 : c2[i + 1]) - m2) / sd2, PhiU2, phiU2);
      const double G2 = PhiU2 - PhiL2;
      const double a2U = sd2 * phiU2, a2L = sd2 * phiL2;
      const double a1 = fmax1 - m1;
      double PhiL1 = 0.0, phiL1 = 0.0;
      const double* sprow = sp + (size_t)(i + 1) * (k + 1);
      for (int j = -1; j < k - ii; ++j) {
        const double fmax2 = (j == -1) ? r1 : c2[k - j - 1];
        const double cU1 = (j == k - 1) ? r0 : c1[j + 1];
        double PhiU1, phiU1;
        kb_gauss((cU1 - m1) / sd1, PhiU1, phiU1);
        const double a2 = fmax2 - m2;
        const double Psi1 = (sd1 * phiU1 + a1 * PhiU1) - ((j == -1) ? 0.0 : (sd1 * phiL1 + a1 * PhiL1));
        const double Psi2 = (a2U + a2 * PhiU2) - ((i == -1) ? 0.0 : (a2L + a2 * PhiL2));
        const double c = Psi1 * Psi2 - sprow[j + 1] * (PhiU1 - PhiL1) * G2;
        total += fmax(c, 0.0);
        PhiL1 = PhiU1; phiL1 = phiU1;
      }
    }
    if (out) out[p] = total;
    kb_ehvi_merge(best, total, idx_offset + p);
  }
  // block argmax (ties -> lowest index)
  for (int o = 16; o > 0; o >>= 1) {
    const double v = __shfl_down_sync(0xffffffffu, best.val, o);
    const long long ix = __shfl_down_sync(0xffffffffu, best.idx, o);
    kb_ehvi_merge(best, v, ix);
  }
  __shared__ KbEhviBest wb[8];
  __syncthreads();
  if ((threadIdx.x & 31) == 0) wb[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) kb_ehvi_merge(best, wb[w].val, wb[w].idx);
    partials[blockIdx.x] = best;
  }
}

__global__ void kb_ehvi_final_kernel(const KbEhviBest* __restrict__ partials, int n, KbEhviBest* __restrict__ result,
                                     int accumulate) {
  KbEhviBest best{-INFINITY, 0x7fffffffffffffffLL};
  for (int i = threadIdx.x; i < n; i += 32) kb_ehvi_merge(best, partials[i].val, partials[i].idx);
  for (int o = 16; o > 0; o >>= 1) {
    const double v = __shfl_down_sync(0xffffffffu, best.val, o);
    const long long ix = __shfl_down_sync(0xffffffffu, best.idx, o);
    kb_ehvi_merge(best, v, ix);
  }
  if (threadIdx.x == 0) {
    if (accumulate) kb_ehvi_merge(best, result->val, result->idx);
    *result = best;
  }
}

int kb_ehvi_grid(long long m, int num_sms) {
  const long long blocks = (m + 255) / 256;
  const long long cap = (long long)num_sms * 8;
  return (int)(blocks < cap ? blocks : cap);
}

// partials must hold kb_ehvi_grid(m, num_sms) entries; result is {double val; long long idx}.
cudaError_t kb_launch_ehvi2d(cudaStream_t st, long long m, const double* mu1, const double* s1, const double* mu2,
                             const double* s2, const double* table_dev, int k, long long idx_offset, double* out,
                             void* partials, void* result, int accumulate, int num_sms) {
  const int grid = kb_ehvi_grid(m, num_sms);
  const size_t smem = kb_ehvi_table_doubles(k) * sizeof(double);
  if (smem <= 96 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kb_ehvi2d_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kb_ehvi2d_kernel<true><<<grid, 256, smem, st>>>(m, mu1, s1, mu2, s2, table_dev, k, idx_offset, out,
                                                   (KbEhviBest*)partials);
  } else {
    kb_ehvi2d_kernel<false><<<grid, 256, 0, st>>>(m, mu1, s1, mu2, s2, table_dev, k, idx_offset, out,
                                                 (KbEhviBest*)partials);
  }
  kb_ehvi_final_kernel<<<1, 32, 0, st>>>((const KbEhviBest*)partials, grid, (KbEhviBest*)result, accumulate);
  return cudaGetLastError();
}
