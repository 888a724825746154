// Counter-based Philox4x32-10 (Salmon et al. 2011) and the uniform / normal variates built on it.
// Shared by the population generator (kb_rng.cu) and the device CMA-ES (kb_cmaes.cu): site index i,
// dimension pair j -> one block of four 32-bit words -> two uniforms -> two variates.
#pragma once
#include <cstdint>

__device__ __forceinline__ void kb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1, uint32_t (&out)[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 52-bit uniform strictly inside (0, 1): (((hi << 20) | (lo >> 12)) + 0.5) * 2^-52  (exact in FP64)
__device__ __forceinline__ double kb_u53(uint32_t hi, uint32_t lo) {
  const unsigned long long v = ((unsigned long long)hi << 20) | (unsigned long long)(lo >> 12);
  return ((double)v + 0.5) * 2.220446049250313e-16;
}

// The two standard normal variates of (site gi, dimension pair j): Box-Muller on the two uniforms,
// first variate for dimension 2j, second for 2j + 1 (the mapping kb_points_generate uses).
__device__ __forceinline__ void kb_philox_normal_pair(unsigned long long gi, int j, unsigned long long seed, double& v0,
                                                      double& v1) {
  uint32_t r[4];
  kb_philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)j, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double u1 = kb_u53(r[0], r[1]), u2 = kb_u53(r[2], r[3]);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  v0 = rad * cs; v1 = rad * sn;
}
