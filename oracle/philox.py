"""Host restatement (numpy) of the device Monte-Carlo population generator (csrc/kb_rng.cu):
Philox4x32-10 keyed by the seed, counter = (site index, dimension pair), two 52-bit uniforms per block,
Box-Muller / inverse transforms as in mcpopgen (reliability_analysis/akmcs.py:241-260).
TEST INFRASTRUCTURE ONLY (parity of sub-sets of a device-generated population)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & MASK, np.uint64(k1) & MASK
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def u52(hi, lo):
    v = (hi << np.uint64(20)) | (lo >> np.uint64(12))
    return (v.astype(np.float64) + 0.5) * 2.0 ** -52


def population(indices, d, dist="normal", p0=0.0, p1=1.0, lb=None, ub=None, seed=0):
    """Rows `indices` (global site indices) of the population kb_points_generate builds."""
    idx = np.asarray(indices, dtype=np.uint64)
    npair = (d + 1) // 2
    out = np.empty((len(idx), 2 * npair))
    for j in range(npair):
        r = philox4x32_10(idx & MASK, idx >> np.uint64(32), np.full(len(idx), j, np.uint64), np.zeros(len(idx), np.uint64),
                          np.uint64(seed) & MASK, np.uint64(seed) >> np.uint64(32))
        u1, u2 = u52(r[0], r[1]), u52(r[2], r[3])
        if dist == "uniform":
            v0, v1 = u1, u2
        elif dist == "gumbel":
            sc = (p1 / np.pi) * np.sqrt(6.0)
            v0, v1 = p0 - sc * np.log(-np.log(u1)), p0 - sc * np.log(-np.log(u2))
        else:
            rad = np.sqrt(-2.0 * np.log(u1))
            v0, v1 = rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)
            if dist == "normal":
                v0, v1 = p1 * v0 + p0, p1 * v1 + p0
            else:
                var = p1 ** 2
                sig = np.sqrt(np.log(var / p0 ** 2 + 1))
                mu = np.log(p0 ** 2 / np.sqrt(var + p0 ** 2))
                v0, v1 = np.exp(sig * v0 + mu), np.exp(sig * v1 + mu)
        out[:, 2 * j], out[:, 2 * j + 1] = v0, v1
    out = out[:, :d]
    if dist == "uniform":
        out = np.asarray(lb) + (np.asarray(ub) - np.asarray(lb)) * out
    return out
