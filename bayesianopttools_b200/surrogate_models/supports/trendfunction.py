"""Trend basis on the host (reference: surrogate_models/supports/trendfunction.py).

F is an INPUT of the kernels (n x nind, built once per standardize()); prediction sites
evaluate the same orthonormal Legendre basis on the device (csrc/kb_predict.cu).
"""
from itertools import combinations

import numpy as np


def polytruncation(nix, nvar, q=1):
    """Total-order multi-indices in the reference's order (trendfunction.py:7-41)."""
    blocks = []
    for dgr in range(nix, -1, -1):
        rows = []
        for c in combinations(range(1, dgr + nvar), nvar - 1):
            cuts = (0,) + c + (dgr + nvar,)
            rows.append([cuts[i + 1] - cuts[i] - 1 for i in range(nvar)])
        blocks.append(np.array(rows, dtype=float).reshape(-1, nvar))
    idx = np.flip(np.vstack(blocks), 0)
    if q < 1:
        idp = np.sum(idx ** q, 1) ** (1.0 / q)
        idx = idx[idp <= (nix + 0.000001), :]
    return idx


def compute_regression_mat(idx, X, bounds, polytype=None, **kwargs):
    """Orthonormal Legendre regression matrix (trendfunction.py:70-107, polytype == 1)."""
    X = np.asarray(X, dtype=float)
    idx = np.asarray(idx).astype(int)
    bounds = np.asarray(bounds, dtype=float)
    if polytype is not None and np.any(np.asarray(polytype) != 1):
        raise NotImplementedError("only the Legendre basis (polytype 1) is on the Kriging path")
    N = 2 * ((X - bounds[0]) / (bounds[1] - bounds[0])) - 1
    maxo = int(idx.max()) if idx.size else 0
    P = np.empty((maxo + 1,) + N.shape)
    P[0] = 1.0
    if maxo >= 1:
        P[1] = N
    for k in range(1, maxo):
        P[k + 1] = ((2 * k + 1) * N * P[k] - k * P[k - 1]) / (k + 1)
    for k in range(maxo + 1):
        P[k] = P[k] * np.sqrt(2 * k + 1)
    F = np.ones((N.shape[0], idx.shape[0]))
    for a in range(idx.shape[0]):
        for j in range(idx.shape[1]):
            if idx[a, j]:
                F[:, a] *= P[idx[a, j], :, j]
    return F
