"""Host-side sampling plans and standardisation (reference: misc/sampling/samplingplan.py).

Only what the Kriging hot path and its callers need: ``sampling`` (halton / sobol / rlh),
``realval`` and ``standardize``.  Sample sets are inputs to the kernels, not part of them.
"""
import os

import numpy as np

_DIRS = None


def _sobol_dirs():
    global _DIRS
    if _DIRS is None:
        _DIRS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sobol_dirs_toms659.npy"))
    return _DIRS


def sobol(nvar, nsamp):
    """Gray-code Sobol points 1..nsamp with the TOMS-659 direction numbers
    (same sequence as the reference's i4_sobol_generate, sobol_seq.py:118-134)."""
    V = _sobol_dirs()
    if nvar > V.shape[0]:
        raise ValueError("sobol start points support at most %d dimensions" % V.shape[0])
    x = np.zeros(nvar, dtype=np.int64)
    out = np.empty((nsamp, nvar))
    for j in range(1, nsamp + 1):
        c, t = 0, j - 1
        while t & 1:
            t >>= 1
            c += 1
        x ^= V[:nvar, c]
        out[j - 1] = x * (2.0 ** -30)
    return out


def _primes(k):
    out, c = [], 2
    while len(out) < k:
        if all(c % p for p in out):
            out.append(c)
        c += 1
    return out


def halton(nvar, nsamp):
    """Van der Corput radical inverse per prime base, first point 0 (haltonsampling.py:30-40)."""
    out = np.zeros((nsamp, nvar))
    for j, base in enumerate(_primes(nvar)):
        for i in range(nsamp):
            f, r, k = 1.0, 0.0, i
            while k > 0:
                k, rem = divmod(k, base)
                f = f / base
                r = r + f * rem
            out[i, j] = r
    return out


def rlh(nvar, nsamp, rng=None):
    """Random Latin hypercube (reference: misc/sampling/rlh.py)."""
    rng = np.random if rng is None else rng
    out = np.empty((nsamp, nvar))
    for j in range(nvar):
        out[:, j] = (rng.permutation(nsamp) + 0.5) / nsamp
    return out


def realval(lb, ub, samp):
    """samplingplan.py:46-57"""
    lb, ub = np.asarray(lb, dtype=float), np.asarray(ub, dtype=float)
    samp = np.asarray(samp, dtype=float)
    if len(ub) != len(lb):
        raise ValueError("Lower and upper bound have to be in the same dimension")
    if len(ub) != np.size(samp, axis=1):
        raise ValueError("sample and bound are not in the same dimension")
    return samp * (ub - lb) + lb


def sampling(option, nvar, nsamp, **kwargs):
    """samplingplan.py:14-44 -- returns (normalised, real) samples."""
    ret = kwargs.get("result", "normalized")
    ub = kwargs.get("upbound", np.array([None]))
    lb = kwargs.get("lobound", np.array([None]))
    opt = option.lower()
    if opt == "halton":
        samplenorm = halton(nvar, nsamp)
    elif opt in ("sobol", "sobolnew"):
        samplenorm = sobol(nvar, nsamp)
    elif opt == "rlh":
        samplenorm = rlh(nvar, nsamp)
    else:
        raise NameError("sampling plan unavailable!")
    if ret.lower() == "real":
        if np.asarray(lb).any() is None or np.asarray(ub).any() is None:
            raise ValueError("lb and ub must have value")
        if np.any(np.asarray(ub, float) - np.asarray(lb, float) < 0):
            raise ValueError("Upper bound must bigger than lower bound!")
        return samplenorm, realval(lb, ub, samplenorm)
    return samplenorm, np.zeros(np.shape(samplenorm))


def standardize(X, y=None, type="default", norm_y=False, **kwargs):
    """samplingplan.py:59-108.  'default': x -> ((x-lb)/(ub-lb) - 0.5)*2 ; 'std': z-score."""
    ranges = kwargs.get("range", None)
    X = np.asarray(X, dtype=float)
    if type.lower() == "default":
        if ranges is None:
            raise ValueError("Default normalization requires range value!")
        ranges = np.asarray(ranges, dtype=float)
        nv = X.shape[1]
        X_norm = ((X - ranges[0, :nv]) / (ranges[1, :nv] - ranges[0, :nv]) - 0.5) * 2
        if norm_y:
            y = np.asarray(y, dtype=float)
            y_norm = ((y - ranges[0, nv:]) / (ranges[1, nv:] - ranges[0, nv:]) - 0.5) * 2
            return X_norm, y_norm
        return X_norm
    if type.lower() == "std":
        X_mean = np.mean(X, axis=0)
        X_std = X.std(axis=0, ddof=1)
        X_std[X_std == 0.0] = 1.0
        X_norm = (X - X_mean) / X_std
        if norm_y:
            y = np.asarray(y, dtype=float)
            y_mean = np.mean(y, axis=0)
            y_std = y.std(axis=0, ddof=1)
            y_std[y_std == 0.0] = 1.0
            return X_norm, (y - y_mean) / y_std, X_mean, y_mean, X_std, y_std
        return X_norm, X_mean, X_std
    raise ValueError("standardization type not recognised")
