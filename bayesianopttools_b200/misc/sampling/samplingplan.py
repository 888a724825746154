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


_JK = None


def _joekuo():
    global _JK
    if _JK is None:
        _JK = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sobol_joekuo_256.npy"))
    return _JK


def sobol_direction_numbers(d, L=32):
    """The first L 32-bit direction numbers v_1..v_L (v_k = m_k << (32 - k)) of the first d Joe & Kuo
    dimensions, shape (d, L) uint64 (sobol_new.py:44-80).  Also the table handed to the device generator
    (`_lib.DevicePoints.generate(dist='sobol')`)."""
    JK = _joekuo()
    if d - 1 > len(JK):
        raise ValueError("sobolnew supports up to %d dimensions" % (len(JK) + 1))
    V = np.zeros((d, L), dtype=np.uint64)
    for j in range(d):
        v = [0] * (L + 1)
        if j == 0:
            for i in range(1, L + 1):
                v[i] = 1 << (32 - i)
        else:
            s_, a_ = int(JK[j - 1, 0]), int(JK[j - 1, 1])
            m = [0] + [int(x) for x in JK[j - 1, 2:2 + s_]]
            for i in range(1, min(L, s_) + 1):
                v[i] = m[i] << (32 - i)
            for i in range(s_ + 1, L + 1):
                v[i] = v[i - s_] ^ (v[i - s_] >> s_)
                for k in range(1, s_):
                    v[i] ^= ((a_ >> (s_ - 1 - k)) & 1) * v[i - k]
        V[j] = v[1:]
    return V


def sobol_points(n, d):
    """Joe & Kuo Sobol points in Gray-code order, first point 0 (reference: misc/sampling/sobol_new.py:5-98,
    the `sobolnew` option; direction numbers new-joe-kuo-6.21201, first 257 dimensions packed in
    sobol_joekuo_256.npy by the table script next to the golden-fixture generator).

    x_0 = 0, x_i = x_{i-1} XOR v[c_{i-1}] with c_i = 1 + index of the lowest zero bit of i and the
    32-bit direction numbers v_k = m_k << (32 - k), m_k from the primitive-polynomial recurrence
    m_k = m_{k-s} XOR (m_{k-s} << s ... ) (Joe & Kuo 2008)."""
    n, d = int(n), int(d)
    L = max(int(np.ceil(np.log(n) / np.log(2.0))), 1)
    idx = np.arange(n, dtype=np.int64)
    c = np.ones(n, dtype=np.int64)              # 1 + number of trailing one bits of i
    t = idx.copy()
    while True:
        odd = (t & 1) == 1
        if not odd.any():
            break
        c[odd] += 1
        t[odd] >>= 1
    pts = np.zeros((n, d))
    V = sobol_direction_numbers(d, L)
    for j in range(d):
        v = np.concatenate(([0], V[j], [0])).astype(np.int64)
        # x_i = XOR of v[c_0..c_{i-1}]: cumulative XOR along the sequence
        steps = v[c[:-1]] if n > 1 else np.zeros(0, dtype=np.int64)
        x = np.zeros(n, dtype=np.int64)
        if n > 1:
            x[1:] = np.bitwise_xor.accumulate(steps)
        pts[:, j] = x / float(2 ** 32)
    return pts


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
    elif opt == "sobol":
        samplenorm = sobol(nvar, nsamp)
    elif opt == "sobolnew":
        samplenorm = sobol_points(nsamp + 1, nvar)[1:, :]      # the zero point is dropped (samplingplan.py:24-25)
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
