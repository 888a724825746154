"""CPU oracle: a numpy/scipy restatement of KADAL's Kriging hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the shipped package
(``bayesianopttools_b200``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
there only as the checker or the timed CPU baseline.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference
(imported from ``/root/reference``) on seeded inputs and on the notebook inputs
whose outputs the reference stores (``Kriging_Demo.ipynb:341-345``,
``Single_Objective_Bayesian_Optimization.ipynb:307-310``); the resulting
fixtures live in ``tests/golden/`` and ``tests/test_oracle_golden.py`` checks
every function below against them.  Paths with no reference behaviour to pin
(prediction variance with ``TrendOrder > 0`` -- the reference crashes, see
SURVEY.md App. B) say so at the function.

Every function cites the reference lines it restates (paths relative to the
reference root).
"""
import numpy as np
from scipy.linalg import cholesky as _chol, solve_triangular as _trsv
from scipy.special import erf as _erf

KERNEL_IDS = {"gaussian": 0, "exponential": 1, "matern32": 2, "matern52": 3,
              "iso_gaussian": 0}


# --------------------------------------------------------------------------
# correlation  (surrogate_models/supports/kernelfunc.py:4-82)
# --------------------------------------------------------------------------
def corr(XN, XM, theta, kernel="gaussian", plscoeff=None):
    """Correlation matrix (n, m) between rows of XN (n,d) and XM (m,d).

    ``theta`` are length-scales that DIVIDE (kernelfunc.py:42,58).
    gaussian  kernelfunc.py:36-51   exp(-1/2 sum ((a-b)/theta)^2)
    kpls      kernelfunc.py:43-49   exp(-1/2 sum_k [sum_i w_ik^2 (a_i-b_i)^2]/theta_k^2)
    exponential :53-60, matern32 :62-71, matern52 :73-82
    """
    XN = np.asarray(XN, dtype=np.float64)
    XM = np.asarray(XM, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64).ravel()
    diff = np.abs(XN[:, None, :] - XM[None, :, :])           # (n, m, d)
    k = kernel.lower()
    if k in ("gaussian", "iso_gaussian"):
        if plscoeff is None:
            return np.exp(-0.5 * np.sum((diff ** 2) / theta ** 2, axis=2))
        w2 = np.asarray(plscoeff, dtype=np.float64) ** 2     # (d, h)
        md = np.dot(diff ** 2, w2) / theta ** 2              # (n, m, h)
        return np.exp(-0.5 * np.sum(md, axis=2))
    h = diff / theta
    if k == "exponential":
        return np.exp(-np.sum(h, axis=2))
    if k == "matern32":
        return np.prod(1 + np.sqrt(3) * h, axis=2) * np.exp(-np.sqrt(3) * np.sum(h, axis=2))
    if k == "matern52":
        return (np.prod(1 + np.sqrt(5) * h + (5 / 3) * h ** 2, axis=2)
                * np.exp(-np.sqrt(5) * np.sum(h, axis=2)))
    raise ValueError("Kernel option not available!")


def corr_chunked(XN, XM, theta, kernel="gaussian", plscoeff=None, chunk=2048):
    """Same as :func:`corr` without the (n, m, d) temporary (large n*m*d): column chunks sized so that the
    temporary stays below about 128 MB."""
    XM = np.asarray(XM, dtype=np.float64)
    n, d = np.shape(XN)
    chunk = int(max(1, min(chunk, (1 << 24) // max(1, n * d))))
    out = np.empty((np.shape(XN)[0], XM.shape[0]))
    for s in range(0, XM.shape[0], chunk):
        out[:, s:s + chunk] = corr(XN, XM[s:s + chunk], theta, kernel, plscoeff)
    return out


# --------------------------------------------------------------------------
# hyper-parameter vector decoding (likelihood_func.py:71-79, nuggetset :121-165)
# --------------------------------------------------------------------------
def decode_hyper(x, nvar_eff, nkernel, fixed_nugget, trainvar):
    """x (log10) -> dict(theta, sigma2 | None, nugget, eps, wgkf).

    Layout by LENGTH: [theta(nvar_eff) | sigma2? | nugget? | weights(nkernel)?].
    ``nvar_eff`` is n_princomp for KPLS (likelihood_func.py:67-68).
    """
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    off = nvar_eff + (1 if trainvar else 0)
    rest = len(x) - off
    if rest == 0:
        nugget, w = fixed_nugget, np.array([1.0])
    elif rest == 1:
        nugget, w = x[off], np.array([1.0])
    elif rest == nkernel:
        nugget, w = fixed_nugget, x[off:off + nkernel]
    elif rest == nkernel + 1:
        nugget, w = x[off], x[off + 1:off + 1 + nkernel]
    else:
        raise ValueError("Nugget setting is not available")
    return dict(theta=10.0 ** x[:nvar_eff],
                sigma2=(10.0 ** x[nvar_eff]) if trainvar else None,
                nugget=nugget, eps=10.0 ** nugget, wgkf=w / np.sum(w))


def build_psi(X, theta, kernels, wgkf, plscoeff=None):
    """Psi = sum_k wgkf[k] * corr_k(X, X)   (likelihood_func.py:93-102)."""
    Psi = np.zeros((X.shape[0], X.shape[0]))
    for kname, w in zip(kernels, wgkf):
        Psi += w * corr_chunked(X, X, theta, kname, plscoeff)
    return Psi


# --------------------------------------------------------------------------
# likelihood  (likelihood_func.py:6-118 + maincalc :168-213)
# --------------------------------------------------------------------------
def nll(x, X, y, F, kernels=("gaussian",), fixed_nugget=-6.0, trainvar=False,
        plscoeff=None, n_princomp=None, full=False):
    """Negative (concentrated | full) ln-likelihood for one hyper-parameter vector.

    maincalc restated with a Cholesky factor and triangular solves instead of
    the reference's LU solves on triangular matrices (likelihood_func.py:183-193);
    the eigvals print-check (:172-173) is dropped (no effect on results).
    Raises numpy.linalg.LinAlgError when R is not positive definite, like the
    reference (:175 is outside its try block).
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    F = np.asarray(F, dtype=np.float64).reshape(X.shape[0], -1)
    n = X.shape[0]
    nvar_eff = n_princomp if (plscoeff is not None) else X.shape[1]
    hp = decode_hyper(x, nvar_eff, len(kernels), fixed_nugget, trainvar)
    Psi = build_psi(X, hp["theta"], kernels, hp["wgkf"], plscoeff)
    R = Psi + np.eye(n) * hp["eps"]                                   # :171
    L = _chol(R, lower=True)                                          # :175
    lndet = 2.0 * np.sum(np.log(np.abs(np.diag(L))))                  # :180
    Zy = _trsv(L, y, lower=True)
    ZF = _trsv(L, F, lower=True)
    M = ZF.T @ ZF                                                     # F' R^-1 F
    BE = np.linalg.solve(M, ZF.T @ Zy)                                # :187-189
    rt = Zy - ZF @ BE                                                 # L^-1 (y - F BE)
    rRr = float((rt.T @ rt)[0, 0])
    if trainvar:
        s2 = hp["sigma2"]                                             # :197-199
        val = 0.5 * lndet + n / 2 * np.log(2 * np.pi) + n / 2 * np.log(s2) + rRr / (2 * s2)
    else:
        s2 = rRr / n                                                  # :201-203
        val = (n / 2) * np.log(s2) + 0.5 * lndet
    if not full:
        return float(val)
    return dict(NegLnLike=float(val), BE=BE, SigmaSqr=s2, L=L, U=L.T, Psi=Psi,
                Theta=np.atleast_1d(np.asarray(x, dtype=np.float64))[:nvar_eff],
                theta=hp["theta"], nugget=hp["nugget"], wgkf=hp["wgkf"], lndet=lndet,
                alpha=_trsv(L.T, rt, lower=False), M=M, ZF=ZF)


def nll_reference_cost(x, X, y, F, kernels=("gaussian",), fixed_nugget=-6.0, trainvar=False):
    """The same value as `nll`, computed with the reference's OWN operation sequence so that timing it gives the
    cost of the unmodified reference on this host (BASELINE.md section 3 asks for that figure next to the fair one):
    per-dimension (n, n, nvar) distance stack as kernelfunc.py:36-47 builds it, the `eigvals` positive-definiteness
    print-check (likelihood_func.py:172-173), the Cholesky factor (:175) and the six dense `numpy.linalg.solve`
    calls on the triangular factors (:183-193; the reference imports `solve` as `mldivide`).  Ordinary / universal
    Kriging, no KPLS.  Timing aid for bench.py's CPU arms only; `tests/test_oracle_golden.py` pins it to `nll`."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    F = np.asarray(F, dtype=np.float64).reshape(X.shape[0], -1)
    n, nvar = X.shape
    hp = decode_hyper(x, nvar, len(kernels), fixed_nugget, trainvar)
    Psi = np.zeros((n, n))
    for kname, w in zip(kernels, hp["wgkf"]):
        stack = np.empty((n, n, nvar))
        for j in range(nvar):
            dj = np.abs(X[:, j][:, None] - X[:, j][None, :])
            stack[:, :, j] = (dj ** 2) / hp["theta"][j] ** 2 if kname in ("gaussian", "iso_gaussian") else dj / hp["theta"][j]
        if kname in ("gaussian", "iso_gaussian"):
            Psi += w * np.exp(-0.5 * np.sum(stack, 2))
        elif kname == "exponential":
            Psi += w * np.exp(-np.sum(stack, 2))
        elif kname == "matern32":
            Psi += w * np.prod(1 + np.sqrt(3) * stack, 2) * np.exp(-np.sqrt(3) * np.sum(stack, 2))
        else:
            Psi += w * np.prod(1 + np.sqrt(5) * stack + (5 / 3) * stack ** 2, 2) * np.exp(-np.sqrt(5) * np.sum(stack, 2))
    R = Psi + np.eye(n) * hp["eps"]
    np.any(np.linalg.eigvals(R) < 0)                                  # :172 (general, non-symmetric eigen-solver)
    U = np.linalg.cholesky(R).T                                       # :175-176
    lndet = 2.0 * np.sum(np.log(np.abs(np.diag(U))))
    a = np.linalg.solve(U, np.linalg.solve(U.T, y))                   # :183-184  R^-1 y  (two LU solves)
    Bm = np.linalg.solve(U, np.linalg.solve(U.T, F))                  # :185-186  R^-1 F
    BE = np.linalg.solve(F.T @ Bm, F.T @ a)                           # :187-189
    r = y - F @ BE
    c = np.linalg.solve(U, np.linalg.solve(U.T, r))                   # :192-193
    if trainvar:
        s2 = hp["sigma2"]
        return float(0.5 * lndet + n / 2 * np.log(2 * np.pi) + n / 2 * np.log(s2) + (r.T @ c)[0, 0] / (2 * s2))
    s2 = (r.T @ c)[0, 0] / n
    return float((n / 2) * np.log(s2) + 0.5 * lndet)


def nll_batch(xpop, X, y, F, **kw):
    """Population form: one NLL per row of xpop; +inf where R is not PD."""
    out = np.empty(len(xpop))
    info = np.zeros(len(xpop), dtype=np.int32)
    for b, x in enumerate(np.asarray(xpop, dtype=np.float64)):
        try:
            out[b] = nll(x, X, y, F, **kw)
        except np.linalg.LinAlgError:
            out[b], info[b] = np.inf, 1
        if not np.isfinite(out[b]):
            out[b], info[b] = np.inf, max(info[b], 1)
    return out, info


# --------------------------------------------------------------------------
# trend basis (trendfunction.py:7-107; Legendre only, as used by the path)
# --------------------------------------------------------------------------
def total_order_index(order, nvar):
    """Total-order multi-indices in the reference's order (polytruncation :7-41).

    Degree-ascending; within one degree the order MonCof produces (:43-68),
    flipped (:30): reverse-lexicographic generation then flip of the whole
    stack.  Verified against the reference in tests/test_oracle_golden.py.
    """
    from itertools import combinations
    blocks = []
    for dgr in range(order, -1, -1):
        rows = []
        for c in combinations(range(1, dgr + nvar), nvar - 1):
            cuts = (0,) + c + (dgr + nvar,)
            rows.append([cuts[i + 1] - cuts[i] - 1 for i in range(nvar)])
        blocks.append(np.array(rows, dtype=np.int64).reshape(-1, nvar))
    return np.flip(np.vstack(blocks), 0)


def legendre_basis(idx, Xs):
    """Orthonormal Legendre regression matrix on [-1,1] (trendfunction.py:70-107).

    F[:, a] = prod_j P_{idx[a,j]}(Xs[:, j]) * sqrt(2 idx[a,j] + 1).
    """
    Xs = np.asarray(Xs, dtype=np.float64)
    idx = np.asarray(idx).astype(int)
    maxo = int(idx.max()) if idx.size else 0
    P = np.empty((maxo + 1,) + Xs.shape)
    P[0] = 1.0
    if maxo >= 1:
        P[1] = Xs
    for k in range(1, maxo):
        P[k + 1] = ((2 * k + 1) * Xs * P[k] - k * P[k - 1]) / (k + 1)
    for k in range(maxo + 1):
        P[k] *= np.sqrt(2 * k + 1)
    F = np.ones((Xs.shape[0], idx.shape[0]))
    for a in range(idx.shape[0]):
        for j in range(idx.shape[1]):
            if idx[a, j]:
                F[:, a] *= P[idx[a, j], :, j]
    return F


# --------------------------------------------------------------------------
# prediction / acquisition (prediction.py:67-298)
# --------------------------------------------------------------------------
def standardize_x(x, lb=None, ub=None, mean=None, std=None):
    """samplingplan.py:84-88 ('default') or prediction.py:160 ('std')."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    if mean is not None:
        return (x - mean) / std
    return ((x - lb) / (np.asarray(ub, dtype=np.float64) - lb) - 0.5) * 2


def predict(xs, X, y, F, idx, fit, kernels=("gaussian",), plscoeff=None,
            ybest=None, limit=None, limittype=">=", sigmalcb=None):
    """All predictor outputs at STANDARDISED sites xs (m, d) as 1-D arrays.

    prediction.py:183-228 with R^-1 applied through the Cholesky factor.
    Variance for nind > 1 uses the correct quadratic form
    u' (F' R^-1 F)^-1 u  -- the reference's expression (:226) is only valid for
    nind == 1 and crashes otherwise: UNPINNED for nind > 1.
    EI / PoI / PoF / LCB / EBE: prediction.py:244-285 per point (the reference's
    whole-batch special cases at :250-251 and :264-266 are not reproduced).
    U = |pred| / s  (reliability_analysis/akmcs.py:223).
    """
    xs = np.atleast_2d(np.asarray(xs, dtype=np.float64))
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    L, BE, s2 = fit["L"], fit["BE"], fit["SigmaSqr"]
    theta, wgkf = fit["theta"], fit["wgkf"]
    PC = legendre_basis(idx, xs)                                       # :183-184
    fpc = PC @ BE                                                      # :185
    psi = np.zeros((X.shape[0], xs.shape[0]))
    for kname, w in zip(kernels, wgkf):                                # :187-201
        psi += w * corr_chunked(X, xs, theta, kname, plscoeff)
    f = (fpc + psi.T @ fit["alpha"]).ravel()                           # :209-212
    V = _trsv(L, psi, lower=True)                                      # L^-1 psi
    term1 = 1.0 - np.sum(V * V, axis=0)                                # :224
    ux = fit["ZF"].T @ V - PC.T                                        # :225
    term2 = np.sum(ux * np.linalg.solve(fit["M"], ux), axis=0)         # :226 (fixed)
    ssqr = float(s2) * (term1 + term2)                                 # :227
    s = np.abs(ssqr) ** 0.5                                            # :228
    out = dict(pred=f, ssqr=ssqr, s=s, fpc=fpc.ravel(), ebe=-ssqr)
    yb = float(np.min(y)) if ybest is None else float(ybest)           # :249
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (yb - f) / s
        t1 = (yb - f) * (0.5 + 0.5 * _erf(z / np.sqrt(2)))             # :253-256
        t2 = s / np.sqrt(2 * np.pi) * np.exp(-0.5 * (yb - f) ** 2 / ssqr)  # :257-259
        out["ei"] = -(t1 + t2 + np.finfo(float).tiny)                  # :268-269
        out["poi"] = -(0.5 + 0.5 * _erf(z / np.sqrt(2)))               # :271-273
        if limit is not None:
            sign = 1.0 if limittype in (">", ">=") else -1.0           # :275-285
            out["pof"] = 0.5 + 0.5 * _erf(sign * (f - limit) / s / np.sqrt(2))
        out["U"] = np.abs(f) / s
    if sigmalcb is not None:
        out["lcb"] = f - sigmalcb * ssqr                               # :245
    return out


def akmcs_stats(G, s):
    """AK-MCS reductions (akmcs.py:208-238): counts, min U and its first index."""
    G = np.asarray(G).ravel()
    s = np.asarray(s).ravel()
    U = np.abs(G) / s
    i = int(np.argmin(U))
    return dict(n_fail=int(np.sum(G <= 0)),
                n_fail_minus=int(np.sum(G - 1.96 * s <= 0)),
                n_fail_plus=int(np.sum(G + 1.96 * s <= 0)),
                minU=float(U[i]), argminU=i)


# --------------------------------------------------------------------------
# KPLS rotations (kpls_model.py:102-108 -> sklearn PLSRegression, NIPALS PLS1)
# --------------------------------------------------------------------------
def pls1_rotations(X, y, h):
    """x_rotations_ of sklearn.cross_decomposition.PLSRegression(h, scale=True).

    Dependency not vendored by the reference (README.md:23-28, unpinned;
    sklearn 1.9.0 here).  For a single response NIPALS reduces to: centre and
    scale X, y by their (ddof=1) std; w = X'y/|X'y|; t = Xw; p = X't/t't;
    q = y't/t't; deflate X -= t p', y -= t q;  W* = W (P'W)^-1.
    Only w^2 is used downstream (kernelfunc.py:49) so the sign is irrelevant.
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    Xc = X - X.mean(0)
    sd = Xc.std(0, ddof=1)
    sd[sd == 0] = 1.0
    Xc = Xc / sd
    yc = y - y.mean()
    ysd = yc.std(ddof=1)
    yc = yc / (ysd if ysd != 0 else 1.0)
    W = np.zeros((X.shape[1], h))
    P = np.zeros((X.shape[1], h))
    for k in range(h):
        w = Xc.T @ yc
        w /= np.sqrt(w @ w)
        t = Xc @ w
        tt = t @ t
        p = Xc.T @ t / tt
        q = yc @ t / tt
        Xc = Xc - np.outer(t, p)
        yc = yc - t * q
        W[:, k], P[:, k] = w, p
    return W @ np.linalg.pinv(P.T @ W)
