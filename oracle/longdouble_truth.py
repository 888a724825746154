"""80-bit (numpy.longdouble) evaluation of the same likelihood / predictor formulas as kriging_oracle.py.

TEST INFRASTRUCTURE ONLY (see kriging_oracle.py).  On ill-conditioned correlation matrices (cond R ~ 1e10: smooth
kernels, nugget tuned down to 1e-8) two correct FP64 implementations differ by more than the 1e-9 parity bar --
LAPACK itself is 1e-9 ... 5e-9 away from this evaluation.  This module is the arbiter for those cases: a GPU
result counts as correct when it is no further from the 80-bit value than the LAPACK oracle is
(tools/fuzz_parity.py, tools/accuracy_vs_longdouble.py).  Formulas: likelihood_func.py:168-213,
kernelfunc.py:36-82, prediction.py:183-228; pure-numpy loops, so only for n up to about a thousand.
"""
import numpy as np

from . import kriging_oracle as ko

LD = np.longdouble


def corr_ld(XN, XM, theta, kernel):
    XN, XM, theta = np.asarray(XN, dtype=LD), np.asarray(XM, dtype=LD), np.asarray(theta, dtype=LD).ravel()
    diff = np.abs(XN[:, None, :] - XM[None, :, :])
    k = kernel.lower()
    if k in ("gaussian", "iso_gaussian"):
        return np.exp(-LD(0.5) * np.sum((diff ** 2) / theta ** 2, axis=2))
    h = diff / theta
    if k == "exponential":
        return np.exp(-np.sum(h, axis=2))
    if k == "matern32":
        return np.prod(1 + np.sqrt(LD(3)) * h, axis=2) * np.exp(-np.sqrt(LD(3)) * np.sum(h, axis=2))
    if k == "matern52":
        return np.prod(1 + np.sqrt(LD(5)) * h + (LD(5) / LD(3)) * h ** 2, axis=2) * np.exp(-np.sqrt(LD(5)) * np.sum(h, axis=2))
    raise ValueError(kernel)


def chol_ld(A):
    A = np.array(A, dtype=LD)
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        if not A[j, j] > 0:
            raise np.linalg.LinAlgError("not positive definite")
        L[j, j] = np.sqrt(A[j, j])
        L[j + 1:, j] = A[j + 1:, j] / L[j, j]
        A[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
    return L


def fsolve_ld(L, Bm):
    """L^-1 B, column by column (forward substitution)"""
    Bm = np.array(Bm, dtype=LD).reshape(L.shape[0], -1)
    Z = np.zeros_like(Bm)
    for i in range(L.shape[0]):
        Z[i] = (Bm[i] - L[i, :i] @ Z[:i]) / L[i, i]
    return Z


def bsolve_ld(L, Bm):
    """L^-T B"""
    Bm = np.array(Bm, dtype=LD).reshape(L.shape[0], -1)
    Z = np.zeros_like(Bm)
    for i in range(L.shape[0] - 1, -1, -1):
        Z[i] = (Bm[i] - L[i + 1:, i] @ Z[i + 1:]) / L[i, i]
    return Z


def nll(x, X, y, F, kernels=("gaussian",), fixed_nugget=-6.0, trainvar=False, full=False):
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    hp = ko.decode_hyper(x, X.shape[1], len(kernels), fixed_nugget, trainvar)
    # 10**x in long double from the double log10 value (the FP64 paths round 10**x to double first: use that
    # same double so that the comparison is about the linear algebra, not about pow)
    theta = hp["theta"].astype(LD)
    w = np.asarray(hp["wgkf"], dtype=LD)
    Psi = sum(wk * corr_ld(X, X, theta, kn) for kn, wk in zip(kernels, w))
    R = Psi + np.eye(n, dtype=LD) * LD(hp["eps"])
    L = chol_ld(R)
    Zy = fsolve_ld(L, np.asarray(y, dtype=LD).reshape(-1, 1))
    ZF = fsolve_ld(L, np.asarray(F, dtype=LD).reshape(n, -1))
    M = ZF.T @ ZF
    LM = chol_ld(M)
    BE = bsolve_ld(LM, fsolve_ld(LM, ZF.T @ Zy))
    rt = Zy - ZF @ BE
    rr = (rt.T @ rt)[0, 0]
    lndet = 2 * np.sum(np.log(np.diag(L)))
    if trainvar:
        s2 = LD(hp["sigma2"])
        val = LD(0.5) * lndet + LD(n) / 2 * np.log(2 * LD(np.pi)) + LD(n) / 2 * np.log(s2) + rr / (2 * s2)
    else:
        s2 = rr / n
        val = LD(n) / 2 * np.log(s2) + LD(0.5) * lndet
    if not full:
        return float(val)
    return dict(NegLnLike=float(val), L=L, BE=BE, SigmaSqr=s2, theta=theta, wgkf=w, alpha=bsolve_ld(L, rt), ZF=ZF, LM=LM)


def predict(xs, X, y, F, idx, fit, kernels=("gaussian",)):
    """mean and variance at standardised sites (prediction.py:183-228), long double"""
    xs = np.atleast_2d(np.asarray(xs, dtype=np.float64))
    PC = ko.legendre_basis(idx, xs).astype(LD)
    psi = sum(wk * corr_ld(X, xs, fit["theta"], kn) for kn, wk in zip(kernels, fit["wgkf"]))
    f = (PC @ fit["BE"] + psi.T @ fit["alpha"]).ravel()
    V = fsolve_ld(fit["L"], psi)
    term1 = 1 - np.sum(V * V, axis=0)
    ux = fit["ZF"].T @ V - PC.T
    W = fsolve_ld(fit["LM"], ux)
    ssqr = fit["SigmaSqr"] * (term1 + np.sum(W * W, axis=0))
    return dict(pred=f.astype(np.float64), ssqr=ssqr.astype(np.float64))
