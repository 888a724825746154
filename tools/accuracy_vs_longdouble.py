"""Which FP64 path is closer to the truth on ill-conditioned correlation matrices?  GPU NLL and the numpy/LAPACK
oracle NLL against an 80-bit long-double evaluation of the same formula (development aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
from oracle import kriging_oracle as ko


def nll_longdouble(x, X, y, F, nug10):
    ld = np.longdouble
    theta = ld(10.0) ** np.asarray(x[:X.shape[1]], dtype=ld)
    Xl = X.astype(ld)
    D = (Xl[:, None, :] - Xl[None, :, :]) / theta
    R = np.exp(-ld(0.5) * np.sum(D * D, axis=2)) + np.eye(len(X), dtype=ld) * ld(10.0) ** ld(nug10)
    n = len(X)
    L = np.zeros_like(R)
    A = R.copy()
    for j in range(n):
        L[j, j] = np.sqrt(A[j, j])
        L[j + 1:, j] = A[j + 1:, j] / L[j, j]
        A[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
    def fsolve(b):
        z = np.zeros_like(b)
        for i in range(n):
            z[i] = (b[i] - L[i, :i] @ z[:i]) / L[i, i]
        return z
    zy, zf = fsolve(y.astype(ld)), fsolve(F[:, 0].astype(ld))
    be = (zf @ zy) / (zf @ zf)
    r = zy - zf * be
    s2 = (r @ r) / n
    return float(ld(n) / 2 * np.log(s2) + np.sum(np.log(np.diag(L))))


rng = np.random.default_rng(3)
for n, d, nug in ((150, 2, -8.0), (150, 2, -6.0), (300, 4, -7.5), (300, 4, -6.0)):
    X = rng.uniform(-1, 1, (n, d))
    y = np.sum(np.sin(2 * X) + 0.3 * X ** 2, axis=1)
    F = np.ones((n, 1))
    xpop = np.hstack([rng.uniform(-0.3, 0.8, (4, d)), np.full((4, 1), nug)])
    mdl = _lib.DeviceModel(X, y, F, ["gaussian"])
    g = mdl.nll_batch(xpop, False, -6.0)
    for b in range(4):
        o = ko.nll(xpop[b], X, y, F)
        t = nll_longdouble(xpop[b], X, y, F, nug)
        print(f"n={n} d={d} nugget=1e{nug}: truth {t:.10f}  gpu rel err {abs(g[b]-t)/abs(t):.2e}  oracle(LAPACK) rel err {abs(o-t)/abs(t):.2e}  gpu-vs-oracle {abs(g[b]-o)/abs(o):.2e}", flush=True)
    mdl.close()
