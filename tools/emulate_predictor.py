"""CPU emulation of predictor-variance variants against an 80-bit evaluation (development aid, no GPU).
   inv      v = W psi with W = L^-1 explicit (round 1 / current strip kernel)
   invr     v = v0 + W (psi - L v0)              (one residual correction per site, 3x the flops)
   blk<T>   blocked forward substitution with TxT explicit diagonal inverses, off-diagonal blocks from L
   blk<T>r  the same with the residual correction applied inside each diagonal block only
   trsv     LAPACK triangular solve
"""
import sys
import numpy as np
from scipy.linalg import solve_triangular, cholesky
sys.path.insert(0, ".")
from oracle import kriging_oracle as ko, longdouble_truth as lt


def vsolve(L, W, psi, how):
    n = L.shape[0]
    if how == "inv":
        return W @ psi
    if how == "invr":
        v = W @ psi
        return v + W @ (psi - L @ v)
    if how == "trsv":
        return solve_triangular(L, psi, lower=True)
    T = int(how[3:].rstrip("r"))
    v = np.zeros_like(psi)
    for i0 in range(0, n, T):
        i1 = min(i0 + T, n)
        Wd = np.tril(solve_triangular(L[i0:i1, i0:i1], np.eye(i1 - i0), lower=True))
        rhs = psi[i0:i1] - L[i0:i1, :i0] @ v[:i0]
        x = Wd @ rhs
        if how.endswith("r"):
            x = x + Wd @ (rhs - L[i0:i1, i0:i1] @ x)
        v[i0:i1] = x
    return v


rng = np.random.default_rng(5)
hows = ["inv", "invr", "blk128", "blk128r", "blk64", "blk64r", "blk16", "trsv"]
worst = {h: 0.0 for h in hows}
for n, d, nug, kern in ((266, 2, -6.0, "gaussian"), (320, 1, -6.0, "gaussian"), (444, 1, -6.0, "matern52"), (300, 3, -6.0, "gaussian"),
                        (588, 2, -8.0, "gaussian"), (679, 1, -7.0, "matern52"), (500, 6, -6.0, "gaussian"), (120, 2, -6.0, "gaussian")):
    X = rng.uniform(-1, 1, (n, d))
    y = np.sum(np.sin(2 * X) + 0.3 * X ** 2, axis=1)
    F = np.ones((n, 1))
    x = np.concatenate([rng.uniform(-0.3, 0.6, d), [nug]])
    fit = ko.nll(x, X, y, F, kernels=(kern,), full=True)
    tf = lt.nll(x, X, y, F, kernels=(kern,), full=True)
    xq = rng.uniform(-1, 1, (200, d))
    tp = lt.predict(xq, X, y, F, np.zeros((1, d), int), tf, kernels=(kern,))
    L = fit["L"]
    W = np.tril(solve_triangular(L, np.eye(n), lower=True))
    psi = ko.corr_chunked(X, xq, fit["theta"], kern)
    zf = fit["ZF"]
    out = []
    for h in hows:
        V = vsolve(L, W, psi, h)
        term1 = 1.0 - np.sum(V * V, axis=0)
        ux = zf.T @ V - 1.0
        ss = float(fit["SigmaSqr"]) * (term1 + (ux ** 2 / fit["M"][0, 0]).ravel())
        e = np.max(np.abs(ss - tp["ssqr"])) / np.max(np.abs(tp["ssqr"]))
        worst[h] = max(worst[h], e)
        out.append(f"{h} {e:.1e}")
    print(f"n={n} d={d} {kern} nug={nug}: " + "  ".join(out), flush=True)
print("worst: " + "  ".join(f"{h} {e:.1e}" for h, e in worst.items()))
