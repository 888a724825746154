"""CPU emulation of the tile Cholesky's panel-solve variants against an 80-bit evaluation (development aid, no GPU).

The device factorisation forms L[i,k] from the updated tile C and the diagonal tile L[k,k].  Round 1 multiplied by
the explicitly inverted 64x64 block; this script measures, in numpy, how far each candidate replacement leaves
the concentrated likelihood from a long-double evaluation on ill-conditioned Gaussian designs (nugget 1e-8):
   inv64   C @ inv(Lkk)^T                                  (round 1)
   sub16   block forward substitution, 16-column panels, explicit 16x16 inverses
   sub8    block forward substitution,  8-column panels, explicit  8x8  inverses
   sub8s   as sub8 with the off-diagonal blocks pre-multiplied by the 8x8 inverses (one accumulator per panel)
   trsm    column-by-column triangular solve              (LAPACK's dtrsm)
   python tools/emulate_blocking.py [exp_ulp_noise]
"""
import sys

import numpy as np
from scipy.linalg import solve_triangular

T = 64
STATS = {}


def panel_solve(C, Lkk, how):
    if how == "inv64":
        W = solve_triangular(Lkk, np.eye(Lkk.shape[0]), lower=True)
        return C @ np.tril(W).T
    if how == "trsm":
        return solve_triangular(Lkk, C.T, lower=True).T
    if how.startswith("adapt:"):   # residual correction only where the diagonal of Lkk spans more than tau
        d = np.diag(Lkk)
        how = "inv64r" if d.max() / d.min() > float(how[6:]) else "inv64"
        STATS[how] = STATS.get(how, 0) + 1
        return panel_solve(C, Lkk, how)
    if how == "inv64r":     # explicit inverse + one residual correction
        W = np.tril(solve_triangular(Lkk, np.eye(Lkk.shape[0]), lower=True))
        X = C @ W.T
        return X + (C - X @ Lkk.T) @ W.T
    w = {"sub16": 16, "sub8": 8, "sub8s": 8, "sub4": 4}[how]
    m = Lkk.shape[0]
    X = np.zeros_like(C)
    for j0 in range(0, m, w):
        j1 = min(j0 + w, m)
        Wjj = np.tril(solve_triangular(Lkk[j0:j1, j0:j1], np.eye(j1 - j0), lower=True))
        if how == "sub8s":
            acc = C[:, j0:j1] @ Wjj.T
            for p0 in range(0, j0, w):
                M = Wjj @ Lkk[j0:j1, p0:p0 + w]
                acc = acc - X[:, p0:p0 + w] @ M.T
            X[:, j0:j1] = acc
        else:
            Tj = C[:, j0:j1] - X[:, :j0] @ Lkk[j0:j1, :j0].T
            X[:, j0:j1] = Tj @ Wjj.T
    return X


def tile_nll(R, y, F, how):
    n = R.shape[0]
    q = 1 + F.shape[1]
    A = np.zeros((n + q, n))
    A[:n] = R
    A[n] = y
    A[n + 1:] = F.T
    nb = (n + T - 1) // T
    L = np.zeros_like(A)
    for k in range(nb):
        c0, c1 = k * T, min((k + 1) * T, n)
        Ckk = A[c0:c1, c0:c1] - L[c0:c1, :c0] @ L[c0:c1, :c0].T
        Lkk = np.linalg.cholesky(Ckk)
        L[c0:c1, c0:c1] = Lkk
        if c1 < n + q:
            C = A[c1:, c0:c1] - L[c1:, :c0] @ L[c0:c1, :c0].T
            L[c1:, c0:c1] = panel_solve(C, Lkk, how)
    zy, zf = L[n], L[n + 1]
    be = (zf @ zy) / (zf @ zf)
    r = zy - zf * be
    s2 = (r @ r) / n
    return n / 2 * np.log(s2) + np.sum(np.log(np.diag(L[:n])))


def lapack_nll(R, y, F):
    n = R.shape[0]
    L = np.linalg.cholesky(R)
    zy = solve_triangular(L, y, lower=True)
    zf = solve_triangular(L, F[:, 0], lower=True)
    be = (zf @ zy) / (zf @ zf)
    r = zy - zf * be
    return n / 2 * np.log((r @ r) / n) + np.sum(np.log(np.diag(L)))


def nll_longdouble(theta10, X, y, F, nug10):
    ld = np.longdouble
    theta = ld(10.0) ** np.asarray(theta10, dtype=ld)
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


def sweep():
    """error vs the refinement threshold on ill- and well-conditioned designs"""
    rng = np.random.default_rng(11)
    taus = ["adapt:1", "adapt:10", "adapt:30", "adapt:100", "adapt:300", "adapt:1000", "inv64", "trsm"]
    for n, d, nug, lo, hi in ((234, 1, -8.0, -0.3, 0.8), (260, 2, -8.0, -0.3, 0.8), (300, 2, -6.0, -0.3, 0.8),
                              (400, 6, -6.0, 0.3, 1.0), (500, 10, -6.0, -1.0, 1.0), (500, 3, -7.0, -0.3, 0.8)):
        X = rng.uniform(-1, 1, (n, d))
        y = np.sum(np.sin(2 * X) + 0.3 * X ** 2, axis=1)
        F = np.ones((n, 1))
        for b in range(3):
            th10 = rng.uniform(lo, hi, d)
            D = (X[:, None, :] - X[None, :, :]) / 10.0 ** th10
            R = np.exp(-0.5 * np.sum(D * D, axis=2)) + np.eye(n) * 10.0 ** nug
            t = nll_longdouble(th10, X, y, F, nug)
            out = []
            for h in taus:
                STATS.clear()
                try:
                    e = abs(tile_nll(R, y, F, h) - t) / abs(t)
                except np.linalg.LinAlgError:
                    e = float("nan")
                out.append(f"{h.replace('adapt:', 't')} {e:.1e}" + (f"[{STATS.get('inv64r', 0)}/{STATS.get('inv64r', 0) + STATS.get('inv64', 0)}]" if h.startswith("adapt") else ""))
            el = abs(lapack_nll(R, y, F) - t) / abs(t)
            print(f"n={n} d={d} nug={nug}: lapack {el:.1e} | " + "  ".join(out), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        return sweep()
    noise = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
    rng = np.random.default_rng(3)
    hows = ["inv64", "inv64r", "sub16", "sub8", "sub8s", "sub4", "trsm"]
    worst = {h: 0.0 for h in hows + ["lapack"]}
    for n, d, nug in ((150, 2, -8.0), (150, 1, -8.0), (300, 4, -7.5), (260, 2, -8.0), (234, 1, -8.0)):
        X = rng.uniform(-1, 1, (n, d))
        y = np.sum(np.sin(2 * X) + 0.3 * X ** 2, axis=1)
        F = np.ones((n, 1))
        for b in range(4):
            th10 = rng.uniform(-0.3, 0.8, d)
            theta = 10.0 ** th10
            D = (X[:, None, :] - X[None, :, :]) / theta
            R = np.exp(-0.5 * np.sum(D * D, axis=2))
            if noise:
                R = R * (1.0 + noise * 1.1e-16 * rng.uniform(-1, 1, R.shape))
                R = np.tril(R) + np.tril(R, -1).T
            R = R + np.eye(n) * 10.0 ** nug
            t = nll_longdouble(th10, X, y, F, nug)
            try:
                errs = {h: abs(tile_nll(R, y, F, h) - t) / abs(t) for h in hows}
                errs["lapack"] = abs(lapack_nll(R, y, F) - t) / abs(t)
            except np.linalg.LinAlgError:
                print(f"n={n} d={d}: not PD")
                continue
            for h, e in errs.items():
                worst[h] = max(worst[h], e)
            print(f"n={n} d={d} nug={nug}: " + "  ".join(f"{h} {e:.1e}" for h, e in errs.items()), flush=True)
    print("worst: " + "  ".join(f"{h} {e:.1e}" for h, e in worst.items()))


if __name__ == "__main__":
    main()
