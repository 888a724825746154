"""Host restatement (numpy) of the device CMA-ES generation loop (csrc/kb_cmaes.cu).  TEST INFRASTRUCTURE ONLY.

The reference delegates this loop to pycma (`cma.fmin2`, surrogate_models/kriging_model.py:420-424), a
third-party dependency that is neither vendored nor pinned nor installed here, so the trajectory itself has
no reference oracle: PARITY UNPINNED at the trajectory level (SURVEY section 8c says the same).  What this
file pins is that the device kernel computes the algorithm it claims to: same Philox draws
(oracle/philox.py), symmetric square-root sampling, the update equations of optim_tools/cmaes.py, LAPACK
`eigh` in place of the device's parallel Jacobi."""
import numpy as np

from oracle import philox


def run(fun, x0, sigma0, lb=None, ub=None, scaling=None, popsize=None, maxiter=None, seed=0, tolfun=1e-11,
        tolx=1e-11, record=None):
    """Minimise the batched objective `fun(X) -> f`.  Returns dict(best_x, best_f, gen, stop, hist);
    `record` (a list) receives the evaluated population of every generation."""
    x0 = np.asarray(x0, dtype=float)
    N = len(x0)
    lam = int(popsize) if popsize else 4 + int(3 * np.log(N))
    mu = lam // 2
    w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
    w = w / w.sum()
    mueff = 1.0 / np.sum(w ** 2)
    cc = (4 + mueff / N) / (N + 4 + 2 * mueff / N)
    cs = (mueff + 2) / (N + mueff + 5)
    c1 = 2 / ((N + 1.3) ** 2 + mueff)
    cmu = min(1 - c1, 2 * (mueff - 2 + 1 / mueff) / ((N + 2) ** 2 + mueff))
    damps = 1 + 2 * max(0.0, np.sqrt((mueff - 1) / (N + 1)) - 1) + cs
    chiN = np.sqrt(N) * (1 - 1 / (4 * N) + 1 / (21 * N * N))
    maxiter = int(maxiter) if maxiter else 100 + 150 * (N + 3) ** 2 // int(np.sqrt(lam))
    window = 10 + int(30 * N / lam)
    scale = np.ones(N) if scaling is None else np.asarray(scaling, dtype=float)
    has_b = lb is not None
    if has_b:
        lb, ub = np.asarray(lb, dtype=float), np.asarray(ub, dtype=float)
    clip = (lambda x: np.clip(x, lb, ub)) if has_b else (lambda x: x)
    mean, sigma = x0.copy(), float(sigma0)
    pc, ps = np.zeros(N), np.zeros(N)
    C, A, D = np.eye(N), np.eye(N), np.ones(N)
    best_x, best_f, gen, stop, hist = x0.copy(), np.inf, 0, 0, []
    while True:
        # ask: generation `gen`, individual i = site gen * lam + i of the standard normal population
        z = philox.population(gen * lam + np.arange(lam), N, dist="normal", p0=0.0, p1=1.0, seed=seed)
        y = z @ A.T
        x = mean + sigma * y * scale
        xf = clip(x)
        if record is not None:
            record.append(xf.copy())
        f = np.asarray(fun(xf), dtype=float).copy()
        # tell
        finite = np.isfinite(f)
        worst = np.max(f[finite]) if finite.any() else 0.0
        if finite.any():
            ib = int(np.argmin(np.where(finite, f, np.inf)))
            if f[ib] < best_f:
                best_f, best_x = float(f[ib]), xf[ib].copy()
        f[~finite] = worst + 1e3 * (1.0 + abs(worst))
        if has_b:
            f = f + (1.0 + abs(worst)) * np.sum(((x - xf) / np.maximum(ub - lb, 1e-300)) ** 2, axis=1)
        order = np.argsort(f, kind="stable")[:mu]
        y_w, z_w = w @ y[order], w @ z[order]
        mean = mean + sigma * y_w * scale
        ps = (1 - cs) * ps + np.sqrt(cs * (2 - cs) * mueff) * z_w
        gen += 1
        nps = np.linalg.norm(ps)
        hsig = float((nps / np.sqrt(1 - (1 - cs) ** (2 * gen)) / chiN) < (1.4 + 2 / (N + 1)))
        pc = (1 - cc) * pc + hsig * np.sqrt(cc * (2 - cc) * mueff) * y_w
        ys = y[order]
        C = ((1 - c1 - cmu) * C + c1 * (np.outer(pc, pc) + (1 - hsig) * cc * (2 - cc) * C)
             + cmu * (ys.T * w) @ ys)
        C = np.triu(C) + np.triu(C, 1).T
        sigma *= np.exp((cs / damps) * (nps / chiN - 1))
        d2, B = np.linalg.eigh(C)
        D = np.sqrt(np.maximum(d2, 1e-300))
        A = (B * D) @ B.T
        hist.append(float(np.min(f)))
        if gen >= maxiter:
            stop = 1
        elif sigma * np.max(D) * np.max(scale) < tolx:
            stop = 2
        elif np.max(D) > 1e7 * np.min(D):
            stop = 4
        elif gen > window and (max(hist[-window:]) - min(hist[-window:])) < tolfun:
            stop = 3
        if stop:
            break
    return dict(best_x=best_x, best_f=best_f, gen=gen, stop=stop, hist=np.array(hist), mean=mean, sigma=sigma)
