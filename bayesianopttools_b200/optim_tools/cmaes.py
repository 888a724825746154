"""Ask/tell CMA-ES whose whole generation is evaluated by ONE batched objective call.

Replaces the reference's `cma.fmin2(likelihood, x0, sigma0, {'bounds', 'scaling_of_variables'})`
(surrogate_models/kriging_model.py:420-424, optim_tools/acquifunc_opt.py:45-50; pycma is a
third-party dependency, absent here) with the (mu/mu_w, lambda)-CMA-ES of Hansen's tutorial.
The point of owning the loop: `ask()` yields the population that goes to the GPU in one
C-ABI call.  Box constraints: candidates are clipped into the box for evaluation and a
quadratic distance penalty is added (so non-finite objective values never enter the update).
"""
import numpy as np


class CMAES:
    def __init__(self, x0, sigma0, bounds=None, scaling=None, popsize=None, seed=None, maxiter=None,
                 tolfun=1e-11, tolx=1e-11):
        self.N = N = len(x0)
        self.mean = np.asarray(x0, dtype=float).copy()
        self.sigma = float(sigma0)
        self.scale = np.ones(N) if scaling is None else np.asarray(scaling, dtype=float)
        self.lb = None if bounds is None else np.asarray(bounds[0], dtype=float)
        self.ub = None if bounds is None else np.asarray(bounds[1], dtype=float)
        self.lam = int(popsize) if popsize else 4 + int(3 * np.log(N))
        self.mu = self.lam // 2
        w = np.log(self.mu + 0.5) - np.log(np.arange(1, self.mu + 1))
        self.w = w / w.sum()
        self.mueff = 1.0 / np.sum(self.w ** 2)
        self.cc = (4 + self.mueff / N) / (N + 4 + 2 * self.mueff / N)
        self.cs = (self.mueff + 2) / (N + self.mueff + 5)
        self.c1 = 2 / ((N + 1.3) ** 2 + self.mueff)
        self.cmu = min(1 - self.c1, 2 * (self.mueff - 2 + 1 / self.mueff) / ((N + 2) ** 2 + self.mueff))
        self.damps = 1 + 2 * max(0.0, np.sqrt((self.mueff - 1) / (N + 1)) - 1) + self.cs
        self.chiN = np.sqrt(N) * (1 - 1 / (4 * N) + 1 / (21 * N * N))
        self.pc, self.ps = np.zeros(N), np.zeros(N)
        self.C, self.B, self.D = np.eye(N), np.eye(N), np.ones(N)
        self.rng = np.random.default_rng(seed)
        self.maxiter = int(maxiter) if maxiter else 100 + 150 * (N + 3) ** 2 // int(np.sqrt(self.lam))
        self.tolfun, self.tolx = tolfun, tolx
        self.gen = 0
        self.best_x, self.best_f = self.mean.copy(), np.inf
        self.hist = []
        self._z = self._y = self._x = None

    def ask(self):
        self._z = self.rng.standard_normal((self.lam, self.N))
        self._y = (self._z * self.D) @ self.B.T
        self._x = self.mean + self.sigma * self._y * self.scale
        return self.feasible(self._x)

    def feasible(self, x):
        return x if self.lb is None else np.clip(x, self.lb, self.ub)

    def tell(self, fvals):
        f = np.asarray(fvals, dtype=float).copy()
        xf = self.feasible(self._x)
        finite = np.isfinite(f)
        worst = np.max(f[finite]) if finite.any() else 0.0
        f[~finite] = worst + 1e3 * (1.0 + abs(worst))
        i_best = int(np.argmin(np.where(finite, f, np.inf))) if finite.any() else None
        if i_best is not None and f[i_best] < self.best_f:
            self.best_f, self.best_x = float(f[i_best]), xf[i_best].copy()
        if self.lb is not None:
            span = np.maximum(self.ub - self.lb, 1e-300)
            f = f + (1.0 + abs(worst)) * np.sum(((self._x - xf) / span) ** 2, axis=1)
        order = np.argsort(f, kind="stable")[:self.mu]
        y_w = self.w @ self._y[order]
        z_w = self.w @ self._z[order]
        self.mean = self.mean + self.sigma * y_w * self.scale
        N = self.N
        self.ps = (1 - self.cs) * self.ps + np.sqrt(self.cs * (2 - self.cs) * self.mueff) * (self.B @ z_w)
        self.gen += 1
        hsig = (np.linalg.norm(self.ps) / np.sqrt(1 - (1 - self.cs) ** (2 * self.gen)) / self.chiN) < (1.4 + 2 / (N + 1))
        self.pc = (1 - self.cc) * self.pc + hsig * np.sqrt(self.cc * (2 - self.cc) * self.mueff) * y_w
        ys = self._y[order]
        self.C = ((1 - self.c1 - self.cmu) * self.C
                  + self.c1 * (np.outer(self.pc, self.pc) + (1 - hsig) * self.cc * (2 - self.cc) * self.C)
                  + self.cmu * (ys.T * self.w) @ ys)
        self.sigma *= np.exp((self.cs / self.damps) * (np.linalg.norm(self.ps) / self.chiN - 1))
        self.C = np.triu(self.C) + np.triu(self.C, 1).T
        d2, self.B = np.linalg.eigh(self.C)
        self.D = np.sqrt(np.maximum(d2, 1e-300))
        self.hist.append(float(np.min(f)))

    def stop(self):
        if self.gen >= self.maxiter:
            return True
        if self.sigma * np.max(self.D) * np.max(self.scale) < self.tolx:
            return True
        k = 10 + int(30 * self.N / self.lam)
        if len(self.hist) > k and (max(self.hist[-k:]) - min(self.hist[-k:])) < self.tolfun:
            return True
        if np.max(self.D) > 1e7 * np.min(self.D):
            return True
        return False


def fmin_batched(batch_fun, x0, sigma0, bounds=None, scaling=None, popsize=None, seed=None, maxiter=None):
    """Minimise with one `batch_fun(X) -> f` call per generation. Returns (xbest, fbest, es)."""
    es = CMAES(x0, sigma0, bounds=bounds, scaling=scaling, popsize=popsize, seed=seed, maxiter=maxiter)
    while not es.stop():
        X = es.ask()
        es.tell(batch_fun(X))
    return es.best_x, es.best_f, es
