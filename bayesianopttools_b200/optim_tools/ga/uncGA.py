"""Real-coded GA with batched fitness evaluation (reference: optim_tools/ga/uncGA.py:8-158).

Same operators as the reference (tournament selection, SBX crossover, Gaussian mutation,
elitism); its three per-individual evaluation sites (:46-50, :108-113, :132-136) become three
batched calls `batchfcn(pop) -> fitness`.
"""
import numpy as np


def uncGA(batchfcn, lb, ub, npop=100, maxg=100, pc=0.9, pm=0.1, eta_c=15.0, sigma_m=0.1, seed=None,
          tol=1e-12, stall=40):
    rng = np.random.default_rng(seed)
    lb = np.atleast_1d(np.asarray(lb, dtype=float))
    ub = np.atleast_1d(np.asarray(ub, dtype=float))
    nvar = len(lb)
    pop = lb + (ub - lb) * rng.random((npop, nvar))
    fit = np.asarray(batchfcn(pop), dtype=float)
    fit[~np.isfinite(fit)] = np.inf
    hist = []
    for g in range(maxg):
        # tournament selection
        a, b = rng.integers(0, npop, npop), rng.integers(0, npop, npop)
        parents = pop[np.where(fit[a] <= fit[b], a, b)]
        # SBX crossover
        kids = parents.copy()
        for i in range(0, npop - 1, 2):
            if rng.random() < pc:
                u = rng.random(nvar)
                beta = np.where(u <= 0.5, (2 * u) ** (1 / (eta_c + 1)), (1 / (2 * (1 - u))) ** (1 / (eta_c + 1)))
                p1, p2 = parents[i], parents[i + 1]
                kids[i] = 0.5 * ((1 + beta) * p1 + (1 - beta) * p2)
                kids[i + 1] = 0.5 * ((1 - beta) * p1 + (1 + beta) * p2)
        kids = np.clip(kids, lb, ub)
        fk = np.asarray(batchfcn(kids), dtype=float)
        # Gaussian mutation
        mask = rng.random(npop) < pm
        if mask.any():
            sg = sigma_m * (1.0 - g / maxg) ** 2
            mut = np.clip(kids[mask] + sg * (ub - lb) * rng.standard_normal((int(mask.sum()), nvar)), lb, ub)
            fm = np.asarray(batchfcn(mut), dtype=float)
            kids[mask], fk[mask] = mut, fm
        fk[~np.isfinite(fk)] = np.inf
        # elitist survivor selection
        allp, allf = np.vstack((pop, kids)), np.concatenate((fit, fk))
        keep = np.argsort(allf, kind="stable")[:npop]
        pop, fit = allp[keep], allf[keep]
        hist.append(fit[0])
        if len(hist) > stall and abs(hist[-stall] - hist[-1]) < tol:
            break
    return pop[0], float(fit[0]), np.array(hist)
