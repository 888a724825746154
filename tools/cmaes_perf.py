"""Generations per second of the CMA-ES hyper-parameter search: generation loop on the device
(kb_train_cmaes) vs the host ask/tell loop around the batched likelihood (optim_tools/cmaes.py).
Usage: python tools/cmaes_perf.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from bayesianopttools_b200 import _lib                                     # noqa: E402
from bayesianopttools_b200.optim_tools.cmaes import CMAES                   # noqa: E402

for n, d, lam, gens in ((50, 2, 8, 200), (200, 10, 16, 200), (1000, 10, 64, 100), (2000, 10, 64, 40)):
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (n, d))
    y = 0.5 * np.sum((5 * X) ** 4 - 16 * (5 * X) ** 2 + 5 * (5 * X), axis=1)
    mdl = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"])
    lb, ub = -5 * np.ones(d), 5 * np.ones(d)
    kw = dict(lb=lb, ub=ub, popsize=lam, maxiter=gens, seed=1, tolfun=1e-300, tolx=1e-300)
    mdl.train_cmaes(np.zeros(d), 2.0, False, -6.0, **dict(kw, maxiter=5))        # warm-up (workspaces)
    t0 = time.perf_counter()
    xb, fb, g, why, hist = mdl.train_cmaes(np.zeros(d), 2.0, False, -6.0, **kw)
    td = time.perf_counter() - t0
    es = CMAES(np.zeros(d), 2.0, bounds=[lb, ub], popsize=lam, seed=1, maxiter=gens, tolfun=1e-300, tolx=1e-300)
    t0 = time.perf_counter()
    gh = 0
    while not es.stop():
        P = es.ask()
        es.tell(mdl.nll_batch(P, False, -6.0))
        gh += 1
    th = time.perf_counter() - t0
    print("n=%4d d=%2d popsize=%3d: device loop %4d generations in %.3f s (%.0f gen/s, best %.6f, stop %d) | "
          "host loop %4d generations in %.3f s (%.0f gen/s, best %.6f)"
          % (n, d, lam, g, td, g / td, fb, why, gh, th, gh / th, es.best_f), flush=True)
    mdl.close()
