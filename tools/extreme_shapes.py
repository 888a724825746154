"""Parity at the edges of the supported shape space (development aid; bounded memory and oracle time):
many sites with few candidates, many dimensions, many regressors, populations far larger than the chain
workers, wide KPLS.  GPU vs the numpy oracle on NLL (all candidates or a subset) and on a few predictions."""
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
from oracle import kriging_oracle as ko

CASES = [
    # name, n, d, order, B, kernel, kpls_h, check (how many candidates go through the oracle)
    ("many sites", 6000, 5, 0, 2, "gaussian", 0, 1),
    ("many sites, trend 1", 4500, 6, 1, 3, "matern52", 0, 1),
    ("many dimensions", 300, 64, 0, 5, "gaussian", 0, 5),
    ("d = 150 (no KPLS)", 260, 150, 0, 3, "exponential", 0, 3),
    ("many regressors", 1500, 13, 2, 4, "gaussian", 0, 2),
    ("nind close to n", 130, 4, 3, 6, "matern32", 0, 6),
    ("population 600", 200, 3, 0, 600, "gaussian", 0, 12),
    ("population 1500, tiny model", 30, 2, 0, 1500, "matern52", 0, 20),
    ("wide KPLS", 900, 170, 0, 6, "gaussian", 8, 2),
]
worst = 0.0
for name, n, d, order, B, kern, h, ncheck in CASES:
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    X = rng.uniform(-1, 1, (n, d))
    y = np.sum(np.sin(2 * X[:, :8]) + 0.3 * X[:, :8] ** 2, axis=1) + 0.05 * rng.standard_normal(n)
    idx = ko.total_order_index(order, d)
    F = ko.legendre_basis(idx, X)
    kw = dict(kernels=(kern,), trainvar=False, fixed_nugget=-6.0)
    pls = None
    nth = d
    if h:
        pls = rng.normal(size=(d, h)) / np.sqrt(d)
        kw.update(plscoeff=pls, n_princomp=h)
        nth = h
    xpop = rng.uniform(-0.4, 0.8, (B, nth))
    t0 = time.perf_counter()
    mdl = _lib.DeviceModel(X, y, F, [kern], model_type="kpls" if h else "kriging", plscoeff=pls)
    nll, info = mdl.nll_batch(xpop, False, -6.0, return_info=True)
    tg = time.perf_counter() - t0
    pick = np.linspace(0, B - 1, ncheck).astype(int)
    e = 0.0
    for b in pick:
        ref = ko.nll(xpop[b], X, y, F, **kw)
        assert info[b] == 0, (name, b, info[b])
        e = max(e, abs(nll[b] - ref) / max(1.0, abs(ref)))
    # predictions at the first candidate
    mdl.fit_state(xpop[0], False, -6.0, want_U=False, want_Psi=False)
    fit = ko.nll(xpop[0], X, y, F, full=True, **kw)
    mdl.set_trend(idx)
    mdl.set_norm(_lib.NORM_NONE)
    xq = rng.uniform(-1, 1, (300, d))
    outs, _ = mdl.predict(xq, ("pred", "ssqr"), acq="pred")
    p = ko.predict(xq, X, y, F, idx, fit, kernels=(kern,), plscoeff=pls)
    ep = float(np.max(np.abs(outs["pred"] - p["pred"])) / np.max(np.abs(p["pred"])))
    es = float(np.max(np.abs(outs["ssqr"] - p["ssqr"])) / np.max(np.abs(p["ssqr"])))
    worst = max(worst, e)
    print(f"{name:28s} n={n:5d} d={d:3d} nind={F.shape[1]:3d} B={B:4d}: nll rel {e:.2e}  pred {ep:.2e}  ssqr {es:.2e}  "
          f"(gpu create+population {tg:.2f} s)", flush=True)
    mdl.close()
print("worst nll rel", worst)
