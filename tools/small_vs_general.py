"""One-tile fused likelihood kernel (kb_small.cu) against the general pipeline on the same inputs
(run twice: default and KB_SMALL_NLL=0, compare the printed digests).  Development aid."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
from oracle import kriging_oracle as ko
worst = 0.0
for seed, (n, d, nind, kern, trainvar) in enumerate([(5, 1, 1, ["gaussian"], False), (40, 2, 1, ["matern52"], True),
                                                     (64, 3, 4, ["exponential"], False), (33, 6, 7, ["gaussian", "matern32"], False),
                                                     (50, 2, 15, ["gaussian"], False), (64, 8, 1, ["matern32"], False)]):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sum(np.sin(2 * X), axis=1) + 0.01 * rng.standard_normal(n)
    F = np.hstack([np.ones((n, 1)), rng.normal(size=(n, nind - 1))]) if nind > 1 else np.ones((n, 1))
    B = 9
    cols = [rng.uniform(-0.5, 0.7, (B, d))]
    if trainvar:
        cols.append(rng.uniform(-1, 1, (B, 1)))
    if len(kern) > 1:
        cols.append(rng.uniform(0.1, 1, (B, len(kern))))
    xpop = np.hstack(cols)
    mdl = _lib.DeviceModel(X, y, F, kern)
    nll, info = mdl.nll_batch(xpop, trainvar, -6.0, return_info=True)
    beta, s2 = mdl.nll_extras(B)
    ref, rinfo = ko.nll_batch(xpop, X, y, F, kernels=tuple(kern), trainvar=trainvar, fixed_nugget=-6.0)
    e = float(np.max(np.abs(nll - ref) / np.maximum(1.0, np.abs(ref))))
    worst = max(worst, e)
    print(f"n={n} d={d} nind={nind} {kern} trainvar={trainvar}: rel err vs oracle {e:.2e}  digest {np.sum(nll):.15e} {np.sum(beta):.15e} {np.sum(s2):.15e} flags {int(np.sum(info != 0))}/{int(np.sum(rinfo != 0))}")
    mdl.close()
print("worst", worst)
