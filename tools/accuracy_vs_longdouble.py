"""Which FP64 path is closer to the truth on ill-conditioned correlation matrices?  GPU NLL and the numpy/LAPACK
oracle NLL against an 80-bit long-double evaluation of the same formula (oracle/longdouble_truth.py).
Development aid;  KB_CHOL_REFINE=0 shows the round-1 behaviour (no residual correction), =1 corrects every tile."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
from oracle import kriging_oracle as ko
from oracle import longdouble_truth as lt

rng = np.random.default_rng(3)
worst_g = worst_o = 0.0
closer = total = 0
for n, d, nug, kern in ((17, 1, -6.0, "matern52"), (31, 1, -8.0, "gaussian"), (60, 2, -8.0, "gaussian"), (150, 2, -8.0, "gaussian"),
                        (150, 2, -6.0, "gaussian"), (234, 1, -8.0, "gaussian"), (205, 1, -8.0, "matern52"), (260, 2, -8.0, "gaussian"),
                        (300, 4, -7.5, "gaussian"), (300, 4, -6.0, "gaussian"), (600, 2, -8.0, "gaussian")):
    X = rng.uniform(-1, 1, (n, d))
    y = np.sum(np.sin(2 * X) + 0.3 * X ** 2, axis=1)
    F = np.ones((n, 1))
    xpop = np.hstack([rng.uniform(-0.3, 0.8, (4, d)), np.full((4, 1), nug)])
    mdl = _lib.DeviceModel(X, y, F, [kern])
    g = mdl.nll_batch(xpop, False, -6.0)
    for b in range(4):
        o = ko.nll(xpop[b], X, y, F, kernels=(kern,))
        t = lt.nll(xpop[b], X, y, F, kernels=(kern,))
        eg, eo = abs(g[b] - t) / abs(t), abs(o - t) / abs(t)
        worst_g, worst_o = max(worst_g, eg), max(worst_o, eo)
        closer += eg <= eo
        total += 1
        print(f"n={n} d={d} {kern} nugget=1e{nug}: truth {t:.10f}  gpu rel err {eg:.2e}  oracle(LAPACK) rel err {eo:.2e}  gpu-vs-oracle {abs(g[b]-o)/abs(o):.2e}", flush=True)
    mdl.close()
print(f"worst distance from the 80-bit evaluation: gpu {worst_g:.2e}, LAPACK oracle {worst_o:.2e}; gpu closer in {closer} of {total}")
