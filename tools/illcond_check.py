"""Predictor accuracy on ill-conditioned fits against an 80-bit evaluation (development aid):
   python tools/illcond_check.py            (run it with KB_PRED_REFINE=0 / 1 to compare the plain and the corrected paths)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
from oracle import kriging_oracle as ko
from oracle import longdouble_truth as lt

CASES = [(45, 1, "gaussian", 5000), (67, 1, "matern52", 3000), (60, 2, "matern52", 3000), (110, 1, "matern32", 4000),
         (126, 2, "gaussian", 2500), (378, 1, "gaussian", 3000), (617, 1, "matern52", 300)]
for n, d, kernel, m in CASES:
    rng = np.random.default_rng(1000 + n)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sum(np.sin(2 * X) + 0.3 * X ** 2, axis=1) + 0.05 * rng.standard_normal(n)
    idx = ko.total_order_index(0, d)
    F = ko.legendre_basis(idx, X)
    x = rng.uniform(-0.2, 0.6, d)
    kw = dict(kernels=(kernel,))
    mdl = _lib.DeviceModel(X, y, F, [kernel])
    mdl.fit_state(x, False, -6.0, want_U=False, want_Psi=False)
    fit = ko.nll(x, X, y, F, full=True, **kw)
    mdl.set_norm(_lib.NORM_NONE)
    xq = rng.uniform(-1, 1, (m, d))
    outs, _ = mdl.predict(xq, ("pred", "ssqr"), acq="pred")
    rp = ko.predict(xq, X, y, F, idx, fit, **kw)
    tf = lt.nll(x, X, y, F, full=True, **kw)
    tp = lt.predict(xq, X, y, F, idx, tf, **kw)
    msg = []
    for k in ("pred", "ssqr"):
        sc = np.max(np.abs(tp[k]))
        msg.append(f"{k}: gpu {np.max(np.abs(outs[k] - tp[k])) / sc:.1e} lapack {np.max(np.abs(rp[k] - tp[k])) / sc:.1e}")
    print(f"n={n} d={d} {kernel} m={m} KB_PRED_REFINE={os.environ.get('KB_PRED_REFINE', 'auto')}: " + " | ".join(msg), flush=True)
    mdl.close()
