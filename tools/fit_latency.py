"""Wall time of kb_fit_state (likelihood mode='all': factor + L^-T + predictor state, optional U / Psi export)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
for n, d in ((50, 2), (200, 10), (1000, 10), (2000, 10), (4000, 20)):
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (n, d)); y = np.sum(X ** 2, axis=1)
    mdl = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"])
    th = np.zeros(d)
    out = []
    for kw in (dict(want_U=False, want_Psi=False), dict(want_U=True, want_Psi=True)):
        mdl.fit_state(th, False, -6.0, **kw)
        t0 = time.perf_counter(); reps = 5
        for _ in range(reps):
            mdl.fit_state(th, False, -6.0, **kw)
        out.append((time.perf_counter() - t0) / reps * 1e3)
    print("n=%4d d=%2d: fit_state %.2f ms without exports, %.2f ms with U and Psi copied to the host" % (n, d, out[0], out[1]), flush=True)
    mdl.close()
