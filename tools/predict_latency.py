"""Host wall time of one kb_predict call on a handful of sites (what a local acquisition optimiser does
one design at a time, acquifunc_opt.py:45-76) and on a GA / CMA-ES generation.  Development aid."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
for n, d, m in ((30, 2, 1), (50, 2, 1), (200, 10, 1), (200, 10, 11), (1000, 10, 1), (1000, 10, 11), (1000, 10, 100),
                (1000, 10, 1000), (1000, 10, 1100), (4000, 20, 1), (4000, 20, 100)):
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (n, d)); y = np.sum(X ** 2, axis=1)
    mdl = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"])
    mdl.fit_state(np.zeros(d), False, -6.0, want_U=False, want_Psi=False)
    mdl.set_norm(_lib.NORM_NONE)
    xq = rng.uniform(-1, 1, (m, d))
    for _ in range(10):
        mdl.predict(xq, ("ei",), acq="ei", ybest=float(y.min()))
    t0 = time.perf_counter(); reps = 100
    for _ in range(reps):
        mdl.predict(xq, ("ei",), acq="ei", ybest=float(y.min()))
    print("n=%4d d=%2d sites=%4d: %.1f us per predict call (host wall)" % (n, d, m, (time.perf_counter() - t0) / reps * 1e6), flush=True)
    mdl.close()
