"""Host wall time of one `ehvicalc`-shaped call (kb_ehvi2d on 1 / 16 / 1000 designs): what MOBO's local
acquisition optimisers do in a loop.  Development aid."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
for n, d in ((50, 2), (1000, 10)):
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (n, d))
    mdls = []
    for j in range(2):
        y = np.sum((X - 0.3 * j) ** 2, axis=1) * (1 - 2 * j) + 2 * j
        m = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"])
        m.fit_state(np.full(d, -0.2), False, -6.0, want_U=False, want_Psi=False)
        m.set_norm(_lib.NORM_NONE)
        mdls.append(m)
    f1 = np.sort(rng.uniform(0, 1, 12))
    P = np.column_stack([f1, np.sort(rng.uniform(-1, 1, 12))[::-1]])
    r = np.array([1.5, 1.5])
    for msites in (1, 16, 1000):
        xq = rng.uniform(-1, 1, (msites, d))
        for _ in range(5):
            _lib.ehvi2d_models(mdls[0], mdls[1], xq, P, r)
        t0 = time.perf_counter(); reps = 100
        for _ in range(reps):
            _lib.ehvi2d_models(mdls[0], mdls[1], xq, P, r)
        print("n=%4d d=%2d designs=%4d: %.1f us per ehvi call" % (n, d, msites, (time.perf_counter() - t0) / reps * 1e6), flush=True)
    for m in mdls:
        m.close()
