"""Host-side latency of one likelihood population call at KrigDemo / AK-MCS sizes (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
for n, d, B in ((20, 2, 3), (40, 2, 3), (50, 2, 16), (64, 2, 3), (100, 3, 4), (200, 10, 11)):
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (n, d)); y = np.sum(X ** 2, axis=1)
    mdl = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"])
    xpop = rng.uniform(-0.5, 0.5, (B, d))
    for _ in range(20):
        mdl.nll_batch(xpop, False, -6.0)
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        mdl.nll_batch(xpop, False, -6.0)
    dt = (time.perf_counter() - t0) / reps
    mdl.set_profiling(True)
    mdl.nll_batch(xpop, False, -6.0)
    st = mdl.stage_ms(0)
    print(f"n={n} d={d} B={B}: {dt*1e6:.0f} us per call (host wall); device stages ms decode {st[0]:.3f} build {st[1]:.3f} chol {st[2]:.3f} gls {st[3]:.3f}", flush=True)
    mdl.close()
