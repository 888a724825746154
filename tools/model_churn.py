"""Cost of re-creating the device model every update (what SOBO / AK-MCS / MOBO do when the design grows):
create, a few hundred likelihood calls of two shapes, a fit, a few predictions, destroy."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
rng = np.random.default_rng(0)
def cycle(n, d=2, calls=200):
    X = rng.uniform(-1, 1, (n, d)); y = np.sum(X ** 2, axis=1)
    t0 = time.perf_counter()
    mdl = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"])
    t1 = time.perf_counter()
    for i in range(calls):
        mdl.nll_batch(rng.uniform(-0.5, 0.5, (3 if i % 2 else 1, d)), False, -6.0)
    t2 = time.perf_counter()
    mdl.fit_state(np.zeros(d), False, -6.0)
    for _ in range(5):
        mdl.predict(rng.uniform(-1, 1, (1, d)), ("ei",), acq="ei")
    t3 = time.perf_counter()
    mdl.close()
    t4 = time.perf_counter()
    return np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3]) * 1e3
cycle(20)
for n in (20, 30, 100):
    acc = np.median([cycle(n) for _ in range(15)], axis=0)
    print("n=%3d: create %.2f ms | 200 likelihood calls %.2f ms (%.1f us each) | fit + 5 predicts %.2f ms | destroy %.2f ms"
          % (n, acc[0], acc[1], acc[1] * 5, acc[2], acc[3]), flush=True)
