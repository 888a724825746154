"""Host wall time of kb_nll_batch (host buffers) at BASELINE size with and without graph replay."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
n, d, B = int(sys.argv[1]), 10, int(sys.argv[2])
rng = np.random.default_rng(1)
X = rng.uniform(-1, 1, (n, d)); y = np.sum((5 * X) ** 2, axis=1)
mdl = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"])
xpop = rng.uniform(-0.3, 0.7, (B, d))
for _ in range(5):
    mdl.nll_batch(xpop, False, -6.0)
t0 = time.perf_counter(); reps = 50
for _ in range(reps):
    mdl.nll_batch(xpop, False, -6.0)
print("n=%d B=%d: %.1f us per host call" % (n, B, (time.perf_counter() - t0) / reps * 1e6))
