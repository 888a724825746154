"""Throughput of the batched 2-objective EHVI (kb_ehvi2d: two predictor sweeps + the cell-sum kernel)
against the CPU restatement on a bounded sample.  Usage: python tools/ehvi_perf.py [n] [k] [log2 m]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from bayesianopttools_b200 import _lib                                    # noqa: E402
from oracle import ehvi_oracle as eo                                      # noqa: E402  (checker / CPU leg only)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 20)
d = 6
rng = np.random.default_rng(0)
X = rng.uniform(-1, 1, (n, d))
models = []
for j in range(2):
    y = np.sum((X - 0.2 * j) ** 2, axis=1) * (1 - 2 * j) + j
    dm = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"])
    dm.fit_state(np.full(d, -0.3), False, -6.0, want_U=False, want_Psi=False)
    models.append(dm)
f1 = np.sort(rng.uniform(0, 1, k))
P = np.column_stack([f1, np.sort(rng.uniform(0, 1, k))[::-1]])
r = np.array([1.2, 1.2])
xc = rng.uniform(-1, 1, (m, d))
for _ in range(2):
    _lib.ehvi2d_models(models[0], models[1], xc, P, r, want_values=False)
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    out, idx, val = _lib.ehvi2d_models(models[0], models[1], xc, P, r, want_values=False)
dt = (time.perf_counter() - t0) / reps
print("kb_ehvi2d n=%d k=%d m=%d: %.2f ms per call, %.2f M candidates/s (host buffers, argmax only)" % (n, k, m, dt * 1e3, m / dt / 1e6))
# cell-sum kernel alone through kb_exi2d (host arrays in and out)
mu = rng.uniform(-0.3, 1.3, (m, 2))
s = rng.uniform(1e-3, 0.6, (m, 2))
_lib.exi2d_batch(P, r, mu[:1000], s[:1000])
t0 = time.perf_counter()
got = _lib.exi2d_batch(P, r, mu, s)
dt2 = time.perf_counter() - t0
ms = 20000
t0 = time.perf_counter()
want = eo.exi2d(P, r, mu[:ms], s[:ms])
dtc = time.perf_counter() - t0
print("kb_exi2d m=%d: %.2f ms (%.2f M/s incl. copies); numpy restatement %d candidates: %.2f s (%.3f M/s); max rel diff %.2e"
      % (m, dt2 * 1e3, m / dt2 / 1e6, ms, dtc, ms / dtc / 1e6, np.max(np.abs(got[:ms] - want)) / np.max(want)))
