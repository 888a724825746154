"""Where the first second goes: import, library load, CUDA context, first calls of each kernel family."""
import os, sys, time
t0 = time.perf_counter()
import numpy as np
t1 = time.perf_counter()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
_lib.load()
t2 = time.perf_counter()
_lib.require_gpu()
t3 = time.perf_counter()
rng = np.random.default_rng(0)
X = rng.uniform(-1, 1, (200, 4)); y = np.sum(X ** 2, axis=1)
mdl = _lib.DeviceModel(X, y, np.ones((200, 1)), ["gaussian"])
t4 = time.perf_counter()
mdl.nll_batch(np.zeros((3, 4)), False, -6.0)
t5 = time.perf_counter()
mdl.nll_batch(np.zeros((3, 4)), False, -6.0)
t6 = time.perf_counter()
mdl.fit_state(np.zeros(4), False, -6.0)
t7 = time.perf_counter()
mdl.set_norm(_lib.NORM_NONE)
mdl.predict(X[:5], ("ei",), acq="ei")
t8 = time.perf_counter()
mdl.predict(rng.uniform(-1, 1, (5000, 4)), ("ei",), acq="ei")
t9 = time.perf_counter()
print("numpy import %.2f s | library load %.3f s | device check %.3f s | first model %.3f s | first likelihood call %.3f s | "
      "second %.4f s | first fit %.3f s | first few-sites predict %.3f s | first strip predict %.3f s"
      % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7, t9 - t8))
