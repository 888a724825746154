"""Wall time of the reference-facing Python calls at single-design granularity: Kriging.predict(x, 'EI')
and likelihood(x, KrigInfo) -- what scipy's local optimisers call in a loop.  Development aid."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
from bayesianopttools_b200.surrogate_models.supports.likelihood_func import likelihood
for n, d in ((40, 2), (1000, 10)):
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (n, d)); y = 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1).reshape(-1, 1)
    K = initkriginfo("single")
    K.update(X=X, y=y, lb=-5 * np.ones(d), ub=5 * np.ones(d), kernel=["gaussian"], nugget=-6, TrendOrder=0,
             optimizer="lbfgsb", nrestart=1)
    o = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    t0 = time.perf_counter(); o.train(disp=False); tt = time.perf_counter() - t0
    x = rng.uniform(-5, 5, d)
    for _ in range(20):
        o.predict(x, "EI")
    t0 = time.perf_counter(); reps = 300
    for _ in range(reps):
        o.predict(x, "EI")
    tp = (time.perf_counter() - t0) / reps
    th = np.asarray(o.KrigInfo["Theta"]).ravel()
    for _ in range(20):
        likelihood(th, o.KrigInfo, mode="default", trainvar=False)
    t0 = time.perf_counter()
    for _ in range(reps):
        likelihood(th, o.KrigInfo, mode="default", trainvar=False)
    tl = (time.perf_counter() - t0) / reps
    print("n=%4d d=%2d: train %.3f s | Kriging.predict(x,'EI') %.1f us | likelihood(x, KrigInfo) %.1f us" % (n, d, tt, tp * 1e6, tl * 1e6), flush=True)
