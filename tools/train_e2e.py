"""Wall time of `Kriging.train` + `loocvcalc` + a 1e6-site EI sweep through the reference-facing API at the
BASELINE C2 shape and at the KrigDemo shape.  Development aid (numbers for the README)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo

def run(n, d, optimizer, nrestart):
    rng = np.random.default_rng(1)
    X = rng.uniform(-5, 5, (n, d))
    y = 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1).reshape(-1, 1)
    K = initkriginfo("single")
    K.update(X=X, y=y, lb=-5 * np.ones(d), ub=5 * np.ones(d), kernel=["gaussian"], nugget=-6, TrendOrder=0,
             optimizer=optimizer, nrestart=nrestart, seed=3)
    t0 = time.perf_counter()
    o = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    o.train(disp=False)
    t1 = time.perf_counter()
    err, _ = o.loocvcalc()
    t2 = time.perf_counter()
    xq = rng.uniform(-5, 5, (1_000_000, d))
    res, _ = o.predict_sweep(xq, acq="EI")
    t3 = time.perf_counter()
    print("n=%4d d=%2d %-6s nrestart=%d: train %.3f s (NLL %.4f) | LOOCV %.3f s | EI sweep over 1e6 sites + arg-min %.3f s"
          % (n, d, optimizer, nrestart, t1 - t0, float(np.ravel(o.KrigInfo["NegLnLike"])[0]), t2 - t1, t3 - t2), flush=True)

run(40, 2, "lbfgsb", 7)          # warm-up of the CUDA context and the library
for args in ((40, 2, "lbfgsb", 7), (40, 2, "cmaes", 2), (1000, 10, "lbfgsb", 5), (1000, 10, "cmaes", 1)):
    run(*args)
