"""End-to-end AK-MCS through the reference-facing Python API at a large Monte-Carlo population
(BASELINE configs[4] shape): four-branch limit state, d = 2, N points ~ N(0,1)^2, initial DoE of 12
Halton points, `maxupdate` Kriging updates.   python tools/akmcs_e2e.py [n_order] [maxupdate]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200.misc.sampling.samplingplan import sampling
from bayesianopttools_b200.reliability_analysis.akmcs import AKMCS, mcpopgen_device
from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
from bayesianopttools_b200.testcase.RA.testcase import evaluate

n_order = int(sys.argv[1]) if len(sys.argv) > 1 else 7
maxupdate = int(sys.argv[2]) if len(sys.argv) > 2 else 20
np.random.seed(0)
lb, ub = -5.0 * np.ones(2), 5.0 * np.ones(2)
_, X = sampling("halton", 2, 12, result="real", upbound=ub, lobound=lb)
y = evaluate(X, type="fourbranches")
K = initkriginfo("single")
K.update(X=X, y=y, nvar=2, problem="fourbranches", nsamp=12, nrestart=5, ub=ub, lb=lb, optimizer="lbfgsb",
         kernel=["gaussian"], nugget=-6)
t0 = time.time()
krig = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
krig.train(disp=False)
t_train = time.time() - t0
t0 = time.time()
pop = mcpopgen_device(type="gaussian", ndim=2, n_order=n_order, n_coeff=1, stddev=1, mean=0, seed=1)
t_pop = time.time() - t0
t0 = time.time()
ak = AKMCS(krig, dict(init_samp=pop, maxupdate=maxupdate, problem="fourbranches", keep_fields=False))
ak.run(disp=False)
t_run = time.time() - t0
print("N = 1e%d, %d updates: population on device %.3f s, initial train %.2f s, AK-MCS loop %.2f s (%.3f s per update incl. retraining), "
      "Pf = %.6g, cov = %.4g, stop criterion = %.4g, final design size %d"
      % (n_order, len(ak.updateX) - 1, t_pop, t_train, t_run, t_run / max(1, len(ak.updateX) - 1), ak.Pf, ak.cov,
         ak.stop_criteria, krig.KrigInfo["nsamp"]))
