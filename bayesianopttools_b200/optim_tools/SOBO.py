"""Single-objective Bayesian optimisation loop (reference: optim_tools/SOBO.py:22-197)."""
from copy import deepcopy

import numpy as np

from bayesianopttools_b200.testcase.analyticalfcn.cases import evaluate
from .acquifunc_opt import run_single_opt


class SOBO:
    def __init__(self, soboInfo, krigobj, autoupdate=True, expconst=None, chpconst=None):
        self.soboInfo = soboInfocheck(soboInfo, autoupdate)
        self.krigobj = krigobj
        self.autoupdate = autoupdate
        self.krigconstlist = expconst
        self.cheapconstlist = chpconst

    def run(self, disp=True):
        """SOBO.py:39-112: acquisition arg-min -> evaluate -> append -> standardize + train."""
        K = self.krigobj.KrigInfo
        self.nup = 0
        self.Xall, self.yall = K["X"], K["y"]
        self.yhist = np.array([np.min(self.yall)])
        self.istall = 0
        evalfun = self.soboInfo.get("evaluate", lambda x: evaluate(x, K["problem"]))
        print("Begin single-objective Bayesian optimization process.")
        while self.nup < self.soboInfo["nup"]:
            if self.autoupdate and disp:
                print(f"Update no.: {self.nup + 1}, F-count: {np.size(K['X'], 0)}, "
                      f"Best f(x): {self.yhist.ravel()[self.nup]}, Stall counter: {self.istall}")
            self.xnext, self.metricnext = run_single_opt(self.krigobj, self.soboInfo, self.krigconstlist,
                                                         self.cheapconstlist)
            if self.autoupdate is False:
                break
            self.ynext = np.asarray(evalfun(self.xnext), dtype=float).reshape(1, -1)
            if np.isnan(self.ynext).any():
                SSqr, y_hat = self.krigobj.predict(self.xnext, ["SSqr", "pred"])
                self.ynext = np.asarray(y_hat + SSqr, dtype=float).reshape(1, -1)
            K["X"] = np.vstack((K["X"], self.xnext))
            K["y"] = np.vstack((K["y"], self.ynext))
            K["nsamp"] = K["X"].shape[0]
            self.krigobj.standardize()
            self.krigobj.train(disp=False)
            if self.nup == 0:
                self.xupdate, self.yupdate = deepcopy(self.xnext), deepcopy(self.ynext)
            else:
                self.xupdate = np.vstack((self.xupdate, self.xnext))
                self.yupdate = np.vstack((self.yupdate, self.ynext))
            self.nup += 1
            self.yhist = np.vstack((self.yhist.reshape(-1, 1), np.min(K["y"])))
            if self.yhist[self.nup, 0] == self.yhist[self.nup - 1, 0]:
                self.istall += 1
                if self.istall == self.soboInfo["stalliteration"]:
                    break
            else:
                self.istall = 0
        print("Optimization finished, now creating the final outputs.")
        return self.xnext, getattr(self, "ynext", None)


def soboInfocheck(soboInfo, autoupdate):
    """SOBO.py:115-197 defaults."""
    if "nup" not in soboInfo:
        if autoupdate is True:
            raise ValueError("Number of updates for Bayesian optimization, soboInfo['nup'], is not specified")
        soboInfo["nup"] = 1
    elif not autoupdate:
        soboInfo["nup"] = 1
    if "stalliteration" not in soboInfo or not autoupdate:
        soboInfo["stalliteration"] = int(np.floor(soboInfo.get("nsamp", 2) / 2)) if autoupdate else 1
    if "acquifunc" not in soboInfo:
        soboInfo["acquifunc"] = "EI"
    elif soboInfo["acquifunc"].lower() not in ["ei", "pred", "lcb", "poi", "ebe"]:
        raise ValueError(soboInfo["acquifunc"], " is not a valid acquisition function.")
    if soboInfo["acquifunc"].lower() == "lcb" and "sigmalcb" not in soboInfo:
        soboInfo["sigmalcb"] = 3
    if "acquifuncopt" not in soboInfo:
        soboInfo["acquifuncopt"] = "lbfgsb"
    elif soboInfo["acquifuncopt"].lower() not in ["lbfgsb", "cobyla", "cmaes", "sweep"]:
        raise ValueError(soboInfo["acquifuncopt"], " is not a valid acquisition function optimizer.")
    if "nrestart" not in soboInfo:
        soboInfo["nrestart"] = 1
    elif soboInfo["nrestart"] < 1:
        raise ValueError("soboInfo['nrestart'] should be at least one")
    if "filename" not in soboInfo:
        soboInfo["filename"] = "temporarydata.mat"
    return soboInfo
