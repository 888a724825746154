"""Single-objective Bayesian optimisation loop (reference: optim_tools/SOBO.py:22-197).

Same constructor, `run(disp)` and settings dictionary as the reference's `SOBO`; each pass is
acquisition arg-min (run_single_opt -> the device predictor) -> evaluate -> append -> standardize + train.
`acquifuncopt='sweep'` is the device-side optimiser: a candidate design scored in one fused
predict + acquisition launch with the arg-min reduced on the GPU."""
import numpy as np

from bayesianopttools_b200.testcase.analyticalfcn.cases import evaluate
from .acquifunc_opt import run_single_opt

_ACQUISITIONS = ("ei", "pred", "lcb", "poi", "ebe")
_OPTIMISERS = ("lbfgsb", "cobyla", "cmaes", "sweep")


class SOBO:
    def __init__(self, soboInfo, krigobj, autoupdate=True, expconst=None, chpconst=None):
        self.soboInfo = soboInfocheck(soboInfo, autoupdate)
        self.krigobj, self.autoupdate = krigobj, autoupdate
        self.krigconstlist, self.cheapconstlist = expconst, chpconst

    def _observe(self, x, fun):
        """Objective at x; a failed (NaN) evaluation is replaced by prediction + variance (SOBO.py:69-73)."""
        y = np.asarray(fun(x), dtype=float).reshape(1, -1)
        if np.isnan(y).any():
            SSqr, y_hat = self.krigobj.predict(x, ["SSqr", "pred"])
            y = np.asarray(y_hat + SSqr, dtype=float).reshape(1, -1)
        return y

    def run(self, disp=True):
        """SOBO.py:39-112."""
        K, S = self.krigobj.KrigInfo, self.soboInfo
        fun = S.get("evaluate", lambda x: evaluate(x, K["problem"]))
        self.nup, self.istall = 0, 0
        self.Xall, self.yall = K["X"], K["y"]
        best = [float(np.min(self.yall))]
        new_x, new_y = [], []
        print("Begin single-objective Bayesian optimization process.")
        while self.nup < S["nup"]:
            if self.autoupdate and disp:
                print(f"Update no.: {self.nup + 1}, F-count: {np.size(K['X'], 0)}, "
                      f"Best f(x): {best[self.nup]}, Stall counter: {self.istall}")
            self.xnext, self.metricnext = run_single_opt(self.krigobj, S, self.krigconstlist, self.cheapconstlist)
            if not self.autoupdate:
                break
            self.ynext = self._observe(self.xnext, fun)
            K["X"], K["y"] = np.vstack((K["X"], self.xnext)), np.vstack((K["y"], self.ynext))
            K["nsamp"] = K["X"].shape[0]
            self.krigobj.standardize()
            self.krigobj.train(disp=False)
            new_x.append(np.array(self.xnext, dtype=float))
            new_y.append(self.ynext)
            self.nup += 1
            best.append(float(np.min(K["y"])))
            # stall counter: consecutive updates without a new best value (SOBO.py:98-104)
            self.istall = self.istall + 1 if best[-1] == best[-2] else 0
            if self.istall == S["stalliteration"]:
                break
        self.yhist = np.array(best).reshape(-1, 1)
        if new_x:
            self.xupdate, self.yupdate = np.vstack(new_x), np.vstack(new_y)
        print("Optimization finished, now creating the final outputs.")
        return self.xnext, getattr(self, "ynext", None)


def soboInfocheck(soboInfo, autoupdate):
    """Validation and defaults of the settings dictionary (SOBO.py:115-197)."""
    if autoupdate and "nup" not in soboInfo:
        raise ValueError("Number of updates for Bayesian optimization, soboInfo['nup'], is not specified")
    if not autoupdate:
        soboInfo["nup"], soboInfo["stalliteration"] = 1, 1
    soboInfo.setdefault("stalliteration", int(np.floor(soboInfo.get("nsamp", 2) / 2)))
    acq = soboInfo.setdefault("acquifunc", "EI").lower()
    if acq not in _ACQUISITIONS:
        raise ValueError(soboInfo["acquifunc"], " is not a valid acquisition function.")
    if acq == "lcb":
        soboInfo.setdefault("sigmalcb", 3)
    if soboInfo.setdefault("acquifuncopt", "lbfgsb").lower() not in _OPTIMISERS:
        raise ValueError(soboInfo["acquifuncopt"], " is not a valid acquisition function optimizer.")
    if soboInfo.setdefault("nrestart", 1) < 1:
        raise ValueError("soboInfo['nrestart'] should be at least one")
    soboInfo.setdefault("filename", "temporarydata.mat")
    return soboInfo
