"""Multi-objective Bayesian optimisation loop (reference: optim_tools/MOBO.py:14-467).

Call surface of the reference's `MOBO` -- constructor arguments, `run(disp, infeasible)`, the attributes
it leaves behind (Xall, yall, ypar, spredall, metricall, nup) and the `moboinfocheck` defaults -- with the
loop written once for both acquisition modes.  Where the time goes is different: the EHVI acquisition is
the device kernel behind optim_tools/ehvi/EHVIcomputation.py (kb_ehvi2d), a Kriging-believer refit at
fixed hyper-parameters (MOBO.py:258-267) is one `likelihood(mode='all')` = one kb_fit_state, and
`acquifuncopt='sweep'` scores a device-generated candidate design in one pass instead of running
multi-start local searches.  `run` returns (xupdate, yupdate, supdate, metricall)."""
import time
from copy import deepcopy

import numpy as np
import scipy.io as sio

from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
from bayesianopttools_b200.surrogate_models.supports.likelihood_func import likelihood
from bayesianopttools_b200.surrogate_models.supports.trendfunction import compute_regression_mat
from bayesianopttools_b200.testcase.analyticalfcn.cases import evaluate
from .acquifunc_opt import run_multi_opt, run_single_opt
from .parego import paregopre
from .searchpareto import paretopoint

_ACQUISITIONS = ("ehvi", "parego")
_OPTIMISERS = ("lbfgsb", "cobyla", "cmaes", "ga", "sweep")


def _scalar_predictions(kriglist, x, what):
    """One row per requested output, one column per objective model, at the single design x."""
    out = np.zeros((len(what), len(kriglist)))
    for j, k in enumerate(kriglist):
        vals = k.predict(x, list(what))
        vals = vals if isinstance(vals, (list, tuple)) else [vals]
        out[:, j] = [float(np.ravel(v)[0]) for v in vals]
    return out


class MOBO:
    def __init__(self, moboInfo, kriglist, autoupdate=True, multiupdate=0, savedata=True, expconst=None,
                 chpconst=None):
        self.moboInfo = moboinfocheck(moboInfo, autoupdate)
        self.kriglist, self.krignum = kriglist, len(kriglist)
        self.autoupdate, self.multiupdate, self.savedata = autoupdate, multiupdate, savedata
        self.krigconstlist, self.cheapconstlist = expconst, chpconst

    # -- public entry -----------------------------------------------------------------------------------
    def run(self, disp=True, infeasible=None):
        """MOBO.py:56-135: collect the current designs, find the front, iterate, return the new designs."""
        lead = self.kriglist[0].KrigInfo
        self.nup = 0
        self.Xall = lead["X"]
        self.yall = np.column_stack([np.asarray(k.KrigInfo["y"], dtype=float)[:, 0] for k in self.kriglist])
        if infeasible is not None:
            self.Xall = np.delete(self.Xall.copy(), infeasible, 0)
            self.yall = np.delete(self.yall.copy(), infeasible, 0)
        self.ypar, _ = paretopoint(self.yall)
        print("Begin multi-objective Bayesian optimization process.")
        self._progress(disp and self.autoupdate)
        mode = self.moboInfo["acquifunc"].lower()
        if mode == "parego":
            self.KrigScalarizedInfo = deepcopy(lead)
            self._fit_scalarised(self.Xall, paregopre(self.yall))
            self.paregoupdate(disp)
        else:
            self.ehviupdate(disp)
        if disp:
            print("Optimization finished, now creating the final outputs.")
        tail = self.moboInfo["nup"] * max(1, self.multiupdate)
        supdate = self.spredall[-tail:, :] if mode == "ehvi" else None
        return self.Xall[-tail:, :], self.yall[-tail:, :], supdate, self.metricall

    def ehviupdate(self, disp):
        """MOBO.py:138-190."""
        self.spredall = np.zeros_like(self.yall)
        self._iterate(disp, ehvi=True)

    def paregoupdate(self, disp):
        """MOBO.py:193-240."""
        self._iterate(disp, ehvi=False)

    # -- the loop -----------------------------------------------------------------------------------------
    def _progress(self, show):
        if show:
            print(f"Update no.: {self.nup + 1}, F-count: {np.size(self.Xall, 0)}, "
                  f"Maximum no. updates: {self.moboInfo['nup'] + 1}")

    def _suggest(self, ehvi, disp):
        """Next design(s) with the objective models' predictions there: (x, ypred, spred or None, metric)."""
        single = self.multiupdate in (0, 1)
        if ehvi:
            if not single:
                return tuple(self.simultpredehvi(disp))
            x, metric = run_multi_opt(self.kriglist, self.moboInfo, self.ypar, self.krigconstlist, self.cheapconstlist)
            ps = _scalar_predictions(self.kriglist, x, ("pred", "s"))
            return x, ps[0], ps[1], metric
        if not single:
            x, yp, metric = self.simultpredparego()
            return x, yp, None, metric
        x, metric = run_single_opt(self.scalkrig, self.moboInfo, self.krigconstlist, self.cheapconstlist)
        return x, _scalar_predictions(self.kriglist, x, ("pred",))[0], None, metric

    def _iterate(self, disp, ehvi):
        info = self.moboInfo
        if self.multiupdate < 0:
            raise ValueError("Number of multiple update must be greater or equal to 0")
        while self.nup < info["nup"]:
            if ehvi and info["refpointtype"].lower() == "dynamic":       # MOBO.py:152-153
                top, low = np.max(self.yall, 0), np.min(self.yall, 0)
                info["refpoint"] = top + 2 * (top - low)
            x, ypred, spred, metric = self._suggest(ehvi, disp)
            self.metricall = metric if self.nup == 0 else np.vstack((self.metricall, metric))
            if spred is not None:
                self.spredall = np.vstack((self.spredall, spred))
            if not self.autoupdate:                                      # manual mode: report and stop
                self.Xall = np.vstack((self.Xall, x))
                self.yall = np.vstack((self.yall, ypred))
                return
            self.enrich(x)
            self.nup += 1
            self._progress(disp)

    # -- several designs per iteration ----------------------------------------------------------------------
    def simultpredehvi(self, disp=False):
        """Kriging believer (MOBO.py:243-303): after each suggestion the objective models take their own
        prediction there as an observation and are re-fitted at FIXED hyper-parameters (one device
        factorisation each); the next suggestion is sought against the front updated the same way."""
        believers = [deepcopy(k) for k in self.kriglist]          # the device handle copies to an empty holder
        nvar = believers[0].KrigInfo["nvar"]
        unit_box = np.vstack((-np.ones((1, nvar)), np.ones((1, nvar))))
        front, yall = self.ypar, self.yall
        rows = []
        for it in range(self.multiupdate):
            tic = time.time()
            if disp:
                print(f"update number {it + 1}")
            x, metric = run_multi_opt(believers, self.moboInfo, front, self.krigconstlist, self.cheapconstlist)
            ps = _scalar_predictions(believers, x, ("pred", "s"))
            for j, kb in enumerate(believers):
                K = kb.KrigInfo
                K["X"] = np.vstack((K["X"], x))
                K["y"] = np.vstack((K["y"], ps[0, j]))
                K["nsamp"] = K["X"].shape[0]
                kb.standardize()
                K = kb.KrigInfo
                K["F"] = compute_regression_mat(K["idx"], K["X_norm"], unit_box, np.ones(nvar))
                kb.KrigInfo = likelihood(K["Theta"], K, mode="all", trainvar=kb.trainvar)
            rows.append((np.array(x, dtype=float), ps[0].copy(), ps[1].copy(), metric))
            yall = np.vstack((yall, ps[0]))
            front, _ = paretopoint(yall)
            if disp:
                print("time: ", time.time() - tic, " s")
        return [np.vstack([r[c] for r in rows]) for c in range(4)]

    def simultpredparego(self):
        """MOBO.py:306-347: one design per randomly drawn scalarisation weight."""
        rows = []
        for it, widx in enumerate(np.random.choice(11, self.multiupdate)):
            print(f"update number {it + 1}")
            model = self._fit_scalarised(self.Xall, paregopre(self.yall, widx), keep=False)
            x, metric = run_single_opt(model, self.moboInfo, self.krigconstlist, self.cheapconstlist)
            rows.append((np.array(x, dtype=float), _scalar_predictions(self.kriglist, x, ("pred",))[0], metric))
        return tuple(np.vstack([r[c] for r in rows]) for c in range(3))

    # -- model upkeep -------------------------------------------------------------------------------------
    def _fit_scalarised(self, X, y, keep=True):
        info = deepcopy(self.KrigScalarizedInfo)
        info["X"], info["y"] = X, y
        model = Kriging(info, standardization=True, standtype="default", normy=False, trainvar=False)
        model.train(disp=False)
        if keep:
            self.KrigScalarizedInfo, self.scalkrig = info, model
        return model

    def enrich(self, xnext):
        """Evaluate the suggested design(s), append, recompute the front, retrain (MOBO.py:350-402)."""
        lead = self.kriglist[0].KrigInfo
        fun = self.moboInfo.get("evaluate", lambda x: evaluate(x, lead["problem"]))
        xs = np.atleast_2d(np.asarray(xnext, dtype=float))
        ys = np.vstack([np.asarray(fun(row), dtype=float).reshape(1, -1) for row in xs])
        # a failed evaluation is replaced by a pessimistic prediction (Forrester, Sobester & Keane 2006)
        for r in np.flatnonzero(np.isnan(ys).any(axis=1)):
            ps = _scalar_predictions(self.kriglist, xs[r], ("pred", "SSqr"))
            ys[r] = ps[0] + ps[1]
        self.Xall, self.yall = np.vstack((self.Xall, xs)), np.vstack((self.yall, ys))
        self.ypar, members = paretopoint(self.yall)
        if self.moboInfo["acquifunc"].lower() == "parego":
            self._fit_scalarised(self.Xall, paregopre(self.yall))
        for j, k in enumerate(self.kriglist):
            k.KrigInfo.update(X=self.Xall, y=self.yall[:, j].reshape(-1, 1), nsamp=self.Xall.shape[0])
            k.standardize()
            k.train(disp=False)
        if self.savedata:
            sio.savemat(self.moboInfo["filename"],
                        {"xbest": self.Xall[np.asarray(members).astype(int), :], "ybest": self.ypar})


def moboinfocheck(moboInfo, autoupdate):
    """Validation and defaults of the settings dictionary (MOBO.py:405-467)."""
    if autoupdate and "nup" not in moboInfo:
        raise ValueError("Number of updates for Bayesian optimization, moboInfo['nup'], is not specified")
    if not autoupdate:
        moboInfo["nup"] = 1                     # manual mode always stops after one suggestion
    acq = moboInfo.setdefault("acquifunc", "EHVI").lower()
    if acq not in _ACQUISITIONS:
        raise ValueError(moboInfo["acquifunc"], " is not a valid acquisition function.")
    if acq == "ehvi":
        moboInfo["refpointtype"] = "static" if "refpoint" in moboInfo else "dynamic"
    else:
        moboInfo["krignum"] = 1
        moboInfo.setdefault("paregoacquifunc", "EI")
    if moboInfo.setdefault("acquifuncopt", "lbfgsb").lower() not in _OPTIMISERS:
        raise ValueError(moboInfo["acquifuncopt"], " is not a valid acquisition function optimizer.")
    if moboInfo.setdefault("nrestart", 1) < 1:
        raise ValueError("BayesInfo['nrestart'] should be at least one")
    moboInfo.setdefault("filename", "temporarydata.mat")
    return moboInfo
