"""Multi-objective Bayesian optimisation loop (reference: optim_tools/MOBO.py:14-467).

Same call surface and bookkeeping as the reference's `MOBO` (run / ehviupdate / paregoupdate /
simultpredehvi / simultpredparego / enrich, `moboinfocheck` defaults).  What changes is where the time
goes: the acquisition is the device EHVI (optim_tools/ehvi/EHVIcomputation.py -> kb_ehvi2d), the
Kriging-believer refits at fixed hyper-parameters (MOBO.py:258-267) are one `likelihood(mode='all')`
= one kb_fit_state each, and `acquifuncopt='sweep'` replaces the multi-start local search by a
one-launch scoring of a large candidate set.  `run` returns the reference's four values
(xupdate, yupdate, supdate, metricall)."""
import time
from copy import deepcopy

import numpy as np
import scipy.io as sio

from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
from bayesianopttools_b200.surrogate_models.supports.likelihood_func import likelihood
from bayesianopttools_b200.surrogate_models.supports.trendfunction import compute_regression_mat
from bayesianopttools_b200.testcase.analyticalfcn.cases import evaluate
from . import searchpareto
from .acquifunc_opt import run_multi_opt, run_single_opt
from .parego import paregopre


class MOBO:
    def __init__(self, moboInfo, kriglist, autoupdate=True, multiupdate=0, savedata=True, expconst=None,
                 chpconst=None):
        self.moboInfo = moboinfocheck(moboInfo, autoupdate)
        self.kriglist = kriglist
        self.krignum = len(kriglist)
        self.autoupdate = autoupdate
        self.multiupdate = multiupdate
        self.savedata = savedata
        self.krigconstlist = expconst
        self.cheapconstlist = chpconst

    def run(self, disp=True, infeasible=None):
        """MOBO.py:56-135."""
        K0 = self.kriglist[0].KrigInfo
        self.nup = 0
        self.Xall = K0["X"]
        self.yall = np.column_stack([np.asarray(k.KrigInfo["y"], dtype=float)[:, 0] for k in self.kriglist])
        if infeasible is not None:
            self.yall = np.delete(self.yall.copy(), infeasible, 0)
            self.Xall = np.delete(self.Xall.copy(), infeasible, 0)
        self.ypar, _ = searchpareto.paretopoint(self.yall)
        print("Begin multi-objective Bayesian optimization process.")
        if self.autoupdate and disp:
            print(f"Update no.: {self.nup + 1}, F-count: {np.size(self.Xall, 0)}, "
                  f"Maximum no. updates: {self.moboInfo['nup'] + 1}")
        acq = self.moboInfo["acquifunc"].lower()
        if acq == "parego":
            self.KrigScalarizedInfo = deepcopy(K0)
            self._train_scalarised(self.Xall, paregopre(self.yall))
            self.paregoupdate(disp)
        elif acq == "ehvi":
            self.ehviupdate(disp)
        else:
            raise ValueError(self.moboInfo["acquifunc"], " is not a valid acquisition function.")
        if disp:
            print("Optimization finished, now creating the final outputs.")
        ntail = self.moboInfo["nup"] * (self.multiupdate if self.multiupdate > 1 else 1)
        xupdate, yupdate = self.Xall[-ntail:, :], self.yall[-ntail:, :]
        supdate = self.spredall[-ntail:, :] if hasattr(self, "spredall") else None
        return xupdate, yupdate, supdate, self.metricall

    def _train_scalarised(self, X, y):
        info = deepcopy(self.KrigScalarizedInfo)
        info["X"], info["y"] = X, y
        self.KrigScalarizedInfo = info
        self.scalkrig = Kriging(info, standardization=True, standtype="default", normy=False, trainvar=False)
        self.scalkrig.train(disp=False)
        return self.scalkrig

    def _record(self, metricnext):
        self.metricall = metricnext if self.nup == 0 else np.vstack((self.metricall, metricnext))

    def ehviupdate(self, disp):
        """MOBO.py:138-190."""
        self.spredall = np.zeros_like(self.yall)
        while self.nup < self.moboInfo["nup"]:
            if self.moboInfo["refpointtype"].lower() == "dynamic":
                hi, lo = np.max(self.yall, 0), np.min(self.yall, 0)
                self.moboInfo["refpoint"] = hi + (hi - lo) * 2
            if self.multiupdate < 0:
                raise ValueError("Number of multiple update must be greater or equal to 0")
            if self.multiupdate in (0, 1):
                xnext, metricnext = run_multi_opt(self.kriglist, self.moboInfo, self.ypar, self.krigconstlist,
                                                  self.cheapconstlist)
                yprednext, sprednext = np.zeros(self.krignum), np.zeros(self.krignum)
                for ii, krigobj in enumerate(self.kriglist):
                    yprednext[ii], sprednext[ii] = _pred_s(krigobj, xnext)
            else:
                xnext, yprednext, sprednext, metricnext = self.simultpredehvi(disp)
            self._record(metricnext)
            if self.autoupdate is False:
                self.Xall = np.vstack((self.Xall, xnext))
                self.yall = np.vstack((self.yall, yprednext))
                self.spredall = np.vstack((self.spredall, sprednext))
                break
            self.spredall = np.vstack((self.spredall, sprednext))
            self.enrich(xnext)
            self.nup += 1
            if disp:
                print(f"Update no.: {self.nup + 1}, F-count: {np.size(self.Xall, 0)}, "
                      f"Maximum no. updates: {self.moboInfo['nup'] + 1}")

    def paregoupdate(self, disp):
        """MOBO.py:193-240."""
        while self.nup < self.moboInfo["nup"]:
            if self.multiupdate < 0:
                raise ValueError("Number of multiple update must be greater or equal to 0")
            if self.multiupdate in (0, 1):
                xnext, metricnext = run_single_opt(self.scalkrig, self.moboInfo, self.krigconstlist,
                                                   self.cheapconstlist)
                yprednext = np.array([float(np.ravel(k.predict(xnext, ["pred"]))[0]) for k in self.kriglist])
            else:
                xnext, yprednext, metricnext = self.simultpredparego()
            self._record(metricnext)
            if self.autoupdate is False:
                self.Xall = np.vstack((self.Xall, xnext))
                self.yall = np.vstack((self.yall, yprednext))
                break
            self.enrich(xnext)
            self.nup += 1
            if disp:
                print(f"Update no.: {self.nup + 1}, F-count: {np.size(self.Xall, 0)}, "
                      f"Maximum no. updates: {self.moboInfo['nup'] + 1}")

    def simultpredehvi(self, disp=False):
        """Kriging believer (MOBO.py:243-303): after each suggested design the objective models take their
        own prediction there as if it were an observation and are re-fitted at FIXED hyper-parameters
        (one device factorisation each), then the next design is sought against the updated front."""
        krigtemp = [_clone(obj) for obj in self.kriglist]
        nvar = krigtemp[0].KrigInfo["nvar"]
        bound = np.vstack((-np.ones((1, nvar)), np.ones((1, nvar))))
        ypartemp, yall = self.ypar, self.yall
        xs, ys, ss, ms = [], [], [], []
        for ii in range(self.multiupdate):
            t1 = time.time()
            if disp:
                print(f"update number {ii + 1}")
            xnext, metrictemp = run_multi_opt(krigtemp, self.moboInfo, ypartemp, self.krigconstlist,
                                              self.cheapconstlist)
            yprednext, sprednext = np.zeros(len(krigtemp)), np.zeros(len(krigtemp))
            for jj, kt in enumerate(krigtemp):
                yprednext[jj], sprednext[jj] = _pred_s(kt, xnext)
                KI = kt.KrigInfo
                KI["X"] = np.vstack((KI["X"], xnext))
                KI["y"] = np.vstack((KI["y"], yprednext[jj]))
                KI["nsamp"] = KI["X"].shape[0]
                kt.standardize()
                KI = kt.KrigInfo
                KI["F"] = compute_regression_mat(KI["idx"], KI["X_norm"], bound, np.ones(nvar))
                kt.KrigInfo = likelihood(KI["Theta"], KI, mode="all", trainvar=kt.trainvar)
            xs.append(np.array(xnext, dtype=float)); ys.append(yprednext.copy()); ss.append(sprednext.copy())
            ms.append(metrictemp)
            yall = np.vstack((yall, yprednext))
            ypartemp, _ = searchpareto.paretopoint(yall)
            if disp:
                print("time: ", time.time() - t1, " s")
        return [np.vstack(xs), np.vstack(ys), np.vstack(ss), np.vstack(ms)]

    def simultpredparego(self):
        """MOBO.py:306-347: several designs per iteration from different scalarisation weights."""
        idxs = np.random.choice(11, self.multiupdate)
        xs, ys, ms = [], [], []
        for ii, idx in enumerate(idxs):
            print(f"update number {ii + 1}")
            info = deepcopy(self.KrigScalarizedInfo)
            info["X"], info["y"] = self.Xall, paregopre(self.yall, idx)
            krigtemp = Kriging(info, standardization=True, standtype="default", normy=False, trainvar=False)
            krigtemp.train(disp=False)
            xnext, metricnext = run_single_opt(krigtemp, self.moboInfo, self.krigconstlist, self.cheapconstlist)
            yprednext = np.array([float(np.ravel(k.predict(xnext, ["pred"]))[0]) for k in self.kriglist])
            xs.append(np.array(xnext, dtype=float)); ys.append(yprednext); ms.append(metricnext)
        return np.vstack(xs), np.vstack(ys), np.vstack(ms)

    def enrich(self, xnext):
        """Evaluate the suggested design(s), append, recompute the front, retrain (MOBO.py:350-402)."""
        K0 = self.kriglist[0].KrigInfo
        evalfun = self.moboInfo.get("evaluate", lambda x: evaluate(x, K0["problem"]))
        xn = np.atleast_2d(np.asarray(xnext, dtype=float))
        ynext = np.vstack([np.asarray(evalfun(row), dtype=float).reshape(1, -1) for row in xn])
        # failed evaluations are replaced by a pessimistic prediction (Forrester et al. 2006)
        for r in np.flatnonzero(np.isnan(ynext).any(axis=1)):
            for jj, k in enumerate(self.kriglist):
                SSqr, y_hat = k.predict(xn[r], ["SSqr", "pred"])
                ynext[r, jj] = float(np.ravel(y_hat)[0] + np.ravel(SSqr)[0])
        self.yall = np.vstack((self.yall, ynext))
        self.Xall = np.vstack((self.Xall, xn))
        self.ypar, I = searchpareto.paretopoint(self.yall)
        acq = self.moboInfo["acquifunc"].lower()
        if acq not in ("ehvi", "parego"):
            raise ValueError(self.moboInfo["acquifunc"], " is not a valid acquisition function.")
        if acq == "parego":
            self._train_scalarised(self.Xall, paregopre(self.yall))
        for index, krigobj in enumerate(self.kriglist):
            krigobj.KrigInfo["X"] = self.Xall
            krigobj.KrigInfo["y"] = self.yall[:, index].reshape(-1, 1)
            krigobj.KrigInfo["nsamp"] = self.Xall.shape[0]
            krigobj.standardize()
            krigobj.train(disp=False)
        if self.savedata:
            I = np.asarray(I).astype(int)
            sio.savemat(self.moboInfo["filename"], {"xbest": self.Xall[I, :], "ybest": self.ypar})


def _pred_s(krigobj, x):
    pred, s = krigobj.predict(x, ["pred", "s"])
    return float(np.ravel(pred)[0]), float(np.ravel(s)[0])


def _clone(krigobj):
    """Independent copy of a Kriging object; the device handle cached in KrigInfo deep-copies to an
    empty holder (supports/device_state.py), so the copy builds its own on first use."""
    return deepcopy(krigobj)


def moboinfocheck(moboInfo, autoupdate):
    """Defaults and validation of the MOBO settings (MOBO.py:405-467)."""
    if "nup" not in moboInfo:
        if autoupdate is True:
            raise ValueError("Number of updates for Bayesian optimization, moboInfo['nup'], is not specified")
        moboInfo["nup"] = 1
        print("Number of updates for Bayesian optimization has been set to 1")
    elif not autoupdate:
        moboInfo["nup"] = 1
        print("Manual mode is active, number of updates for Bayesian optimization is forced to 1")
    if "acquifunc" not in moboInfo:
        moboInfo["acquifunc"] = "EHVI"
        print("The acquisition function is not specified, set to EHVI")
    elif moboInfo["acquifunc"].lower() not in ("ehvi", "parego"):
        raise ValueError(moboInfo["acquifunc"], " is not a valid acquisition function.")
    if moboInfo["acquifunc"].lower() == "ehvi":
        moboInfo["refpointtype"] = "static" if "refpoint" in moboInfo else "dynamic"
    else:
        moboInfo["krignum"] = 1
        moboInfo.setdefault("paregoacquifunc", "EI")
    if "acquifuncopt" not in moboInfo:
        moboInfo["acquifuncopt"] = "lbfgsb"
        print("The acquisition function optimizer is not specified, set to L-BFGS-B.")
    elif moboInfo["acquifuncopt"].lower() not in ("lbfgsb", "cobyla", "cmaes", "ga", "sweep"):
        raise ValueError(moboInfo["acquifuncopt"], " is not a valid acquisition function optimizer.")
    if "nrestart" not in moboInfo:
        moboInfo["nrestart"] = 1
        print("The number of restart for acquisition function optimization is not specified, "
              "setting BayesInfo.nrestart to 1.")
    elif moboInfo["nrestart"] < 1:
        raise ValueError("BayesInfo['nrestart'] should be at least one")
    if "filename" not in moboInfo:
        moboInfo["filename"] = "temporarydata.mat"
        print("The file name for saving the results is not specified, set the name to temporarydata.mat")
    return moboInfo
