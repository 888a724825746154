"""Acquisition optimisation (reference: optim_tools/acquifunc_opt.py:10-102).

The reference optimises the acquisition with multi-start local searches calling
`krigobj.predict(x, acq)` ONE point per call.  Those optimisers are kept (`lbfgsb`, `cobyla`,
`cmaes` -- the latter evaluating each generation as one batched predict), and a `sweep`
optimiser is added: a large candidate set is scored in one fused predict+acquisition launch
with the arg-min reduced on the device (the batched equivalent of the reference's
`np.argmin` over restarts, :74-76), optionally polished by L-BFGS-B from the winner.
"""
import numpy as np
from scipy.optimize import fmin_cobyla, minimize

from bayesianopttools_b200.misc.sampling.samplingplan import realval
from .cmaes import fmin_batched


def sweep_candidates(krigobj, ncand, rng=None):
    rng = np.random.default_rng() if rng is None else rng
    K = krigobj.KrigInfo
    return realval(K["lb"], K["ub"], rng.random((int(ncand), K["nvar"])))


def run_single_opt(krigobj, soboInfo, krigconstlist=None, cheapconstlist=None):
    if krigconstlist is not None or cheapconstlist is not None:
        raise NotImplementedError("constrained acquisition is outside the Kriging hot path")
    acquifuncopt = soboInfo["acquifuncopt"].lower()
    acquifunc = soboInfo["acquifunc"]
    if acquifunc.lower() == "parego":
        acquifunc = soboInfo["paregoacquifunc"]
    K = krigobj.KrigInfo
    if acquifunc.lower() == "lcb":
        K["sigmalcb"] = soboInfo.get("sigmalcb", 3)
    nres, nvar = soboInfo["nrestart"], K["nvar"]
    lbv, ubv = np.asarray(K["lb"], dtype=float), np.asarray(K["ub"], dtype=float)

    if acquifuncopt == "sweep":
        cand = sweep_candidates(krigobj, soboInfo.get("ncandidates", 1_000_000), soboInfo.get("rng"))
        res, _ = krigobj.predict_sweep(cand, acq=acquifunc)
        xnext, fnext = cand[res["best_idx"]].copy(), res["best_val"]
        if soboInfo.get("sweep_polish", True):
            r = minimize(krigobj.predict, xnext, method="L-BFGS-B", bounds=np.column_stack((lbv, ubv)),
                         args=(acquifunc,))
            if r.fun < fnext:
                xnext, fnext = r.x, r.fun
        return xnext, fnext

    Xrand = realval(lbv, ubv, np.random.rand(nres, nvar))
    xnextcand, fnextcand = np.zeros((nres, nvar)), np.zeros(nres)
    for im in range(nres):
        if acquifuncopt == "cmaes":
            def batch(P):
                out = krigobj.predict(P, [acquifunc])
                return np.asarray(out, dtype=float).ravel()
            xnextcand[im], fnextcand[im], _ = fmin_batched(batch, Xrand[im], 1.0)
        elif acquifuncopt == "lbfgsb":
            r = minimize(krigobj.predict, Xrand[im], method="L-BFGS-B", bounds=np.column_stack((lbv, ubv)),
                         args=(acquifunc,))
            xnextcand[im], fnextcand[im] = r.x, r.fun
        elif acquifuncopt == "cobyla":
            cons = []
            for i in range(nvar):
                cons.append(lambda x, *a, i=i: x[i] - lbv[i])
                cons.append(lambda x, *a, i=i: ubv[i] - x[i])
            r = fmin_cobyla(krigobj.predict, Xrand[im], cons, rhobeg=0.5, rhoend=1e-4, args=(acquifunc,))
            xnextcand[im], fnextcand[im] = r, krigobj.predict(r, acquifunc)
        else:
            raise ValueError(soboInfo["acquifuncopt"], " is not a valid acquisition function optimizer.")
    I = int(np.argmin(fnextcand))
    return xnextcand[I, :], fnextcand[I]
