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

from bayesianopttools_b200 import parallel
from bayesianopttools_b200.misc.sampling.samplingplan import realval
from .cmaes import fmin_batched
from .ehvi.EHVIcomputation import ehvi_best, ehvicalc
from .ga.uncGA import uncGA


_sweep_calls = 0


def sweep_candidates(krigobj, ncand, rng=None, sampler="sobol", seed=None, span=None):
    """Candidate designs of the `sweep` optimisers.  Default: a Sobol design generated ON the device
    (no host->device copy of the candidates; a different stretch of the sequence on every call so that
    successive updates do not re-score the same sites); `sampler='random'` draws device Philox uniforms;
    a caller-supplied numpy `rng` keeps everything on the host (reproducible tests).  `span=(lo, hi)`:
    only candidates lo .. hi-1 of that same set (one rank's shard)."""
    global _sweep_calls
    K = krigobj.KrigInfo
    lo, hi = (0, int(ncand)) if span is None else span
    if rng is not None:
        return realval(K["lb"], K["ub"], rng.random((int(ncand), K["nvar"])))[lo:hi]
    from .. import _lib
    call = _sweep_calls
    _sweep_calls += 1
    if sampler.lower() == "random":
        return _lib.DevicePoints.generate(hi - lo, K["nvar"], dist="random", lb=K["lb"], ub=K["ub"],
                                          seed=call if seed is None else seed, first_index=lo)
    first = 1 + (call * int(ncand)) % max(1, (1 << 32) - 2 * int(ncand))
    return _lib.DevicePoints.generate(hi - lo, K["nvar"], dist="sobol", lb=K["lb"], ub=K["ub"],
                                      first_index=first + lo)


def run_single_opt(krigobj, soboInfo, krigconstlist=None, cheapconstlist=None):
    if krigconstlist is not None or cheapconstlist is not None:
        raise NotImplementedError("constrained single-objective acquisition is outside the Kriging hot path")
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
        # soboInfo["shard_candidates"] under torch.distributed: every rank scores a contiguous
        # share of the SAME candidate set, one 64-byte record per rank is all-gathered (lowest candidate index on
        # ties) and the owner of the winner broadcasts its coordinates; all ranks must make this call.
        ncand = int(soboInfo.get("ncandidates", 1_000_000))
        rank, size = parallel.world() if soboInfo.get("shard_candidates", False) else (0, 1)
        lo, hi = parallel.point_shard(ncand, rank, size)
        cand = sweep_candidates(krigobj, ncand, soboInfo.get("rng"), soboInfo.get("sweep_sampler", "sobol"),
                                soboInfo.get("seed"), span=(lo, hi) if size > 1 else None)
        res, _ = krigobj.predict_sweep(cand, acq=acquifunc)
        if size > 1:
            res = dict(res)
            res["best_idx"] += lo
            res["min_u_idx"] += lo
            res = parallel.gather_sweep_result(res)
            counts = [b - a for a, b in (parallel.point_shard(ncand, r, size) for r in range(size))]
            owner, local = parallel.owner_of(res["best_idx"], counts)
            xnext = parallel.broadcast_row(cand[local] if owner == rank else None, owner, nvar)
        else:
            xnext = np.array(cand[res["best_idx"]], dtype=float)
        fnext = res["best_val"]
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


def _as_list(v):
    return v if isinstance(v, list) else [v]


def multiconstfun(x, ypar, kriglist, moboInfo, krigconstlist=None, cheapconstlist=None):
    """acquifunc_opt.py:283-334: EHVI scaled by the product of the constraint models' probabilities
    of feasibility and of the cheap constraints' 0/1 verdicts.  x may be one design or a batch."""
    x = np.asarray(x, dtype=float)
    xs = x.reshape(1, -1) if x.ndim == 1 else x
    fx = np.atleast_1d(ehvicalc(xs, ypar, moboInfo, kriglist)).astype(float)
    if krigconstlist is not None:
        for kc in _as_list(krigconstlist):
            fx = fx * np.asarray(kc.predict(xs, ["PoF"]), dtype=float).ravel()
    if cheapconstlist is not None:
        for fn in _as_list(cheapconstlist):
            fx = fx * np.array([float(fn(row)) for row in xs])
    return float(fx[0]) if x.ndim == 1 else fx


def run_multi_opt(kriglist, moboInfo, ypar, krigconstlist=None, cheapconstlist=None):
    """Next design by minimising the (negated) expected hypervolume improvement
    (reference: acquifunc_opt.py:105-211).  `lbfgsb` and `cobyla` are the reference's multi-start local
    searches on the one-design call; `cmaes` and `ga` evaluate each generation as one batch; `sweep`
    scores `ncandidates` random designs in one pass (both objective models + the cell sum on the device,
    only the argmax comes back) and polishes the winner with L-BFGS-B."""
    opt = moboInfo["acquifuncopt"].lower()
    if moboInfo["acquifunc"].lower() != "ehvi":
        raise ValueError("Acquisition function handle is not available")
    K = kriglist[0].KrigInfo
    nres, nvar = moboInfo["nrestart"], K["nvar"]
    lbv, ubv = np.asarray(K["lb"], dtype=float), np.asarray(K["ub"], dtype=float)
    constrained = krigconstlist is not None or cheapconstlist is not None
    if constrained:
        fun, args = multiconstfun, (ypar, kriglist, moboInfo, krigconstlist, cheapconstlist)
    else:
        fun, args = ehvicalc, (ypar, moboInfo, kriglist)
    batch = lambda P: np.asarray(fun(np.atleast_2d(P), *args), dtype=float).ravel()

    if opt == "sweep":
        if constrained:          # constraint callbacks need the candidates on the host
            cand = sweep_candidates(kriglist[0], moboInfo.get("ncandidates", 100_000),
                                    moboInfo.get("rng") or np.random.default_rng(moboInfo.get("seed")))
            f = batch(cand)
            I = int(np.argmin(f))
            xnext, fnext = cand[I].copy(), float(f[I])
        else:
            # moboInfo["shard_candidates"]: as in run_single_opt, every rank scores its share of the same set
            ncand = int(moboInfo.get("ncandidates", 1_000_000))
            rank, size = parallel.world() if moboInfo.get("shard_candidates", False) else (0, 1)
            lo, hi = parallel.point_shard(ncand, rank, size)
            cand = sweep_candidates(kriglist[0], ncand, moboInfo.get("rng"), moboInfo.get("sweep_sampler", "sobol"),
                                    moboInfo.get("seed"), span=(lo, hi) if size > 1 else None)
            I, val = ehvi_best(cand, ypar, moboInfo, kriglist)
            if size > 1:
                rec = dict(best_val=-float(val), best_idx=int(I) + lo, min_u=np.inf, min_u_idx=0, n_fail=0,
                           n_fail_minus=0, n_fail_plus=0, n_points=hi - lo)
                rec = parallel.gather_sweep_result(rec)
                counts = [b - a for a, b in (parallel.point_shard(ncand, r, size) for r in range(size))]
                owner, local = parallel.owner_of(rec["best_idx"], counts)
                xnext = parallel.broadcast_row(cand[local] if owner == rank else None, owner, nvar)
                fnext = rec["best_val"]
            else:
                xnext, fnext = np.array(cand[I], dtype=float), -val
        if moboInfo.get("sweep_polish", True):
            r = minimize(fun, xnext, method="L-BFGS-B", bounds=np.column_stack((lbv, ubv)), args=args)
            if r.fun < fnext:
                xnext, fnext = r.x, float(r.fun)
        return xnext, fnext
    if opt == "ga":
        xnext, fnext, _ = uncGA(batch, lb=lbv, ub=ubv, seed=moboInfo.get("seed"))
        return xnext, fnext

    Xrand = realval(lbv, ubv, np.random.rand(nres, nvar))
    xnextcand, fnextcand = np.zeros((nres, nvar)), np.zeros(nres)
    for im in range(nres):
        if opt == "cmaes":
            xnextcand[im], fnextcand[im], _ = fmin_batched(batch, Xrand[im], 1.0)
        elif opt == "lbfgsb":
            r = minimize(fun, Xrand[im], method="L-BFGS-B", bounds=np.column_stack((lbv, ubv)), args=args)
            xnextcand[im], fnextcand[im] = r.x, r.fun
        elif opt == "cobyla":
            cons = []
            for i in range(nvar):
                cons.append(lambda x, *a, i=i: x[i] - lbv[i])
                cons.append(lambda x, *a, i=i: ubv[i] - x[i])
            r = fmin_cobyla(fun, Xrand[im], cons, rhobeg=0.5, rhoend=1e-4, args=args)
            xnextcand[im], fnextcand[im] = r, fun(r, *args)
        else:
            raise ValueError(moboInfo["acquifuncopt"], " is not a valid acquisition function optimizer.")
    I = int(np.argmin(fnextcand))
    return xnextcand[I, :], fnextcand[I]
