"""Kriging model class over the GPU hot path (reference: surrogate_models/kriging_model.py).

Same constructor, attributes, KrigInfo contract and methods as the reference's ``Kriging``;
the arithmetic (likelihood, Cholesky, prediction) runs in libkrigb200.so.  The hyper-parameter
optimisers stay on the host but evaluate POPULATIONS: the finite-difference stencil of
L-BFGS-B / SLSQP (nbhyp+1 candidates), a CMA-ES generation or a GA population is one batched
call instead of one `likelihood()` per candidate (kriging_model.py:420-446).
"""
import threading

import numpy as np
from scipy.optimize import fmin_cobyla, minimize

from bayesianopttools_b200.misc.sampling.samplingplan import sampling, standardize
from bayesianopttools_b200.optim_tools.cmaes import fmin_batched
from bayesianopttools_b200.optim_tools.ga.uncGA import uncGA
from .supports.errperf import errperf
from .supports.likelihood_func import likelihood, likelihood_batch
from .supports.prediction import prediction
from .supports.trendfunction import compute_regression_mat, polytruncation

PENALTY = 1e4      # the reference's penalty value for failed evaluations (likelihood_func.py:207)
_TLS = threading.local()   # which lock-step slot the current optimiser thread owns (see _Lockstep)


class _Lockstep:
    """Runs the objective calls of several optimiser threads as ONE device call.

    The reference's multi-start search (`parallelopt`, kriging_model.py:326-389) runs `nrestart` independent local
    optimisers; with L-BFGS-B each of them asks for nbhyp + 1 likelihoods at a time, which is far too small a
    population to fill the device.  Here every restart runs in its own thread; a thread that needs likelihood
    values deposits its candidates and waits, and when every still-running restart has deposited, the last one
    evaluates the union in one `likelihood_batch` call and hands each thread its slice.  Each optimiser sees
    exactly the values it would see alone (the kernels are deterministic per candidate, whatever else is in the
    population), so every restart ends where the sequential loop would end -- nrestart times fewer device calls."""

    def __init__(self, batch_fn, nthreads):
        self.fn, self.active = batch_fn, nthreads
        self.cv = threading.Condition()
        self.pending, self.results = {}, {}
        self.gen, self.error, self.calls = 0, None, 0

    def _flush(self):                      # lock held; every active thread is waiting on it
        tids = sorted(self.pending)
        sizes = [len(self.pending[t]) for t in tids]
        try:
            f = np.asarray(self.fn(np.vstack([self.pending[t] for t in tids])))
            off = 0
            for t, n in zip(tids, sizes):
                self.results[t] = f[off:off + n]
                off += n
        except BaseException as e:         # noqa: BLE001 -- handed to every waiting thread
            self.error = e
        self.calls += 1
        self.pending.clear()
        self.gen += 1
        self.cv.notify_all()

    def evaluate(self, tid, pts):
        with self.cv:
            if self.error is not None:
                raise self.error
            self.pending[tid] = np.atleast_2d(np.asarray(pts, dtype=float))
            if len(self.pending) >= self.active:
                self._flush()
            else:
                gen = self.gen
                while self.gen == gen:
                    self.cv.wait()
            if self.error is not None:
                raise self.error
            return self.results.pop(tid)

    def finish(self, tid):
        with self.cv:
            self.active -= 1
            if self.active > 0 and len(self.pending) >= self.active:
                self._flush()


def lbfgsb_lockstep(batch_nll, x0s, lbv, ubv, eps=1e-3, maxcor=10, ftol=2.2204460492503131e-09, gtol=1e-5,
                    maxiter=15000, maxls=20):
    """`scipy.optimize.minimize(method="L-BFGS-B", jac=True)` for several start points at once, on ONE thread.

    scipy's driver (`_minimize_lbfgsb`, scipy/optimize/_lbfgsb_py.py) is a reverse-communication loop around
    `_lbfgsb.setulb`: the routine hands back an x and asks for (f, g).  Here every restart owns its own `setulb`
    state; each round advances every unfinished restart to its next (f, g) request, evaluates the forward-difference
    stencils of ALL of them in one `batch_nll` call (nrestart x (nbhyp + 1) candidates) and feeds the values back.
    Arguments, evaluation points and values are exactly those of the sequential scipy call, so each restart ends at
    the bit-identical point (tests/test_gpu_branches.py::test_lockstep_restarts_equal_the_sequential_search); the
    thread rendezvous of `_Lockstep` (0.3 ms per round against 0.1 ms of device call on small designs) is gone.
    Returns (x*, f*) rows and the number of device calls."""
    from scipy.optimize import _lbfgsb
    x0s = np.atleast_2d(np.asarray(x0s, dtype=float))
    R, n = x0s.shape
    lbv, ubv = np.asarray(lbv, dtype=float), np.asarray(ubv, dtype=float)
    idt = np.int32
    factr = ftol / np.finfo(float).eps

    diag_i = np.arange(n)

    def stencils(xs):
        """(f, g) at every x of xs from one population: scipy's '2-point' rule with bounds (the step flips where
        x + eps leaves the box), PENALTY for failed evaluations -- the arithmetic of Kriging._fun_and_grad.
        Vectorised over the start points (the per-point numpy calls of a Python loop here were most of the host time of
        a round on small designs: 120 us against 55 us of device call); the evaluation points and the differences are
        formed by the same floating-point operations as before (x + 0.0 is x)."""
        X = np.asarray(xs, dtype=np.float64).reshape(len(xs), n)
        H = np.full(X.shape, eps)
        H[(X + H > ubv) & (X - H >= lbv)] = -eps
        pts = np.repeat(X[:, None, :], n + 1, axis=1)            # (P, n + 1, n): row 0 is x itself
        stepped = X + H
        pts[:, diag_i + 1, diag_i] = stepped
        f = np.asarray(batch_nll(pts.reshape(-1, n)), dtype=np.float64)
        f = np.where(np.isfinite(f), f, PENALTY).reshape(len(xs), n + 1)
        G = (f[:, 1:] - f[:, :1]) / (stepped - X)
        return [(float(f[i, 0]), G[i]) for i in range(len(xs))]

    st = []
    for r in range(R):
        x = np.array(np.clip(x0s[r], lbv, ubv), dtype=np.float64)
        st.append(dict(x=x, x_eval=None, f=np.array(0.0, dtype=np.float64), g=np.zeros(n), nit=0, done=False,
                       wa=np.zeros(2 * maxcor * n + 5 * n + 11 * maxcor * maxcor + 8 * maxcor), iwa=np.zeros(3 * n, dtype=idt),
                       task=np.zeros(2, dtype=idt), ln_task=np.zeros(2, dtype=idt), lsave=np.zeros(4, dtype=idt),
                       isave=np.zeros(44, dtype=idt), dsave=np.zeros(29), cache=None))
    nbd = np.full(n, 2, dtype=idt)
    calls = 1
    # scipy's ScalarFunction evaluates (f, g) at the clipped start point when it is built and answers the routine's
    # first request from that cache
    for s_, fg in zip(st, stencils([s_["x"].copy() for s_ in st])):
        s_["cache"] = (s_["x"].copy(), fg)
    while True:
        pending = []
        for s_ in st:
            while not s_["done"]:
                s_["g"] = s_["g"].astype(np.float64)
                _lbfgsb.setulb(maxcor, s_["x"], lbv, ubv, nbd, s_["f"], s_["g"], factr, gtol, s_["wa"], s_["iwa"],
                               s_["task"], s_["lsave"], s_["isave"], s_["dsave"], maxls, s_["ln_task"])
                t0 = s_["task"][0]
                if t0 == 3:
                    cx, cfg = s_["cache"]
                    if np.array_equal(cx, s_["x"]):
                        s_["f"], s_["g"] = cfg[0], cfg[1]
                        continue
                    pending.append(s_)
                    break
                elif t0 == 1:
                    s_["nit"] += 1
                    if s_["nit"] >= maxiter:
                        s_["task"][0], s_["task"][1] = 5, 504
                else:
                    s_["done"] = True
        if not pending:
            break
        calls += 1
        for s_, fg in zip(pending, stencils([s_["x"].copy() for s_ in pending])):
            s_["cache"] = (s_["x"].copy(), fg)
            s_["f"], s_["g"] = fg[0], fg[1]
    return np.array([s_["x"] for s_ in st]), np.array([float(s_["f"]) for s_ in st]), calls


class Kriging:
    """See the reference docstring (kriging_model.py:16-52) for the KrigInfo keys."""

    def __init__(self, KrigInfo, ub=5, lb=-5, standardization=False, standtype="default", normy=True,
                 trainvar=True, inherit=False):
        KrigInfo["nvar"] = np.size(KrigInfo["X"], 1)
        KrigInfo["nsamp"] = np.size(KrigInfo["X"], 0)
        self.nbhyp = KrigInfo["nvar"] + 1 if trainvar is True else KrigInfo["nvar"]
        self.trainvar = trainvar
        self._lockstep = None
        self.normy = normy
        self.standardization = standardization
        self.standtype = standtype
        self.Y = KrigInfo["y"]
        self.X = KrigInfo["X"]
        self.lb = lb
        self.ub = ub
        if not inherit:
            self.type = "kriging"
            KrigInfo["type"] = self.type
            KrigInfo, scaling = kriginfocheck(KrigInfo, lb, ub, self.nbhyp)
            self.KrigInfo = KrigInfo
            self.scaling = scaling
            self.sigmacmaes = (ub - lb) / 5
            if self.standardization is True:
                self.standardize()

    # ------------------------------------------------------------------
    def standardize(self):
        """kriging_model.py:102-167 (host): X_norm, optional y_norm, trend index and F."""
        K = self.KrigInfo
        K["idx"] = polytruncation(K["TrendOrder"], K["nvar"], 1)
        self.Y = K["y"]
        if self.standardization is True:
            if self.standtype.lower() == "default":
                K["normtype"] = "default"
                bound = np.vstack((-np.ones((1, K["nvar"])), np.ones((1, K["nvar"]))))
                if self.normy is True:
                    # intended behaviour of kriging_model.py:121-129 (the reference passes `normy=` to a
                    # callee whose keyword is `norm_y` and crashes)
                    rng = np.vstack((np.hstack((K["lb"], np.min(self.Y, axis=0))),
                                     np.hstack((K["ub"], np.max(self.Y, axis=0)))))
                    K["X_norm"], K["y_norm"] = standardize(K["X"], K["y"], type="default", norm_y=True, range=rng)
                    K["norm_y"] = True
                else:
                    K["X_norm"] = standardize(K["X"], K["y"], type="default",
                                              range=np.vstack((K["lb"], K["ub"])))
                    K.pop("y_norm", None)
                    K["norm_y"] = False
            else:
                K["normtype"] = "std"
                if self.normy is True:
                    (K["X_norm"], K["y_norm"], K["X_mean"], K["y_mean"], K["X_std"], K["y_std"]) = \
                        standardize(K["X"], K["y"], type="std", norm_y=True)
                    K["norm_y"] = True
                else:
                    K["X_norm"], K["X_mean"], K["X_std"] = standardize(K["X"], K["y"], type="std")
                    K.pop("y_norm", None)
                    K["norm_y"] = False
                bound = np.vstack((np.min(K["X_norm"], axis=0), np.max(K["X_norm"], axis=0)))
            K["standardization"] = True
            K["F"] = compute_regression_mat(K["idx"], K["X_norm"], bound, np.ones(K["nvar"]))
        else:
            K["standardization"] = False
            K["norm_y"] = False
            bound = np.vstack((np.min(K["X"], axis=0), np.max(K["X"], axis=0)))
            K["F"] = compute_regression_mat(K["idx"], K["X"], bound, np.ones(K["nvar"]))

    # ------------------------------------------------------------------
    def _batch_nll(self, xpop):
        ls = getattr(self, "_lockstep", None)
        if ls is not None:
            tid = getattr(_TLS, "tid", None)
            if tid is not None:
                return ls.evaluate(tid, xpop)
        return likelihood_batch(xpop, self.KrigInfo, trainvar=self.trainvar)

    def _fun_and_grad(self, x, lbv, ubv, eps=1e-3):
        """Objective and forward-difference gradient from ONE batched evaluation; the stencil
        follows scipy's '2-point' rule with bounds (step flipped where x+eps leaves the box)."""
        n = len(x)
        h = np.full(n, eps)
        flip = (x + h > ubv) & (x - h >= lbv)
        h[flip] = -eps
        pts = np.vstack([x[None, :], x[None, :] + np.diag(h)])
        f = self._batch_nll(pts)
        f = np.where(np.isfinite(f), f, PENALTY)
        dx = (pts[1:, :] - x[None, :])[np.arange(n), np.arange(n)]
        return float(f[0]), (f[1:] - f[0]) / dx

    def train(self, parallel=False, disp=True, pre_theta=None):
        """kriging_model.py:169-260."""
        K = self.KrigInfo
        if disp:
            print("Begin train hyperparam.")
        if K["kernel"] == ["iso_gaussian"]:
            self.nbhyp = 1
        elif len(K["kernel"]) != 1 and "iso_gaussian" in K["kernel"]:
            raise NotImplementedError("Isotropic Gaussian kernel is not available for composite kernel")
        elif len(K["ubhyp"]) != self.nbhyp:
            self.nbhyp = len(K["ubhyp"])
        nhyp = len(K["ubhyp"]) if self.nbhyp != 1 or K["kernel"] != ["iso_gaussian"] else 1

        if K["nrestart"] < 1:
            xhyp = np.zeros((1, nhyp))
        else:
            _, xhyp = sampling("sobol", nhyp, K["nrestart"], result="real",
                               upbound=K["ubhyp"][:nhyp], lobound=K["lbhyp"][:nhyp])
        if pre_theta is not None:
            pre_theta = np.asarray(pre_theta, dtype=float)
            xhyp = np.vstack((pre_theta, np.random.rand(K["nrestart"] - 1, nhyp) + pre_theta))
        if K.get("shard_population", False) or K.get("shard_restarts", False):
            # ranks that share one training must start from the same points (np.random.rand above is per process)
            from bayesianopttools_b200 import parallel as par
            xhyp = par.broadcast_array(xhyp, 0)

        if self.nbhyp == 1:
            best_x, neglnlikecand = self._train_1d()
            if disp:
                print(f"Best hyperparameter is {best_x}")
                print(f"With NegLnLikelihood of {neglnlikecand}")
        else:
            if disp:
                print(f"Training {K['nrestart']} hyperparameter(s)")
            bestxcand, neglnlikecand = self.parallelopt(xhyp, parallel, None, disp)
            I = int(np.argmin(neglnlikecand))
            best_x = bestxcand[I, :]
            if disp:
                print("Single Objective, train hyperparam, end.")
                print(f"Best hyperparameter is {best_x}")
                print(f"With NegLnLikelihood of {neglnlikecand[I]}")
        self.KrigInfo = likelihood(best_x, K, mode="all", trainvar=self.trainvar)

    def _train_1d(self):
        """nbhyp == 1 (kriging_model.py:213-227): batched bracketing + golden section, or batched GA."""
        K = self.KrigInfo
        if K["optimizer"] == "ga":
            bx, bf, _ = uncGA(self._batch_nll, lb=self.lb, ub=self.ub, npop=100, maxg=100)
            return np.atleast_1d(bx), bf
        grid = np.linspace(self.lb, self.ub, 65)
        f = self._batch_nll(grid[:, None])
        f = np.where(np.isfinite(f), f, np.inf)
        i = int(np.argmin(f))
        a, b = grid[max(i - 1, 0)], grid[min(i + 1, len(grid) - 1)]
        gr = (np.sqrt(5) - 1) / 2
        for _ in range(60):
            c, d = b - gr * (b - a), a + gr * (b - a)
            fc, fd = self._batch_nll(np.array([[c], [d]]))
            if not np.isfinite(fc):
                fc = np.inf
            if not np.isfinite(fd):
                fd = np.inf
            if fc < fd:
                b = d
            else:
                a = c
            if abs(b - a) < 1e-9:
                break
        x = np.array([(a + b) / 2])
        return x, float(self._batch_nll(x[None, :])[0])

    def parallelopt(self, xhyp, parallel=False, optimbound=None, disp=True):
        """kriging_model.py:326-389: one local search per start point, arg-min taken by the caller.

        The reference offers a process pool over the restarts (`parallel=True`).  Here the restarts always run
        concurrently in two nested ways, both transparent to the result:
          * on one GPU they run in LOCKSTEP (`_Lockstep`): their objective calls are merged into one population
            per device call (nrestart x (nbhyp + 1) candidates for L-BFGS-B instead of nbhyp + 1);
          * with KrigInfo["shard_restarts"] under torch.distributed, restart i runs on rank i mod G and the
            (x*, NLL*) rows are all-gathered -- what AK-MCS / SOBO use to stop replicating the retraining.
        KrigInfo["lockstep_restarts"] = False restores the sequential loop (debug / A-B aid)."""
        K = self.KrigInfo
        nres = xhyp.shape[0]
        bestxcand = np.zeros((nres, xhyp.shape[1]))
        neglnlikecand = np.full(nres, np.inf)
        mine = list(range(nres))
        rank, size = 0, 1
        if K.get("shard_restarts", False) and nres > 1:
            from bayesianopttools_b200 import parallel as par
            rank, size = par.world()
            if size > 1:
                mine = [i for i in range(nres) if i % size == rank]
        # the ranks now do DIFFERENT work: a population-sharded likelihood (a collective per call) cannot run inside
        shard_pop = K.get("shard_population", False)
        if size > 1:
            K["shard_population"] = False
        device_loop = K["optimizer"].lower() == "cmaes" and K.get("cmaes_device", True) and xhyp.shape[1] <= 64
        if mine and K["optimizer"].lower() == "lbfgsb" and K.get("lockstep_restarts", True):
            # L-BFGS-B: every restart's reverse-communication state on this thread, one population per round
            bx, bf, self.lockstep_calls = lbfgsb_lockstep(
                lambda P: likelihood_batch(P, K, trainvar=self.trainvar), xhyp[mine, :],
                np.asarray(K["lbhyp"], dtype=float), np.asarray(K["ubhyp"], dtype=float))
            bestxcand[mine, :], neglnlikecand[mine] = bx, bf
        elif len(mine) > 1 and K.get("lockstep_restarts", True) and not device_loop:
            ls = _Lockstep(lambda P: likelihood_batch(P, K, trainvar=self.trainvar), len(mine))
            self._lockstep = ls
            errors = []

            def work(slot, ii):
                _TLS.tid = slot
                try:
                    bestxcand[ii, :], neglnlikecand[ii] = self.tune_hyperparameters(xhyp[ii, :])
                except BaseException as e:          # noqa: BLE001 -- re-raised on the calling thread
                    errors.append(e)
                finally:
                    _TLS.tid = None
                    ls.finish(slot)

            threads = [threading.Thread(target=work, args=(slot, ii)) for slot, ii in enumerate(mine)]
            try:
                for t in threads:
                    t.start()
                for t in threads:
                    t.join()
            finally:
                self._lockstep = None
            if errors:
                raise errors[0]
            self.lockstep_calls = ls.calls
        else:
            for ii in mine:
                if disp:
                    print(f"Training hyperparameter candidate no.{ii + 1}")
                bestxcand[ii, :], neglnlikecand[ii] = self.tune_hyperparameters(xhyp[ii, :])
        if size > 1:
            from bayesianopttools_b200 import parallel as par
            K["shard_population"] = shard_pop
            rows = np.hstack([bestxcand, neglnlikecand[:, None]])
            rows[[i for i in range(nres) if i not in mine]] = 0.0
            allrows = par.allgather_rows(rows.ravel()).reshape(size, nres, -1)
            for i in range(nres):
                bestxcand[i, :], neglnlikecand[i] = allrows[i % size, i, :-1], allrows[i % size, i, -1]
        return bestxcand, neglnlikecand

    def _seed(self):
        """KrigInfo["seed"], or -- when several ranks train one design together -- a seed drawn by rank 0, so that
        the stochastic optimisers (GA, CMA-ES) submit identical populations on every rank."""
        K = self.KrigInfo
        seed = K.get("seed")
        if seed is None and K.get("shard_population", False):
            from bayesianopttools_b200 import parallel as par
            seed = int(par.broadcast_array(np.array([float(np.random.randint(0, 2 ** 31 - 1))]), 0)[0])
        return seed

    def tune_hyperparameters(self, x0):
        """kriging_model.py:392-455 with batched objective evaluations."""
        K = self.KrigInfo
        lbv, ubv = np.asarray(K["lbhyp"], dtype=float), np.asarray(K["ubhyp"], dtype=float)
        opt = K["optimizer"].lower()

        def scalar(x, *a):
            f = self._batch_nll(np.asarray(x, dtype=float)[None, :])[0]
            return float(f) if np.isfinite(f) else PENALTY

        if opt == "cmaes" and K.get("cmaes_device", True) and len(x0) <= 64:
            # the whole generation loop runs on the device (kb_train_cmaes): no host round trip per generation
            from .supports.device_state import fixed_nugget_of, get_model
            seed = self._seed()
            bx, bf, _, _, _ = get_model(K).model.train_cmaes(
                x0, self.sigmacmaes, self.trainvar, fixed_nugget_of(K), lb=lbv, ub=ubv, scaling=self.scaling,
                popsize=K.get("cmaes_popsize") or 0, maxiter=K.get("cmaes_maxiter") or 0,
                seed=int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed))
            return bx, bf
        if opt == "cmaes":
            bx, bf, _ = fmin_batched(lambda P: self._batch_nll(P), x0, self.sigmacmaes, bounds=[lbv, ubv],
                                     scaling=self.scaling, popsize=K.get("cmaes_popsize"),
                                     seed=self._seed(), maxiter=K.get("cmaes_maxiter"))
            return bx, bf
        if opt == "lbfgsb":
            res = minimize(lambda x: self._fun_and_grad(x, lbv, ubv, 1e-3), x0, method="L-BFGS-B", jac=True,
                           bounds=np.column_stack((lbv, ubv)))
            return res.x, res.fun
        if opt == "slsqp":
            res = minimize(lambda x: self._fun_and_grad(x, lbv, ubv, 1.4901161193847656e-08), x0,
                           method="SLSQP", jac=True, bounds=np.column_stack((lbv, ubv)))
            return res.x, res.fun
        if opt == "cobyla":
            cons = []
            for i in range(len(ubv)):
                cons.append(lambda x, i=i: x[i] - lbv[i])
                cons.append(lambda x, i=i: ubv[i] - x[i])
            res = fmin_cobyla(scalar, x0, cons, rhobeg=0.5, rhoend=1e-4)
            return res, scalar(res)
        if opt == "ga":
            bx, bf, _ = uncGA(self._batch_nll, lb=lbv, ub=ubv, npop=K.get("ga_npop", 100),
                              maxg=K.get("ga_maxg", 100), seed=self._seed())
            return bx, bf
        raise KeyError(f"{K['optimizer']} in KrigInfo['Optimizer'] is not recognised.")

    # ------------------------------------------------------------------
    def loocvcalc(self, metrictype="mape", drm=None):
        """kriging_model.py:262-296 / krigloocv2.py:7-33: leave-one-out predictions at fixed theta.

        The reference refits n models of n-1 points; the identical numbers follow from ONE
        factorisation (Dubrule 1983): with Q = R^-1 - R^-1 F (F'R^-1F)^-1 F'R^-1,
        y_i - yhat_-i = (Q y)_i / Q_ii.  diag(R^-1), R^-1 F and Q y come from the device fit state."""
        from .supports.loocv import loocv_device
        K = self.KrigInfo
        pred = loocv_device(K)
        K["LOOCVerror"], K["LOOCVpred"] = errperf(K["y"], pred, metrictype), pred
        return K["LOOCVerror"], K["LOOCVpred"]

    def predict(self, x, predtypes=["pred"], drmmodel=None):
        """kriging_model.py:298-324."""
        return prediction(np.asarray(x, dtype=float), self.KrigInfo, predtypes=predtypes, drm=drmmodel)

    def predict_sweep(self, x, acq="EI", want=()):
        """Reduce-only sweep over many sites: arg-min of one acquisition plus the AK-MCS
        counters, without materialising per-site outputs.  Returns (result dict, outputs dict).
        `x` is an (m, d) array of raw sites or a `_lib.DevicePoints` set already on the device."""
        from .supports.prediction import _configure, _ensure_fitted
        from .supports.device_state import get_model
        K = self.KrigInfo
        holder = get_model(K)
        _ensure_fitted(K, holder)
        _configure(K, holder.model)
        kw = dict(ybest=float(np.min(holder.model.y)))
        if "limit" in K:
            kw.update(limit=float(K["limit"]), limit_sign=1 if K.get("limittype", ">=") in (">", ">=") else -1)
        if "sigmalcb" in K:
            kw["sigmalcb"] = float(K["sigmalcb"])
        from .. import _lib
        if isinstance(x, _lib.DevicePoints):
            outs, res = holder.model.predict_points(x, tuple(w.lower() for w in want), acq=acq.lower(), **kw)
        else:
            outs, res = holder.model.predict(np.asarray(x, dtype=float), tuple(w.lower() for w in want),
                                             acq=acq.lower(), **kw)
        return res, outs


def kriginfocheck(KrigInfo, lb, ub, nbhyp):
    """kriging_model.py:458-571: defaults and the hyper-parameter box."""
    eps = np.finfo(float).eps
    if "nrestart" not in KrigInfo:
        KrigInfo["nrestart"] = 1
    elif KrigInfo["nrestart"] < 1:
        raise ValueError('Minimum value of KrigInfo["nrestart"] is 1')
    if "optimizer" not in KrigInfo:
        KrigInfo["optimizer"] = "lbfgsb"
        print("The optimizer is not specified, set to lbfgsb")
    elif KrigInfo["optimizer"].lower() not in ["lbfgsb", "cmaes", "cobyla", "slsqp", "ga"]:
        raise ValueError(KrigInfo["optimizer"], " is not a valid optimizer.")
    if "TrendOrder" not in KrigInfo:
        KrigInfo["TrendOrder"] = 0
    elif KrigInfo["TrendOrder"] < 0:
        raise ValueError("The order of the polynomial trend should be a positive value.")
    nnugget = 1
    if "nugget" not in KrigInfo:
        KrigInfo["nugget"] = -6
        KrigInfo["nuggetparam"] = "fixed"
        print("Nugget is not defined, set nugget to 1e-06")
    elif type(KrigInfo["nugget"]) is not list:
        KrigInfo["nuggetparam"] = "fixed"
    elif len(KrigInfo["nugget"]) == 2:
        KrigInfo["nuggetparam"] = "tuned"
        nnugget = 2
        if KrigInfo["nugget"][0] > KrigInfo["nugget"][1]:
            raise TypeError("The lower bound of the nugget should be lower than the upper bound.")
    if "kernel" not in KrigInfo:
        KrigInfo["kernel"] = ["gaussian"]
        print("Kernel is not defined, set kernel to gaussian")
    elif type(KrigInfo["kernel"]) is not list:
        KrigInfo["kernel"] = [KrigInfo["kernel"]]
    nkernel = len(KrigInfo["kernel"])
    KrigInfo["nkernel"] = nkernel
    if KrigInfo["type"] == "kriging":
        KrigInfo["n_princomp"] = False
    if "limit" in KrigInfo and "limittype" not in KrigInfo:
        KrigInfo["limittype"] = ">="
        print("KrigInfo['limittype'] is set to greater than equal'>='.")
    lbhyp = lb * np.ones(nbhyp)
    ubhyp = ub * np.ones(nbhyp)
    if nnugget > 1:
        lbhyp = np.hstack((lbhyp, KrigInfo["nugget"][0]))
        ubhyp = np.hstack((ubhyp, KrigInfo["nugget"][1]))
    if nkernel > 1:
        lbhyp = np.hstack((lbhyp, np.zeros(nkernel) + eps))
        ubhyp = np.hstack((ubhyp, np.ones(nkernel)))
    scaling = np.ones(len(lbhyp))
    if nnugget > 1:
        scaling[-1 - (nkernel if nkernel > 1 else 0)] = (KrigInfo["nugget"][1] - KrigInfo["nugget"][0]) / (ub - lb)
    if nkernel > 1:
        scaling[-nkernel:] = 1 / (ub - lb)
    KrigInfo["lbhyp"] = lbhyp
    KrigInfo["ubhyp"] = ubhyp
    return KrigInfo, scaling
