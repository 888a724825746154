"""AK-MCS reliability analysis (reference: reliability_analysis/akmcs.py:9-290).

Per update the reference predicts (mean, s) over the whole Monte-Carlo population in
10 000-row chunks and then runs three Python list comprehensions over N for the failure
counts.  Here one fused launch returns pred, s and the reductions: #{G<=0},
#{G-1.96s<=0}, #{G+1.96s<=0}, min U and its first index (csrc/kb_predict.cu).

Several GPUs (BASELINE configs[4]): under torch.distributed every rank holds
one contiguous shard of the Monte-Carlo population as its `init_samp`, resident on its GPU across
all updates.  Per update each rank sweeps its shard; the three int64 failure counters are summed with one
all-reduce, and one all-gather carries every rank's (min U, index) record together with the d coordinates
of its own lowest-U site (merged with the lowest whole-population index on ties,
`parallel.gather_sweep_result`), so the new site arrives without a separate broadcast; every rank retrains
the (small) model redundantly from identical inputs.
"""
import time

import numpy as np

from bayesianopttools_b200 import parallel
from bayesianopttools_b200.misc.sampling.samplingplan import realval
from bayesianopttools_b200.testcase.RA.testcase import evaluate


class AKMCS:
    def __init__(self, krigobj, akmcsInfo):
        akmcsInfo = akmcsInfocheck(akmcsInfo)
        self.krigobj = krigobj
        self.akmcsInfo = akmcsInfo
        self.init_samp = akmcsInfo["init_samp"]         # (N, d) array, or a _lib.DevicePoints set
        self.maxupdate = akmcsInfo["maxupdate"]
        self.nsamp_local = self.init_samp.n if hasattr(self.init_samp, "_h") else np.size(self.init_samp, axis=0)
        # akmcsInfo["shard_population"] (default: on whenever torch.distributed runs more than one rank):
        # init_samp is this rank's shard, nsamp the size of the whole population
        self._rank, world_size = parallel.world()
        self.sharded = world_size > 1 and bool(akmcsInfo.get("shard_population", True))
        self._counts = parallel.gather_counts(self.nsamp_local) if self.sharded else [self.nsamp_local]
        self._offset = int(sum(self._counts[:self._rank])) if self.sharded else 0
        self.nsamp = int(sum(self._counts))
        self.Gx = np.zeros((self.nsamp_local, 1))
        self.sigmaG = np.zeros((1, self.nsamp_local))
        self.stop_criteria = 100
        self.keep_fields = akmcsInfo.get("keep_fields", True)
        # the Monte-Carlo population is the same in every update (akmcs.py:66-206): it is uploaded
        # to the device once and every sweep is one launch (`resident_population`, default on)
        self._points = None
        self.resident = akmcsInfo.get("resident_population", True)

    def _population(self):
        if hasattr(self.init_samp, "_h"):            # already a device-resident point set
            return self.init_samp
        if not self.resident:
            return self.init_samp
        if self._points is None:
            from .. import _lib
            self._points = _lib.DevicePoints(self.init_samp)
        return self._points

    def _sweep(self):
        want = ("pred", "s") if self.keep_fields else ()
        res, outs = self.krigobj.predict_sweep(self._population(), acq="u", want=want)
        if self.keep_fields:
            self.Gx = outs["pred"].reshape(-1, 1)
            self.sigmaG = outs["s"].reshape(1, -1)
        if self.sharded:
            res = dict(res)
            res["min_u_idx"] += self._offset          # indices of the whole population
            res["best_idx"] += self._offset
            # every rank attaches the coordinates of its own lowest-U site: the winner's arrive with the records
            mine = int(res["min_u_idx"]) - self._offset
            row = self.init_samp[mine, :] if 0 <= mine < self.nsamp_local else np.zeros(self.krigobj.KrigInfo["nvar"])
            res = parallel.gather_sweep_result(res, payload=np.asarray(row, dtype=float))
            xnew = res.pop("min_u_payload")
        else:
            xnew = self.init_samp[res["min_u_idx"], :]
        self._res = res
        N = self.nsamp
        self.Pf = res["n_fail"] / N                                   # pfcalc   akmcs.py:208-212
        self.minU = res["min_u"]                                       # lfucalc  akmcs.py:222-226
        self.xnew = xnew
        pf0 = res["n_fail"] / N                                        # stopcrit akmcs.py:228-238
        self.stop_criteria = 100 if pf0 == 0 else (res["n_fail_minus"] / N - res["n_fail_plus"] / N) / pf0
        if self.keep_fields:
            self.U = np.abs(self.Gx) / self.sigmaG.reshape(-1, 1)
        self.xnew = np.asarray(self.xnew, dtype=float).reshape(-1)

    def covpf(self):
        """akmcs.py:214-220"""
        return 1000 if self.Pf == 0 else np.sqrt((1 - self.Pf) / (self.Pf * self.nsamp))

    def run(self, autoupdate=True, disp=True, savedatato=None, logging=False, saveimageto=None,
            plotdatapos=None, plotdataneg=None):
        """akmcs.py:49-206 (no comet.ml logging, no plotting)."""
        evalfun = self.akmcsInfo.get("evaluate", lambda x: evaluate(x, type=self.akmcsInfo["problem"]))
        self._sweep()
        self.updateX = np.array([self.xnew])
        self.minUiter = np.array([self.minU])
        if disp:
            print(f"Done iter no: 0, Pf: {self.Pf}, minU: {self.minU}")
        history = []
        while autoupdate:
            for i in range(self.maxupdate):
                t = time.time()
                ynew = np.asarray(evalfun(self.xnew), dtype=float).reshape(1, -1)
                K = self.krigobj.KrigInfo
                K["y"] = np.vstack((K["y"], ynew))
                K["X"] = np.vstack((K["X"], self.xnew))
                K["nsamp"] += 1
                self.krigobj.standardize()
                self.krigobj.train(disp=False)
                self._sweep()
                self.cov = self.covpf()
                self.updateX = np.vstack((self.updateX, self.xnew))
                self.minUiter = np.vstack((self.minUiter, self.minU))
                elapsed = time.time() - t
                history.append([i, self.Pf, self.stop_criteria, elapsed])
                if disp:
                    print(f"iter no: {i + 1}, Pf: {self.Pf}, stopcrit: {self.stop_criteria}, "
                          f"time(s): {elapsed}, ynew: {ynew}")
                if savedatato is not None:
                    np.savetxt(savedatato, np.array(history), delimiter=",", header="iter,Pf,stopcrit,time(s)")
                if self.stop_criteria <= 0.05 and i >= 15:
                    break
            if disp:
                print(f"COV: {self.cov}")
            break          # akmcs.py:206 ("temporary break")


def mcpopgen_device(lb=None, ub=None, n_order=6, n_coeff=1, type="random", ndim=2, stddev=1, mean=0, seed=0,
                    first_index=0, shard=False):
    """mcpopgen (akmcs.py:241-260) with the population generated and kept on the device: same arguments
    and distributions, counter-based Philox stream instead of numpy's global Mersenne Twister (so the
    sample differs from numpy's, like any re-seeding would); returns a `_lib.DevicePoints` set that AKMCS
    accepts as `init_samp`.  `shard=True` under torch.distributed: this rank generates only its contiguous
    share of the same stream (site i has the same coordinates whatever the number of ranks)."""
    from .. import _lib
    nmc = int(n_coeff * 10 ** n_order)
    if shard:
        rank, size = parallel.world()
        start, stop = parallel.point_shard(nmc, rank, size)
        first_index, nmc = first_index + start, stop - start
    if type.lower() == "random":
        if lb is None or ub is None:
            raise ValueError("type 'random' is selected, please input lower bound and upper bound value")
        ndim = len(lb)
    if type.lower() not in _lib.DevicePoints.DISTS:
        raise ValueError("Monte Carlo sampling type not supported")
    return _lib.DevicePoints.generate(nmc, ndim, dist=type, mean=mean, stddev=stddev, lb=lb, ub=ub, seed=seed,
                                      first_index=first_index)


def mcpopgen(lb=None, ub=None, n_order=6, n_coeff=1, type="random", ndim=2, stddev=1, mean=0):
    """akmcs.py:241-260"""
    nmc = int(n_coeff * 10 ** n_order)
    t = type.lower()
    if t in ("normal", "gaussian"):
        return stddev * np.random.randn(nmc, ndim) + mean
    if t == "lognormal":
        var = stddev ** 2
        sigma = np.sqrt(np.log(var / (mean ** 2) + 1))
        mu = np.log((mean ** 2) / np.sqrt(var + mean ** 2))
        return np.exp(sigma * np.random.randn(nmc, ndim) + mu)
    if t == "gumbel":
        return np.random.gumbel(mean, (stddev / np.pi) * np.sqrt(6), (nmc, ndim))
    if t == "random":
        if lb is None or ub is None:
            raise ValueError("type 'random' is selected, please input lower bound and upper bound value")
        return realval(lb, ub, np.random.rand(nmc, len(lb)))
    raise ValueError("Monte Carlo sampling type not supported")


def akmcsInfocheck(akmcsInfo):
    """akmcs.py:263-290"""
    if "init_samp" not in akmcsInfo:
        raise ValueError('akmcsInfo["init_samp"] must be defined')
    if "maxupdate" not in akmcsInfo:
        akmcsInfo["maxupdate"] = 120
    if "problem" not in akmcsInfo and "evaluate" not in akmcsInfo:
        raise ValueError('akmcsInfo["problem"] must be defined')
    return akmcsInfo
