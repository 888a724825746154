"""Sobol sensitivity indices on a Kriging surrogate (reference: sensitivity_analysis/sobol_ind.py:6-172).

SURVEY.md section 8(f) row 2: a drop-in beneficiary of the batched predictor.  Same class, arguments,
attributes and estimators as the reference; every Monte-Carlo matrix (A, B, the nvar first/total-order
matrices C_i and the nvar(nvar-1)/2 second-order matrices C_ij, 2e4 rows each) goes through ONE
`krigobj.predict(..., ['pred'])` call = one fused sweep on the device, instead of the reference's
10 000-row chunk loops (sobol_ind.py:48-58,114-124,154-164), each of which re-factorises R twice.
"""
import numpy as np

from ..misc.sampling.samplingplan import sampling
from ..testcase.RA.testcase import evaluate


class SobolIndices:
    """Create Sobol indices for sensitivity analysis (sobol_ind.py:6-34)."""

    def __init__(self, nvar, krigobj=None, problem=None, ub=None, lb=None):
        self.nvar = nvar
        self.krigobj = krigobj
        self.problem = problem
        self.n = int(2e4)
        if ub is not None and lb is not None:
            _, init_mat = sampling("sobolnew", self.nvar * 2, self.n, result="real", upbound=ub, lobound=lb)
        else:
            init_mat, _ = sampling("sobolnew", self.nvar * 2, self.n)
        self.A = np.ascontiguousarray(init_mat[:, :self.nvar])
        self.B = np.ascontiguousarray(init_mat[:, self.nvar:])
        self.ya = None
        self.yb = None
        self.fo_2 = None
        self.denom = None

    def _model(self, X):
        """Responses at the rows of X as an (n, 1) column (sobol_ind.py:45-76)."""
        if self.krigobj is not None:
            return np.asarray(self.krigobj.predict(X, ["pred"]), dtype=float).reshape(-1, 1)
        if self.problem is not None:
            return np.asarray(evaluate(X, self.problem), dtype=float).reshape(-1, 1)
        raise ValueError("Either krigobj or problem must be not None")

    def analyze(self, first=True, total=False, second=False):
        """sobol_ind.py:36-95 -- returns {'first', 'total', 'second'} as requested."""
        self.ya = self._model(self.A)
        self.yb = self._model(self.B)
        self.fo_2 = (np.sum(self.ya) / self.n) ** 2
        self.denom = (np.sum(self.ya ** 2) / self.n) - self.fo_2
        indices = dict()
        if first is True or total is True:
            indices["first"], indices["total"] = self.calc_ft_order(first, total)
        if second is True:
            if "first" not in indices:
                # the reference's `elif second and not first` branch is unreachable (sobol_ind.py:87-91);
                # the first-order indices it needs are computed here instead of raising KeyError
                indices["first"], indices["total"] = self.calc_ft_order(True, True)
            indices["second"] = self.calc_second_order(indices["first"])
        return indices

    def calc_ft_order(self, first=True, total=False):
        """First / total order indices (sobol_ind.py:97-134)."""
        s1 = np.zeros(self.nvar)
        st = np.zeros(self.nvar)
        for ii in range(self.nvar):
            C_i = self.B.copy()
            C_i[:, ii] = self.A[:, ii]
            yci = self._model(C_i)
            if first:
                s1[ii] = ((1 / self.n) * np.sum(self.ya * yci) - self.fo_2) / self.denom
            if total:
                st[ii] = 1 - (((1 / self.n) * np.sum(self.yb * yci) - self.fo_2) / self.denom)
        return [s1, st]

    def calc_second_order(self, s1):
        """Second order indices (sobol_ind.py:136-172)."""
        s2 = dict()
        for ii in range(self.nvar - 1):
            for jj in range(ii + 1, self.nvar):
                C_ij = self.B.copy()
                C_ij[:, ii] = self.A[:, ii]
                C_ij[:, jj] = self.A[:, jj]
                yci = self._model(C_ij)
                vij = (1 / self.n) * np.sum(self.ya * yci) - self.fo_2
                s2["x" + str(ii + 1) + "-x" + str(jj + 1)] = (vij / self.denom) - s1[ii] - s1[jj]
        return s2
