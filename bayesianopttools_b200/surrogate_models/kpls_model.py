"""KPLS model class (reference: surrogate_models/kpls_model.py).

PLS1 rotations are computed on the host (O(h n d), once per standardize(); the reference calls
sklearn's NIPALS, kpls_model.py:102-108); the kernel then folds w^2/theta^2 into per-dimension
Gaussian weights on the device (csrc/kb_corr.cu kb_decode_kernel, kernelfunc.py:43-49).
"""
import numpy as np

from .kriging_model import Kriging, kriginfocheck


def pls1_rotations(X, y, h):
    """x_rotations_ of PLSRegression(h, scale=True) for a single response (NIPALS PLS1)."""
    X = np.asarray(X, dtype=float)
    y = np.asarray(y, dtype=float).reshape(-1)
    Xc = X - X.mean(0)
    sd = Xc.std(0, ddof=1)
    sd[sd == 0] = 1.0
    Xc = Xc / sd
    yc = y - y.mean()
    ysd = yc.std(ddof=1)
    yc = yc / (ysd if ysd != 0 else 1.0)
    W, P = np.zeros((X.shape[1], h)), np.zeros((X.shape[1], h))
    for k in range(h):
        w = Xc.T @ yc
        w = w / np.sqrt(w @ w)
        t = Xc @ w
        tt = t @ t
        p = Xc.T @ t / tt
        q = yc @ t / tt
        Xc = Xc - np.outer(t, p)
        yc = yc - t * q
        W[:, k], P[:, k] = w, p
    return W @ np.linalg.pinv(P.T @ W)


class KPLS(Kriging):
    def __init__(self, KrigInfo, ub=5, lb=-5, standardization=False, standtype="default", normy=True,
                 trainvar=True):
        if "n_princomp" not in KrigInfo:
            KrigInfo["n_princomp"] = 1
        self.n_princomp = KrigInfo["n_princomp"]
        super().__init__(KrigInfo, ub, lb, standardization, standtype, normy, trainvar, inherit=True)
        self.nbhyp = self.n_princomp + 1 if trainvar is True else self.n_princomp
        self.type = "kpls"
        KrigInfo["type"] = self.type
        KrigInfo, scaling = kriginfocheck(KrigInfo, lb, ub, self.nbhyp)
        for k in KrigInfo["kernel"]:
            if k.lower() not in ("gaussian",):
                raise IndexError("KPLS supports the gaussian kernel only (kernelfunc.py:43-49)")
        self.KrigInfo = KrigInfo
        self.scaling = scaling
        self.sigmacmaes = (ub - lb) / 5
        if self.standardization is True:
            self.standardize()

    def standardize(self):
        """kpls_model.py:93-108."""
        Kriging.standardize(self)
        K = self.KrigInfo
        Xs = K["X_norm"] if self.standardization is True else K["X"]
        K["plscoeff"] = pls1_rotations(Xs, K["y"], self.n_princomp)
