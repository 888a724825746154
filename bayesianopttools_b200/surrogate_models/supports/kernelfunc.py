"""calckernel on the GPU (reference: surrogate_models/supports/kernelfunc.py:4-34)."""
import numpy as np

from bayesianopttools_b200 import _lib


def calckernel(XN, XM, theta, nvar, **kwargs):
    """Correlation matrix (n x m) for 'gaussian', 'exponential', 'matern32', 'matern52'.

    Same signature and errors as the reference; ``plscoeff`` selects the KPLS-weighted
    Gaussian (kernelfunc.py:43-49).  Computed by kb_corr_matrix (csrc/kb_corr.cu).
    """
    ker = kwargs.get("type", "gaussian")
    pls = kwargs.get("plscoeff", None)
    if ker.lower() not in ("gaussian", "iso_gaussian", "exponential", "matern32", "matern52"):
        raise ValueError("Kernel option not available!")
    if pls is not None and not np.asarray(pls).any():
        pls = None
    if pls is not None and ker.lower() not in ("gaussian", "iso_gaussian"):
        raise IndexError("KPLS weights are only defined for the gaussian kernel (kernelfunc.py:43-49)")
    return _lib.corr_matrix(np.asarray(XN, dtype=float), np.asarray(XM, dtype=float), theta, ker, pls)
