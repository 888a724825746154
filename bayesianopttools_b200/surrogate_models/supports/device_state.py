"""Binding between the reference's mutable ``KrigInfo`` dict and a device model handle.

The reference keeps ALL state in KrigInfo (numpy arrays) and recomputes everything per call.
Here the design (X_norm, y, F) lives on the GPU behind a ``kb_model`` handle that is cached
in ``KrigInfo['_kb']``; the cache is keyed by the identity of the arrays it was built from,
so `standardize()` after a design update (SOBO.py:84-89, akmcs.py:110-117) transparently
rebuilds it.  The holder deep-copies / pickles to an empty holder (KrigInfo stays picklable
and `deepcopy`-able as in krigloocv2.py:16).
"""
import numpy as np

from bayesianopttools_b200 import _lib


class _Holder:
    def __init__(self):
        self.model = None
        self.key = None
        self.refs = None
        self.fit_sig = None

    def __deepcopy__(self, memo):
        return _Holder()

    def __reduce__(self):
        return (_Holder, ())

    def drop(self):
        if self.model is not None:
            self.model.close()
        self.model, self.key, self.refs, self.fit_sig = None, None, None, None


def design_arrays(KrigInfo):
    """X, y selection of likelihood_func.py:45-53 / prediction.py:134-142."""
    if KrigInfo["standardization"] is False:
        return KrigInfo["X"], KrigInfo["y"]
    X = KrigInfo["X_norm"]
    if "y_norm" in KrigInfo and np.asarray(KrigInfo["y_norm"]).ravel()[0] != 0:
        return X, KrigInfo["y_norm"]
    return X, KrigInfo["y"]


def get_model(KrigInfo):
    """The DeviceModel for the current design in KrigInfo (built on first use)."""
    X, y = design_arrays(KrigInfo)
    F = KrigInfo["F"]
    kernels = list(KrigInfo["kernel"])
    mtype = KrigInfo["type"].lower()
    pls = KrigInfo.get("plscoeff") if mtype == "kpls" else None
    key = (id(X), np.shape(X), id(y), id(F), np.shape(F), tuple(kernels), mtype,
           id(pls) if pls is not None else None)
    holder = KrigInfo.get("_kb")
    if not isinstance(holder, _Holder):
        holder = _Holder()
        KrigInfo["_kb"] = holder
    if holder.model is None or holder.key != key:
        holder.drop()
        holder.model = _lib.DeviceModel(X, y, F, kernels, model_type=mtype, plscoeff=pls)
        holder.key = key
        holder.refs = (X, y, F, pls)     # keep the arrays alive so ids stay unique
    return holder


def effective_nvar(KrigInfo):
    """likelihood_func.py:67-68"""
    if KrigInfo.get("n_princomp", False) is not False:
        return int(KrigInfo["n_princomp"])
    return int(KrigInfo["nvar"])


def nuggetset(x, KrigInfo, nvar, nkernel, trainvar):
    """Host restatement of likelihood_func.py:121-165 (layout of x by length)."""
    off = nvar + (1 if trainvar else 0)
    rest = len(x) - off
    if rest == 0:
        nugget, wgkf = KrigInfo["nugget"], np.array([1])
    elif rest == 1:
        nugget, wgkf = x[off], np.array([1])
    elif rest == nkernel:
        nugget = KrigInfo["nugget"]
        wgkf = x[off:off + nkernel] / np.sum(x[off:off + nkernel])
    elif rest == nkernel + 1:
        nugget = x[off]
        wgkf = x[off + 1:off + 1 + nkernel] / np.sum(x[off + 1:off + 1 + nkernel])
    else:
        raise ValueError("Nugget setting is not available")
    return nugget, 10.0 ** nugget, wgkf


def fixed_nugget_of(KrigInfo):
    nug = KrigInfo.get("nugget", -6)
    if isinstance(nug, (list, tuple, np.ndarray)):
        return float(np.asarray(nug, dtype=float).ravel()[0])
    return float(nug)
