"""Batched predictor / acquisition on the GPU (reference: surrogate_models/supports/prediction.py)."""
import numpy as np

from bayesianopttools_b200 import _lib
from .device_state import get_model
from .likelihood_func import _fit_signature

_VALID = ("pred", "ssqr", "s", "fpc", "lcb", "ebe", "ei", "poi", "pof")


def _ensure_fitted(KrigInfo, holder):
    """The device predictor state must match KrigInfo's Theta / SigmaSqr / nugget / wgkf
    (the reference reads them from the dict on every call, prediction.py:144-150).  If the
    dict was changed since the last device fit (or unpickled), refit at those values."""
    sig = _fit_signature(KrigInfo)
    if holder.fit_sig == sig and holder.model.fitted:
        return
    theta = np.asarray(KrigInfo["Theta"], dtype=float).ravel()
    s2 = float(np.asarray(KrigInfo["SigmaSqr"]).ravel()[0])
    nug = float(np.asarray(KrigInfo["nugget"]).ravel()[0])
    x = np.concatenate([theta, [np.log10(s2)], [nug]])
    if int(KrigInfo["nkernel"]) > 1:
        x = np.concatenate([x, np.asarray(KrigInfo["wgkf"], dtype=float).ravel()])
    # KrigInfo["predict_refine"]: 'auto' (default) / False / True -- corrected sweeps on ill-conditioned fits
    holder.model.set_predict_refine(KrigInfo.get("predict_refine", "auto"))
    holder.model.fit_state(x, True, nug, want_U=False, want_Psi=False)
    holder.fit_sig = sig


def _configure(KrigInfo, model):
    if KrigInfo["standardization"] is True:
        if KrigInfo["normtype"] == "default":
            model.set_norm(_lib.NORM_DEFAULT, KrigInfo["lb"], KrigInfo["ub"])
        elif KrigInfo["normtype"] == "std":
            model.set_norm(_lib.NORM_STD, KrigInfo["X_mean"], KrigInfo["X_std"])
        else:
            raise ValueError(f"Kriging model dictionary 'normtype' value: '{KrigInfo['normtype']}' is not recognised.")
    else:
        model.set_norm(_lib.NORM_NONE)
    model.set_trend(np.asarray(KrigInfo["idx"]).astype(np.int32))
    # norm_y: the reference maps the mean to real units before anything is derived from it (prediction.py:213-217);
    # the device applies the same map inside the sweep, so acquisitions, U and the AK-MCS counters see real units
    if KrigInfo.get("norm_y", False) is True:
        if KrigInfo["normtype"] == "default":
            y = np.asarray(KrigInfo["y"], dtype=float)
            model.set_ymap(_lib.YMAP_DEFAULT, float(np.max(y) - np.min(y)), float(np.min(y)))
        else:
            model.set_ymap(_lib.YMAP_STD, float(np.ravel(KrigInfo["y_std"])[0]), float(np.ravel(KrigInfo["y_mean"])[0]))
    else:
        model.set_ymap(_lib.YMAP_NONE)


def stdtoreal(f, KrigInfo, num=None):
    """prediction.py:300-313"""
    if KrigInfo["normtype"] == "default":
        y = KrigInfo["y"] if num is None else KrigInfo["y"][num]
        ymax, ymin = np.max(y), np.min(y)
        return (f / 2 + 0.5) * (ymax - ymin) + ymin
    return KrigInfo["y_mean"] + KrigInfo["y_std"] * f


def prediction(x, KrigInfo, predtypes, num=None, drm=None, **kwargs):
    """prediction.py:67-298 -- same call and return conventions:
    'pred','ei','poi','pof','fpc','lcb' -> (m,1);  'SSqr','s','ebe' -> (1,m);
    one requested output -> scalar (`.item()`) when m == 1, else the array; several -> list.

    Per-point EI (the reference's whole-batch degenerate cases, :250-251,:264-266, are not
    reproduced); 'lcb' returns f - sigmalcb*SSqr per point as (m,1)."""
    if num is not None or drm is not None:
        raise NotImplementedError("multi-objective 'num' and KPCA 'drm' are outside the Kriging hot path")
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = np.array([x])
    if isinstance(predtypes, str):
        predtypes = [predtypes]
    names = [p.lower() for p in predtypes]
    for p, raw in zip(names, predtypes):
        if p not in _VALID:
            raise ValueError(f"Specified prediction type: '{raw}' is not recognised.")
    holder = get_model(KrigInfo)
    _ensure_fitted(KrigInfo, holder)
    model = holder.model
    _configure(KrigInfo, model)
    y = model.y
    kw = dict(ybest=float(np.min(y)))
    if "pof" in names:
        lt = KrigInfo["limittype"]
        if lt in (">", ">="):
            kw["limit_sign"] = 1
        elif lt in ("<", "<="):
            kw["limit_sign"] = -1
        else:
            raise ValueError("Limit Type is not available yet!")
        kw["limit"] = float(KrigInfo["limit"])
    if "lcb" in names:
        kw["sigmalcb"] = float(KrigInfo["sigmalcb"])
    outs, _ = model.predict(x, tuple(dict.fromkeys(names)), acq=names[0] if names[0] != "fpc" else "pred",
                            want_result=False, **kw)
    outputs = []
    for p in names:
        v = outs[p]
        outputs.append(v.reshape(1, -1) if p in ("ssqr", "s", "ebe") else v.reshape(-1, 1))
    if len(outputs) == 1:
        try:
            return outputs[0].item()
        except ValueError:
            return outputs[0]
    return outputs
