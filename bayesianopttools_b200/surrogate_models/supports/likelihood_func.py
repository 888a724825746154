"""Negative ln-likelihood on the GPU (reference: surrogate_models/supports/likelihood_func.py).

``likelihood`` keeps the reference signature; ``likelihood_batch`` is the population form the
optimisers use (one C-ABI call per CMA-ES generation / finite-difference stencil).
"""
import numpy as np

from bayesianopttools_b200 import _lib
from .device_state import effective_nvar, fixed_nugget_of, get_model, nuggetset


def _prep_x(x, KrigInfo):
    nvar = int(KrigInfo["nvar"])
    if KrigInfo["kernel"] == ["iso_gaussian"]:
        x = np.array([np.asarray(x, dtype=float).ravel()[0]] * nvar)       # likelihood_func.py:60-61
    if not isinstance(x, np.ndarray):
        x = np.array([x], dtype=float) if np.ndim(x) == 0 else np.asarray(x, dtype=float)
    return np.asarray(x, dtype=float).reshape(-1)


def likelihood_batch(xpop, KrigInfo, trainvar=True, return_info=False):
    """NLL for every row of xpop (B x nbhyp) in one launch sequence.  Candidates whose
    correlation matrix is not positive definite get +inf (info != 0) instead of raising."""
    holder = get_model(KrigInfo)
    xpop = np.atleast_2d(np.asarray(xpop, dtype=float))
    if KrigInfo["kernel"] == ["iso_gaussian"]:
        xpop = np.repeat(xpop[:, :1], int(KrigInfo["nvar"]), axis=1)
    if KrigInfo.get("shard_population", False):
        # candidate b is evaluated by rank b mod G, the NLL slices (8 B per candidate) are
        # all-gathered.  Opt-in, because EVERY rank must then make the same calls with the same population (the
        # optimisers here are deterministic, so ranks that train the same design stay in lockstep).
        from bayesianopttools_b200 import parallel
        rank, size = parallel.world()
        if size > 1 and len(xpop) >= size:
            mine = parallel.candidate_slice(len(xpop), rank, size)
            nll, info = holder.model.nll_batch(xpop[mine], trainvar, fixed_nugget_of(KrigInfo), return_info=True)
            # one collective per call: NLL slices, info flags and the same-population check together
            full, flags = parallel.gather_sharded_nll(nll, info, xpop)
            return (full, flags) if return_info else full
        if size > 1:
            parallel.assert_same_population(xpop)         # every rank, every call: same collective sequence
    return holder.model.nll_batch(xpop, trainvar, fixed_nugget_of(KrigInfo), return_info=return_info)


def likelihood(x, KrigInfo, mode="default", trainvar=True):
    """likelihood_func.py:6-118.  mode 'default' -> NegLnLike ; 'all' -> KrigInfo with
    Theta, U, Psi, BE, SigmaSqr, NegLnLike, nugget, wgkf written (:82-84,107-111)."""
    if mode.lower() not in ("default", "all"):
        raise TypeError("Only have two modes, default and all, default return NegLnLike, all return KrigInfo")
    x = _prep_x(x, KrigInfo)
    nvar = effective_nvar(KrigInfo)
    nkernel = int(KrigInfo["nkernel"])
    nugget, _, wgkf = nuggetset(x, KrigInfo, nvar, nkernel, trainvar)
    holder = get_model(KrigInfo)
    if mode.lower() == "default":
        nll, info = holder.model.nll_batch(x[None, :], trainvar, fixed_nugget_of(KrigInfo), return_info=True)
        if info[0] != 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")      # likelihood_func.py:175
        return float(nll[0])
    try:
        holder.model.set_predict_refine(KrigInfo.get("predict_refine", "auto"))
        st = holder.model.fit_state(x, trainvar, fixed_nugget_of(KrigInfo))
    except _lib.KrigB200Error as e:
        if e.status == _lib.ERR_NOT_PD:
            raise np.linalg.LinAlgError("Matrix is not positive definite") from e
        raise
    KrigInfo["Theta"] = x[0:nvar]
    KrigInfo["nugget"] = nugget
    KrigInfo["wgkf"] = wgkf
    KrigInfo["U"] = st["U"]
    KrigInfo["Psi"] = st["Psi"]
    KrigInfo["BE"] = st["BE"]
    KrigInfo["SigmaSqr"] = st["SigmaSqr"] if trainvar else np.array([[st["SigmaSqr"]]])
    KrigInfo["NegLnLike"] = st["NegLnLike"]
    holder.fit_sig = _fit_signature(KrigInfo)
    return KrigInfo


def _fit_signature(KrigInfo):
    return (np.asarray(KrigInfo["Theta"], dtype=float).tobytes(), float(np.asarray(KrigInfo["SigmaSqr"]).ravel()[0]),
            float(np.asarray(KrigInfo["nugget"]).ravel()[0]),
            np.asarray(KrigInfo["wgkf"], dtype=float).tobytes())
