"""`ehvicalc` drop-in (reference optim_tools/ehvi/EHVIcomputation.py:10-51).

The reference evaluates one design at a time: two Kriging predictions, then `exi2d` with the predictive
VARIANCE handed over as the spread `s` (EHVIcomputation.py:40-45 fills it from 'SSqr'); that is kept.
Here x may also be an (m, d) batch: both objective models sweep all candidates on the device and the
improvement is computed from the predictions without a trip through the host (kb_ehvi2d).

One deliberate difference: the reference replaces an exactly-zero improvement by a RANDOM value in
[tiny, 100 tiny] to keep CMA-ES from seeing a flat population; the replacement here is the deterministic
value `tiny`, so that results are reproducible (same sign, same order of magnitude)."""
import numpy as np

from ... import _lib


def _device_models(kriglist):
    from ...surrogate_models.supports.prediction import _configure, _ensure_fitted
    from ...surrogate_models.supports.device_state import get_model
    models = []
    for obj in kriglist:
        K = obj.KrigInfo
        holder = get_model(K)
        _ensure_fitted(K, holder)
        _configure(K, holder.model)
        models.append(holder.model)
    return models


def ehvicalc(x, ypar, moboInfo, kriglist):
    return EHVI(x, ypar, moboInfo, kriglist)


def EHVI(x, ypar, moboInfo, kriglist):
    if len(kriglist) != 2:
        raise NotImplementedError("the device EHVI handles two objectives (the reference's exi2d)")
    x = np.asarray(x, dtype=float)
    m1, m2 = _device_models(kriglist)
    xs = x.reshape(1, -1) if x.ndim == 1 else x
    vals, _, _ = _lib.ehvi2d_models(m1, m2, xs, ypar, moboInfo["refpoint"], spread="ssqr")
    HV = -vals
    HV[HV == 0] = np.finfo("float").tiny
    return float(HV[0]) if x.ndim == 1 else HV


def ehvi_best(x, ypar, moboInfo, kriglist):
    """Candidate of the batch x ((m, d) array or a device-resident point set) with the largest expected improvement: (index, value); only the
    16-byte reduction leaves the device."""
    m1, m2 = _device_models(kriglist)
    xs = x if isinstance(x, _lib.DevicePoints) else np.asarray(x, dtype=float)
    _, idx, val = _lib.ehvi2d_models(m1, m2, xs, ypar, moboInfo["refpoint"], spread="ssqr", want_values=False)
    return idx, val
