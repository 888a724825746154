"""Leave-one-out cross validation on the device (reference: supports/krigloocv2.py:7-33)."""
import numpy as np

from .device_state import get_model
from .prediction import _ensure_fitted


def loocv_device(KrigInfo):
    holder = get_model(KrigInfo)
    _ensure_fitted(KrigInfo, holder)
    pred = holder.model.loocv().reshape(-1, 1)
    if KrigInfo.get("norm_y", False) is True:
        # the model was trained on normalised responses: back to the units of KrigInfo["y"], which the error
        # metric is taken against (the reference's loocv2 crashes on such a model, krigloocv2.py:18-27)
        from .prediction import stdtoreal
        pred = stdtoreal(pred, KrigInfo)
    return pred


def loocv2(KrigInfo, errtype="rmse"):
    from .errperf import errperf
    pred = loocv_device(KrigInfo)
    return errperf(KrigInfo["y"], pred, errtype), pred
