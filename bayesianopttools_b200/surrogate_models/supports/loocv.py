"""Leave-one-out cross validation on the device (reference: supports/krigloocv2.py:7-33)."""
import numpy as np

from .device_state import get_model
from .prediction import _ensure_fitted


def loocv_device(KrigInfo):
    holder = get_model(KrigInfo)
    _ensure_fitted(KrigInfo, holder)
    return holder.model.loocv().reshape(-1, 1)


def loocv2(KrigInfo, errtype="rmse"):
    from .errperf import errperf
    pred = loocv_device(KrigInfo)
    return errperf(KrigInfo["y"], pred, errtype), pred
