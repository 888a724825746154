"""Error metrics used by loocvcalc (reference: surrogate_models/supports/errperf.py)."""
import numpy as np


def errperf(T, P, type="rmse"):
    T, P = np.asarray(T, dtype=float), np.asarray(P, dtype=float)
    e = T - P
    t = type.lower()
    with np.errstate(divide="ignore", invalid="ignore"):
        re = e / T
    table = {
        "e": lambda: e, "ae": lambda: np.abs(e), "mae": lambda: np.mean(np.abs(e)),
        "se": lambda: e ** 2, "mse": lambda: np.mean(e ** 2), "rmse": lambda: np.sqrt(np.mean(e ** 2)),
        "re": lambda: re, "are": lambda: np.abs(re), "mare": lambda: np.mean(np.abs(re)),
        "sre": lambda: re ** 2, "msre": lambda: np.mean(re ** 2), "rmsre": lambda: np.sqrt(np.mean(re ** 2)),
        "pe": lambda: re * 100, "ape": lambda: np.abs(re) * 100, "mape": lambda: np.mean(np.abs(re) * 100),
        "spe": lambda: (re * 100) ** 2, "mspe": lambda: np.mean((re * 100) ** 2),
        "rmspe": lambda: np.sqrt(np.mean((re * 100) ** 2)),
        "r2": lambda: 1 - np.sum(e ** 2) / np.sum((T - np.mean(T)) ** 2),
    }
    if t not in table:
        raise ValueError("Error type is not available")
    return table[t]()
