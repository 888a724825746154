"""ParEGO scalarisation (reference: optim_tools/parego.py:5-31 `paregopre`): augmented Tchebycheff
max_j(w_j f_j) + 0.05 sum_j(w_j f_j) of the objectives scaled to [0, 1], with w one of the 11 evenly
spaced weight pairs (random unless `idx` is given)."""
import numpy as np


def paregopre(yall, idx=False, rng=None):
    yall = np.asarray(yall, dtype=float)
    w1 = np.arange(11) / 10.0
    WGA = np.column_stack((w1, 1.0 - w1))
    if idx is False:
        rng = np.random.default_rng() if rng is None else rng
        idx = int(rng.integers(0, 11))
    WG = WGA[int(idx)]
    lo, hi = np.min(yall, axis=0), np.max(yall, axis=0)
    span = np.where(hi - lo == 0, 1.0, hi - lo)
    FFN = (yall - lo) / span
    return (np.max(WG * FFN, axis=1) + 0.05 * np.sum(WG * FFN, axis=1)).reshape(-1, 1)
