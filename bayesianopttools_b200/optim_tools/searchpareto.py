"""Non-dominated filtering (reference: optim_tools/searchpareto.py:5-50 `paretopoint`).

Row k of B is a Pareto point when every OTHER row i has at least one objective strictly worse than
row k (B[k, j] - B[i, j] < 0 for some j), the reference's criterion -- so a row that duplicates another
row is not kept.  Evaluated as one (m, m, nobj) comparison instead of the reference's double loop."""
import warnings

import numpy as np


def paretopoint(B):
    B = np.asarray(B, dtype=float)
    m = B.shape[0]
    better_somewhere = np.any(B[:, None, :] - B[None, :, :] < 0, axis=2)      # [k, i]: k beats i in some objective
    np.fill_diagonal(better_somewhere, True)
    keep = np.all(better_somewhere, axis=1)
    if not keep.any():
        warnings.warn("There are no Pareto points. The result is an empty matrix.")
        return [], []
    return B[keep].copy(), np.flatnonzero(keep).astype(float)
