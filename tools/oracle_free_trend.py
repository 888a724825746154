"""Total-order Legendre regression matrix for benchmark inputs (host numpy; mirrors
bayesianopttools_b200/surrogate_models/supports/trendfunction.py, no oracle import)."""
import numpy as np

from bayesianopttools_b200.surrogate_models.supports.trendfunction import compute_regression_mat, polytruncation


def trend_matrix(n, d, order, seed=1):
    X = np.random.default_rng(seed).uniform(-1, 1, (n, d))
    idx = polytruncation(order, d, 1)
    return compute_regression_mat(idx, X, np.vstack([-np.ones(d), np.ones(d)]), np.ones(d))
