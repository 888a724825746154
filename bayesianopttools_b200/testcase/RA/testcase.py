"""Reliability-analysis limit states used as synthetic generators (reference: testcase/RA/testcase.py)."""
import numpy as np


def fourbranches(X, k=7):
    X = np.atleast_2d(X)
    x1, x2 = X[:, 0], X[:, 1]
    y = np.stack([3 + 0.1 * (x1 - x2) ** 2 - (x1 + x2) / np.sqrt(2),
                  3 + 0.1 * (x1 - x2) ** 2 + (x1 + x2) / np.sqrt(2),
                  (x1 - x2) + k / np.sqrt(2),
                  (x2 - x1) + k / np.sqrt(2)], axis=1)
    return np.min(y, axis=1, keepdims=True)


def hidimenRA(X):
    X = np.atleast_2d(X)
    n = X.shape[1]
    return (n + 3 * 0.2 * np.sqrt(n) - np.sum(X, axis=1)).reshape(-1, 1)


def evaluate(X, type="fourbranches"):
    X = np.atleast_2d(np.asarray(X, dtype=float))
    t = type.lower()
    if t == "fourbranches":
        return fourbranches(X)
    if t == "hidimenra":
        return hidimenRA(X)
    if t == "styblinski":
        return 50 - 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1, keepdims=True)
    raise NameError("Test case unavailable!")
