"""Analytic test functions used as synthetic data generators (reference:
testcase/analyticalfcn/cases.py).  Formulas as the reference defines them (its Branin is
non-standard, cases.py:42-47)."""
import numpy as np


def styb(X):
    X = np.atleast_2d(X)
    return 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1, keepdims=True)


def branin(X):
    X = np.atleast_2d(X)
    a, b, c = 5.1 / (4 * np.pi ** 2), 5 / np.pi, 1 - 1 / (8 * np.pi)
    return ((X[:, 1] - a * X[:, 1] + b * X[:, 0] - 6) ** 2 + 10 * (c * np.cos(X[:, 0]) + 1)).reshape(-1, 1)


def evaluate(X, type):
    X = np.atleast_2d(np.asarray(X, dtype=float))
    t = type.lower()
    if t == "styblinski":
        return styb(X)
    if t == "branin":
        return branin(X)
    raise NameError("Test case unavailable!")
