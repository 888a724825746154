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


def schaffer(X):
    """Generalised Schaffer problem with r = 1 (cases.py:92-118): two objectives."""
    X = np.atleast_2d(X)
    n = X.shape[1]
    f = np.column_stack((np.sqrt(np.sum(X ** 2, axis=1)), np.sqrt(np.sum((1 - X) ** 2, axis=1)))) / np.sqrt(n)
    return f[0] if len(f) == 1 else f


def fonseca(X):
    """Fonseca-Fleming with the reference's fixed n = 2 centre 1/sqrt(2) (cases.py:120-136)."""
    X = np.atleast_2d(X)
    c = 1.0 / np.sqrt(2.0)
    f = np.column_stack((1 - np.exp(-np.sum((X - c) ** 2, axis=1)), 1 - np.exp(-np.sum((X + c) ** 2, axis=1))))
    return f[0] if len(f) == 1 else f


def schaffer1(X):
    """cases.py:138-149."""
    X = np.atleast_2d(X)
    f = np.column_stack((X[:, 0] ** 2, (X[:, 0] - 2) ** 2))
    return f[0] if len(f) == 1 else f


def evaluate(X, type):
    X = np.atleast_2d(np.asarray(X, dtype=float))
    t = type.lower()
    if t == "styblinski":
        return styb(X)
    if t == "branin":
        return branin(X)
    if t == "schaffer":
        return schaffer(X)
    if t == "schaffer1":
        return schaffer1(X)
    if t == "fonseca":
        return fonseca(X)
    raise NameError("Test case unavailable!")
