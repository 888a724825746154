import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture
def golden():
    return load_golden


def relerr(a, b):
    """Scale-relative error  max|a-b| / max|b|  (tolerance note)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    scale = max(float(np.max(np.abs(b))), np.finfo(float).tiny)
    return float(np.max(np.abs(a - b))) / scale


def assert_rel(label, a, b, tol=1e-9):
    """Scale-relative comparison that PRINTS the achieved error (pytest -s / the captured output of a failure)."""
    e = relerr(a, b)
    print(f"[achieved] {label}: {e:.1e} (bar {tol:.0e})")
    assert e < tol, (label, e)
