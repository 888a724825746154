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


def assert_rel_or_arbiter(label, got, oracle, truth_fn, tol=1e-9, arbiter_tol=1e-8):
    """Parity with the LAPACK oracle at `tol`; where the two FP64 paths differ by more (an ill-conditioned design:
    two correct implementations then differ by about eps x cond(R)), the 80-bit evaluation `truth_fn()` arbitrates
    (oracle/longdouble_truth.py): the result must be within `arbiter_tol` of it, or no further than 3x the oracle's
    own distance.  (numpy emulation of LAPACK, a column-by-column trsm and this path's blocking on the designs of
    test_seated_right_hand_sides_at_their_limits -- cond(R) ~ 2e8 -- puts ALL of them 1e-10 ... 6e-9 from the
    80-bit beta; which one is closest changes from candidate to candidate.)"""
    e = relerr(got, oracle)
    if e < tol:
        print(f"[achieved] {label}: {e:.1e} (bar {tol:.0e})")
        return
    truth = np.asarray(truth_fn(), dtype=np.float64)
    eg, eo = relerr(got, truth), relerr(oracle, truth)
    print(f"[achieved] {label}: {e:.1e} from the oracle; from the 80-bit value: this path {eg:.1e}, oracle {eo:.1e} "
          f"(bar {arbiter_tol:.0e})")
    assert eg <= max(arbiter_tol, 3.0 * eo), (label, e, eg, eo)
