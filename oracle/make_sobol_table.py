"""Extract the Sobol direction-number matrix (ACM TOMS Algorithm 659, Bratley & Fox 1988,
40 dimensions x 30 bits) that the reference's start-point sampler uses
(misc/sampling/sobol_seq.py:137-381) by initialising the reference generator and reading
its state.  Output: bayesianopttools_b200/misc/sampling/sobol_dirs_toms659.npy (int64).
TEST/DATA INFRASTRUCTURE -- run in the build container only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
import misc.sampling.sobol_seq as ss  # noqa: E402  (reference module)

ss.i4_sobol(40, 0)
V = np.array(ss.v[:40, :ss.maxcol], dtype=np.int64)
assert ss.maxcol == 30 and abs(ss.recipd - 2.0 ** -30) < 1e-30
out = os.path.join(ROOT, "bayesianopttools_b200", "misc", "sampling", "sobol_dirs_toms659.npy")
np.save(out, V)
print("wrote", out, V.shape)
