"""`exi2d` drop-in (reference optim_tools/ehvi/exi2d.py:6-83): expected hypervolume improvement of a
candidate with independent Gaussian objectives N(mu_i, s_i) over the 2-objective Pareto front P with
reference point r.  The staircase-cell table is built once per front and the cells are summed by the
CUDA kernel (csrc/kb_ehvi.cu); `mu`/`s` of shape (m, 2) evaluate m candidates in one launch."""
import numpy as np

from ... import _lib


def exi2d(P, r, mu, s):
    mu = np.asarray(mu, dtype=float)
    out = _lib.exi2d_batch(P, r, mu, s)
    return float(out[0]) if mu.ndim == 1 else out
