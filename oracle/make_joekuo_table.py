"""Pack the first 256 dimensions of the Joe & Kuo direction-number table that the reference's
`sobolnew` sampler reads (misc/sampling/sobolcoeff.csv = new-joe-kuo-6.21201, public data from
https://web.maths.unsw.edu.au/~fkuo/sobol/) into a small integer array:
    row j-1 (dimension j >= 2):  [s, a, m_1 .. m_s, 0 ...]   (18 m-slots)
Output: bayesianopttools_b200/misc/sampling/sobol_joekuo_256.npy (int64, 256 x 20).
DATA INFRASTRUCTURE -- run in the build container only (needs /root/reference).
"""
import csv
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(os.environ.get("KADAL_REFERENCE", "/root/reference"), "misc", "sampling", "sobolcoeff.csv")
rows = []
with open(SRC) as fh:
    for rec in csv.DictReader(fh):
        m = [int(v) for v in rec["m_i"].split()]
        s = int(rec["s"])
        assert len(m) == s and s <= 18
        rows.append([s, int(rec["a"])] + m + [0] * (18 - s))
        if len(rows) == 256:
            break
out = os.path.join(ROOT, "bayesianopttools_b200", "misc", "sampling", "sobol_joekuo_256.npy")
np.save(out, np.array(rows, dtype=np.int64))
print("wrote", out, len(rows))
