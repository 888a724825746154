"""CPU restatement of the reference's 2-objective expected hypervolume improvement
(optim_tools/ehvi/exi2d.py:6-83 with exipsi.py, gaussfcn.py, hvolume2d.py).  TEST INFRASTRUCTURE ONLY.

The cell decomposition (staircase cells, their bounds and the dominated volume S+ of each cell) does
not depend on the candidate, so it is built once per Pareto front (`cell_table`); a candidate then costs
one (Phi, phi) pair per cell boundary.  Pinned: tests/golden/ehvi2d.npz holds exi2d outputs of the
unmodified reference (oracle/make_golden.py::ehvi2d)."""
import numpy as np
from scipy.special import erf


def hvolume2d(P, x):
    """hvolume2d.py:3-13"""
    P = np.atleast_2d(P)
    if P.size == 0:
        return 0.0
    S = P[np.argsort(P[:, 0], kind="stable"), :]
    h = (x[0] - S[0, 0]) * (x[1] - S[0, 1])
    for i in range(1, len(S)):
        h += (x[0] - S[i, 0]) * (S[i - 1, 1] - S[i, 1])
    return float(h)


def cell_table(P, r):
    """exi2d.py:7-63: per cell (i+1, j+1): fmax1, fmax2, cL1, cU1, cL2, cU2, sPlus; cells with
    j >= k - ii do not exist (sPlus = nan marks them)."""
    P = np.atleast_2d(np.asarray(P, dtype=float))
    S = P[np.argsort(P[:, 0], kind="stable"), :]
    k = S.shape[0]
    c1, c2 = np.sort(S[:, 0]), np.sort(S[:, 1])
    tab = np.full((k + 1, k + 1, 7), np.nan)
    for i in range(-1, k):
        ii = 0 if i == -1 else i + 1
        for j in range(-1, k - ii):
            fmax2 = r[1] if j == -1 else c2[k - j - 1]
            fmax1 = r[0] if i == -1 else c1[k - i - 1]
            cL1 = -np.inf if j == -1 else c1[j]
            cL2 = -np.inf if i == -1 else c2[i]
            cU1 = r[0] if j == k - 1 else c1[j + 1]
            cU2 = r[1] if i == k - 1 else c2[i + 1]
            SM = S[(cU1 <= S[:, 0]) & (cU2 <= S[:, 1])]
            sp = hvolume2d(SM, (fmax1, fmax2)) if SM.size else 0.0
            tab[i + 1, j + 1] = (fmax1, fmax2, cL1, cU1, cL2, cU2, sp)
    return c1, c2, tab


def _phi(z):
    return np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)


def _Phi(z):
    return 0.5 * (1 + erf(z / np.sqrt(2)))


def exi2d(P, r, mu, s):
    """EHVI of candidates with predicted means mu (m, 2) and spreads s (m, 2) (exi2d.py:6-83)."""
    mu, s = np.atleast_2d(mu).astype(float), np.atleast_2d(s).astype(float)
    _, _, tab = cell_table(P, r)
    k1 = tab.shape[0]
    out = np.zeros(len(mu))
    for a in range(k1):
        for b in range(k1):
            fmax1, fmax2, cL1, cU1, cL2, cU2, sp = tab[a, b]
            if np.isnan(sp):
                continue

            def psi(fm, bnd, m_, s_):
                z = (bnd - m_) / s_
                return s_ * _phi(z) + (fm - m_) * _Phi(z)
            Psi1 = psi(fmax1, cU1, mu[:, 0], s[:, 0]) - (psi(fmax1, cL1, mu[:, 0], s[:, 0]) if np.isfinite(cL1) else 0.0)
            Psi2 = psi(fmax2, cU2, mu[:, 1], s[:, 1]) - (psi(fmax2, cL2, mu[:, 1], s[:, 1]) if np.isfinite(cL2) else 0.0)
            G1 = _Phi((cU1 - mu[:, 0]) / s[:, 0]) - (_Phi((cL1 - mu[:, 0]) / s[:, 0]) if np.isfinite(cL1) else 0.0)
            G2 = _Phi((cU2 - mu[:, 1]) / s[:, 1]) - (_Phi((cL2 - mu[:, 1]) / s[:, 1]) if np.isfinite(cL2) else 0.0)
            out += np.maximum(Psi1 * Psi2 - sp * G1 * G2, 0.0)
    return out
