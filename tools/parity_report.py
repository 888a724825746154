"""Achieved parity errors, printed (the tests only assert): GPU path vs every reference-generated fixture in
tests/golden and vs the oracle at the BASELINE shapes.   python tools/parity_report.py [--big]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bayesianopttools_b200 import _lib
from conftest import load_golden, relerr
from oracle import kriging_oracle as ko

RANDOM_CASES = ["ok_gaussian", "ok_exponential", "ok_matern32", "ok_matern52", "ok_gaussian_trainvar",
                "ok_gaussian_tunednugget", "ok_composite", "ok_composite_tn_tv", "trend1", "trend2", "kpls",
                "ok_gaussian_n200_d10"]


def case_kwargs(g):
    nug = g["nugget"]
    kw = dict(kernels=[str(k) for k in g["kernel"]], trainvar=bool(g["trainvar"]),
              fixed_nugget=float(nug) if nug.ndim == 0 else 0.0)
    if "plscoeff" in g:
        kw.update(plscoeff=g["plscoeff"], n_princomp=int(g["n_princomp"]))
    return kw


def pointwise(a, b):
    """(scale-relative max error, index of the worst point, second-worst error)"""
    a, b = np.ravel(a), np.ravel(b)
    e = np.abs(a - b) / max(float(np.max(np.abs(b))), 1e-300)
    o = np.argsort(e)
    return float(e[o[-1]]), int(o[-1]), float(e[o[-2]]) if len(o) > 1 else 0.0


worst = {}


def note(k, v):
    worst[k] = max(worst.get(k, 0.0), v)


for name in RANDOM_CASES:
    g = load_golden(name)
    kw = case_kwargs(g)
    mdl = _lib.DeviceModel(g["X_norm"], g["y"], g["F"], kw["kernels"], model_type="kpls" if "plscoeff" in kw else "kriging",
                           plscoeff=kw.get("plscoeff"))
    nll, info = mdl.nll_batch(g["xpop"], kw["trainvar"], kw["fixed_nugget"], return_info=True)
    beta, s2 = mdl.nll_extras(len(nll))
    e_nll = float(np.max(np.abs(nll - g["nll"]) / np.abs(g["nll"])))
    line = f"{name:26s} nll {e_nll:.1e} beta {relerr(beta, g['BE']):.1e} sigma2 {relerr(s2, g['sigma2']):.1e}"
    note("nll", e_nll); note("beta", relerr(beta, g["BE"])); note("sigma2", relerr(s2, g["sigma2"]))
    if "p_pred" in g:
        mdl.fit_state(g["xpop"][-1], kw["trainvar"], kw["fixed_nugget"], want_U=False, want_Psi=False)
        mdl.set_norm(_lib.NORM_DEFAULT, g["lb"], g["ub"])
        names = ("pred", "ssqr", "s", "ei", "poi", "pof")
        outs, _ = mdl.predict(g["xq"], names, acq="pred", ybest=float(np.min(g["y"])), limit=float(g["limit"]), limit_sign=1)
        for k in names:
            e, i, e2 = pointwise(outs[k], g["p_" + k])
            line += f" | {k} {e:.1e}@{i} ({e2:.1e})"
            note(k, e); note(k + "_2nd", e2)
    print(line, flush=True)
    mdl.close()

for name in ("krigdemo_nb", "sobo_nb"):
    g = load_golden(name)
    mdl = _lib.DeviceModel(g["X_norm"], g["y"], g["F"], ["gaussian"])
    st = mdl.fit_state(g["theta_log10"], False, -6.0)
    mdl.set_norm(_lib.NORM_DEFAULT, g["lb"], g["ub"])
    outs, _ = mdl.predict(g["xq"], ("pred", "ssqr", "s", "ei", "poi"), acq="ei", ybest=float(np.min(g["y"])))
    line = f"{name:26s} nll {abs(st['NegLnLike'] - float(g['nll'])) / abs(float(g['nll'])):.1e} U {relerr(st['U'], g['U']):.1e}"
    for k in ("pred", "ssqr", "s", "ei", "poi"):
        e, i, e2 = pointwise(outs[k], g["p_" + k])
        line += f" | {k} {e:.1e}@{i} ({e2:.1e})"
        note(k, e)
    lo = mdl.loocv()
    line += f" | loocv {relerr(lo, g['loocv_pred']):.1e}"
    note("loocv", relerr(lo, g["loocv_pred"]))
    print(line, flush=True)
    mdl.close()

print("worst over the fixtures:", {k: f"{v:.1e}" for k, v in worst.items()}, flush=True)

if "--big" in sys.argv:
    # BASELINE configs[1] on the bench's own population, all 64 candidates
    rng = np.random.default_rng(1)
    n, d = 1000, 10
    X = rng.uniform(-5, 5, (n, d))
    y = 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1)
    Xn = ((X + 5.0) / 10.0 - 0.5) * 2
    F = np.ones((n, 1))
    xpop = np.random.default_rng(2).uniform(-1.0, 1.0, (64, d))
    mdl = _lib.DeviceModel(Xn, y, F, ["gaussian"])
    nll, info = mdl.nll_batch(xpop, False, -6.0, return_info=True)
    errs = []
    for b in range(64):
        try:
            ref = ko.nll(xpop[b], Xn, y, F)
            errs.append(abs(nll[b] - ref) / abs(ref) if info[b] == 0 else np.inf)
        except np.linalg.LinAlgError:
            errs.append(0.0 if info[b] != 0 else np.inf)
    errs = np.array(errs)
    print(f"C2 bench population (64 x U(-1,1)): worst nll rel err {errs.max():.2e}, median {np.median(errs):.1e}, "
          f"> 1e-9: {int(np.sum(errs > 1e-9))}, not PD {int(np.sum(info != 0))}", flush=True)
    b = int(np.argmin(np.where(info == 0, nll, np.inf)))
    mdl.fit_state(xpop[b], False, -6.0, want_U=False, want_Psi=False)
    fit = ko.nll(xpop[b], Xn, y, F, full=True)
    mdl.set_norm(_lib.NORM_NONE)
    xq = rng.uniform(-1, 1, (3000, d))
    outs, _ = mdl.predict(xq, ("pred", "ssqr", "ei"), acq="ei", ybest=float(y.min()))
    ref = ko.predict(xq, Xn, y, F, np.zeros((1, d), int), fit)
    print("C2 predictor at the best candidate (3000 sites): " +
          " ".join(f"{k} {relerr(outs[k], ref[k]):.1e}" for k in ("pred", "ssqr", "ei")), flush=True)
    mdl.close()
