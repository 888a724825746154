"""Randomised parity sweep (GPU vs the numpy oracle) over shapes, kernels, trend orders, nugget / variance
options and population sizes -- a development aid that exercises every task kind of the Cholesky (thin and
full right-hand sides, paired and single tiles, one to many chain workers) and the three predictor variants.
   python tools/fuzz_parity.py [ncases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib
from oracle import kriging_oracle as ko

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
KERNELS = ["gaussian", "exponential", "matern32", "matern52"]
worst = dict(nll=0.0, pred=0.0, ssqr=0.0)
fails = 0
for case in range(ncases):
    n = int(rng.choice([rng.integers(2, 70), rng.integers(60, 140), rng.integers(120, 420), rng.integers(400, 900)]))
    d = int(rng.integers(1, 9))
    order = int(rng.choice([0, 0, 0, 1, 2]))
    kern = [str(rng.choice(KERNELS))]
    if rng.random() < 0.2:
        kern = ["gaussian", str(rng.choice(KERNELS[1:]))]
    trainvar = bool(rng.random() < 0.25)
    tuned_nugget = bool(rng.random() < 0.25)
    B = int(rng.choice([1, 2, 5, 17, 40, 70]))
    X = rng.uniform(-1, 1, (n, d))
    y = np.sum(np.sin(2 * X) + 0.3 * X ** 2, axis=1) + 0.05 * rng.standard_normal(n)
    idx = ko.total_order_index(order, d)
    if idx.shape[0] >= n - 1:
        idx = ko.total_order_index(0, d)
    F = ko.legendre_basis(idx, X)
    cols = [rng.uniform(-0.6, 0.8, (B, d))]
    if trainvar:
        cols.append(rng.uniform(-1, 1, (B, 1)))
    if tuned_nugget:
        cols.append(rng.uniform(-8, -3, (B, 1)))
    if len(kern) > 1:
        cols.append(rng.uniform(0.1, 1.0, (B, len(kern))))
    xpop = np.hstack(cols)
    kw = dict(kernels=tuple(kern), trainvar=trainvar, fixed_nugget=-6.0)
    tag = f"case {case}: n={n} d={d} nind={F.shape[1]} kern={kern} trainvar={trainvar} tuned_nugget={tuned_nugget} B={B}"
    try:
        mdl = _lib.DeviceModel(X, y, F, kern)
        nll, info = mdl.nll_batch(xpop, trainvar, -6.0, return_info=True)
        ref, rinfo = ko.nll_batch(xpop, X, y, F, **kw)
        ok = (info == 0) & (rinfo == 0)
        e_nll = float(np.max(np.abs(nll[ok] - ref[ok]) / np.abs(ref[ok]))) if ok.any() else 0.0
        bad_flags = int(np.sum((info != 0) != (rinfo != 0)))
        b = int(np.flatnonzero(ok)[0]) if ok.any() else None
        e_pred = e_ssqr = 0.0
        if b is not None:
            st = mdl.fit_state(xpop[b], trainvar, -6.0, want_U=False, want_Psi=False)
            fit = ko.nll(xpop[b], X, y, F, full=True, **kw)
            mdl.set_trend(idx)
            mdl.set_norm(_lib.NORM_NONE)
            m = int(rng.integers(1, 700))
            xq = rng.uniform(-1, 1, (m, d))
            outs, res = mdl.predict(xq, ("pred", "ssqr"), acq="pred")
            rp = ko.predict(xq, X, y, F, idx, fit, kernels=tuple(kern))
            sc = max(np.max(np.abs(rp["pred"])), 1e-300)
            e_pred = float(np.max(np.abs(outs["pred"] - rp["pred"])) / sc)
            e_ssqr = float(np.max(np.abs(outs["ssqr"] - rp["ssqr"])) / max(np.max(np.abs(rp["ssqr"])), 1e-300))
            if res["best_idx"] != int(np.argmin(outs["pred"])):
                bad_flags += 1
        mdl.close()
        worst["nll"] = max(worst["nll"], e_nll); worst["pred"] = max(worst["pred"], e_pred); worst["ssqr"] = max(worst["ssqr"], e_ssqr)
        # 1e-9 is the parity bar on well-conditioned designs; candidates with cond(R) ~ 1e10 (nugget 1e-8, very
        # smooth kernels) sit where LAPACK itself is 1e-9..1e-8 off a long-double evaluation
        # (profiles/r01_accuracy_vs_longdouble.txt), so the sweep flags only differences beyond 1e-7
        status = "ok" if (e_nll < 1e-7 and e_pred < 2e-8 and e_ssqr < 1e-6 and bad_flags == 0) else "FAIL"
        if status == "ok" and e_nll >= 1e-9:
            status = "ok(ill-conditioned)"
        fails += status == "FAIL"
        illc = globals().get("illc", 0) + (status == "ok(ill-conditioned)")
        globals()["illc"] = illc
        print(f"{status} {tag}: nll {e_nll:.1e} pred {e_pred:.1e} ssqr {e_ssqr:.1e} flags {bad_flags} notpd {int(np.sum(info != 0))}", flush=True)
    except Exception as ex:                      # noqa: BLE001
        fails += 1
        print(f"EXC {tag}: {type(ex).__name__}: {ex}", flush=True)
print("worst:", worst, "ill-conditioned (1e-9 <= nll err < 1e-7):", globals().get("illc", 0), "failures:", fails)
sys.exit(1 if fails else 0)
