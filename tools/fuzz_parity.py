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
from oracle import longdouble_truth as lt

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
KERNELS = ["gaussian", "exponential", "matern32", "matern52"]
worst = dict(nll=0.0, pred=0.0, ssqr=0.0)
fails = narb = 0
arb = dict(nll=[], pred=[], ssqr=[])
failq = dict(nll=0, pred=0, ssqr=0, flags=0)
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
        # 1e-9 is the parity bar.  Where the two FP64 paths differ by more (cond(R) ~ 1e10: nugget tuned down to 1e-8
        # on smooth kernels), an 80-bit evaluation of the same formulas arbitrates: the GPU value passes when it is
        # no further from it than 3x the LAPACK oracle's own distance (both are random rounding walks) or 1e-9.
        status, extra = "ok", ""
        if bad_flags:
            status = "FAIL"
            failq["flags"] += 1
        if e_nll >= 1e-9 and ok.any():
            rel = np.where(ok, np.abs(nll - ref) / np.abs(ref), 0.0)
            bw = int(np.argmax(rel))
            t = lt.nll(xpop[bw], X, y, F, kernels=tuple(kern), trainvar=trainvar, fixed_nugget=-6.0)
            eg, eo = abs(nll[bw] - t) / abs(t), abs(ref[bw] - t) / abs(t)
            extra += f" | nll vs 80-bit: gpu {eg:.1e} lapack {eo:.1e}"
            arb["nll"].append((eg, eo))
            failq["nll"] += eg > max(1e-9, 3 * eo)
            status = "FAIL" if eg > max(1e-9, 3 * eo) else ("ok(arbitrated)" if status == "ok" else status)
        if (e_pred >= 1e-9 or e_ssqr >= 1e-9) and b is not None:
            tf = lt.nll(xpop[b], X, y, F, kernels=tuple(kern), trainvar=trainvar, fixed_nugget=-6.0, full=True)
            tp = lt.predict(xq, X, y, F, idx, tf, kernels=tuple(kern))
            for k, e in (("pred", e_pred), ("ssqr", e_ssqr)):
                sc = max(np.max(np.abs(tp[k])), 1e-300)
                eg, eo = np.max(np.abs(outs[k] - tp[k])) / sc, np.max(np.abs(rp[k] - tp[k])) / sc
                extra += f" | {k} vs 80-bit: gpu {eg:.1e} lapack {eo:.1e}"
                arb[k].append((eg, eo))
                if e >= 1e-9:
                    failq[k] += eg > max(1e-9, 3 * eo)
                    status = "FAIL" if eg > max(1e-9, 3 * eo) else ("ok(arbitrated)" if status == "ok" else status)
        fails += status == "FAIL"
        narb += status == "ok(arbitrated)"
        print(f"{status} {tag}: nll {e_nll:.1e} pred {e_pred:.1e} ssqr {e_ssqr:.1e} flags {bad_flags} notpd {int(np.sum(info != 0))}{extra}", flush=True)
    except Exception as ex:                      # noqa: BLE001
        fails += 1
        print(f"EXC {tag}: {type(ex).__name__}: {ex}", flush=True)
print("worst GPU-vs-LAPACK difference:", {k: f"{v:.1e}" for k, v in worst.items()}, "| cases beyond 1e-9 settled by the 80-bit evaluation:", narb,
      "| failures:", fails, "| by quantity (further from the 80-bit value than max(1e-9, 3 x LAPACK's distance)):",
      {k: int(v) for k, v in failq.items()})
for k, v in arb.items():
    if v:
        a = np.array(v)
        print(f"  {k}: {len(a)} arbitrated values; worst distance from the 80-bit value: gpu {a[:, 0].max():.1e}, lapack {a[:, 1].max():.1e}; "
              f"gpu closer than lapack in {int(np.sum(a[:, 0] <= a[:, 1]))}")
sys.exit(1 if fails else 0)
