"""Device-side throughput of the hot path at the shapes of BASELINE.json's configs (synthetic data,
SURVEY.md section 8d), one JSON line per case with the roofline that bounds it.  Complements bench.py
(which is configs[1] and carries the e2e / CPU arms); run on the GPU box:  python tools/bench_configs.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib


def styb(X5):
    return 0.5 * np.sum(X5 ** 4 - 16 * X5 ** 2 + 5 * X5, axis=1)


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def nll_case(name, n, d, B, F=None, kpls=None, reps=5):
    rng = np.random.default_rng(1)
    X = rng.uniform(-1, 1, (n, d))
    y = styb(5 * X[:, :10]) + (0.1 * np.sum((5 * X[:, 10:]) ** 2, axis=1) if d > 10 else 0.0)
    if F is None:
        F = np.ones((n, 1))
    kw = {}
    nth = d
    if kpls:
        kw = dict(model_type="kpls", plscoeff=rng.normal(size=(d, kpls)) / np.sqrt(d))
        nth = kpls
    mdl = _lib.DeviceModel(X, y, F, ["gaussian"], **kw)
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    mdl.reserve_population(B)
    mdl.set_profiling(True)
    xpop = torch.tensor(np.random.default_rng(2).uniform(-0.3, 0.7, (B, nth)), device="cuda")
    nll = torch.empty(B, dtype=torch.float64, device="cuda")
    ms = timed(lambda: mdl.nll_batch_dev(xpop.data_ptr(), B, nth, False, -6.0, nll.data_ptr()), reps)
    st = mdl.stage_ms(0)
    flop = B * n ** 3 / 3.0
    print(json.dumps({"case": name, "n": n, "d": d, "population": B, "nind": int(F.shape[1]),
                      "ms_per_population": ms, "evals_per_s": B / ms * 1e3,
                      "stage_ms": {"decode": st[0], "corr_build": st[1], "cholesky": st[2], "gls": st[3]},
                      "cholesky_tflops_algorithmic": flop / (st[2] * 1e-3) / 1e12,
                      "cholesky_frac_of_dmma_peak": flop / (st[2] * 1e-3) / 1e12 / PEAK}), flush=True)
    mdl.close()


def sweep_case(name, n, d, m, reps=3, gaussian_pop=False):
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, (n, d))
    y = styb(5 * X[:, :10]) if d >= 10 else np.sum(X ** 2, axis=1) - 0.5
    mdl = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"])
    mdl.fit_state(np.full(d, 0.3 if d >= 10 else -0.2), False, -6.0, want_U=False, want_Psi=False)
    mdl.set_norm(_lib.NORM_NONE)
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    if gaussian_pop:
        xq = torch.randn(m, d, dtype=torch.float64, device="cuda") * 0.4
    else:
        xq = torch.rand(m, d, dtype=torch.float64, device="cuda") * 2 - 1
    res = torch.empty(8, dtype=torch.float64, device="cuda")
    acq = "u" if gaussian_pop else "ei"
    ms = timed(lambda: mdl.predict_dev(xq.data_ptr(), m, acq=acq, ybest=float(y.min()), result_ptr=res.data_ptr()),
               reps, warm=1)
    flop = float(m) * n * n
    print(json.dumps({"case": name, "n": n, "d": d, "sites": m, "ms": ms, "sites_per_s": m / ms * 1e3,
                      "tflops_algorithmic": flop / (ms * 1e-3) / 1e12, "frac_of_dmma_peak": flop / (ms * 1e-3) / 1e12 / PEAK,
                      "input_GBps": m * d * 8 / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": m * d * 8 / (ms * 1e-3) / 1e9 / HBM}),
          flush=True)
    mdl.close()


if __name__ == "__main__":
    PEAK = _lib.measure_peak(0)
    HBM = 6554.2
    try:
        HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps({"dmma_peak_tflops": PEAK, "hbm_peak_GBps": HBM}), flush=True)
    nll_case("C2 likelihood n=1000 d=10 B=64", 1000, 10, 64)
    nll_case("C2 per-GPU share of an 8-GPU population (B=8)", 1000, 10, 8)
    nll_case("north_star Cholesky target n=2000 B=64", 2000, 10, 64)
    from oracle_free_trend import trend_matrix
    nll_case("C3 trend n=4000 d=20 order 2 (nind=231) B=8", 4000, 20, 8, F=trend_matrix(4000, 20, 2), reps=3)
    nll_case("C4 KPLS n=2000 d=100 h=3 B=64", 2000, 100, 64, kpls=3)
    sweep_case("C2 EI sweep n=1000 d=10, 4e6 sites", 1000, 10, 4_000_000)
    sweep_case("sweep n=2000 d=10, 1e6 sites", 2000, 10, 1_000_000)
    sweep_case("C5 AK-MCS U sweep n=40 d=2, 1e8 sites", 40, 2, 100_000_000, gaussian_pop=True)
    sweep_case("C5 AK-MCS late model n=120 d=2, 2e7 sites", 120, 2, 20_000_000, gaussian_pop=True)
