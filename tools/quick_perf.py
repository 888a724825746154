"""Quick device-side timings of the three headline kernels (development aid, not the bench)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib


def styb(X):
    return 0.5 * np.sum((5 * X) ** 4 - 16 * (5 * X) ** 2 + 5 * (5 * X), axis=1)


def time_it(fn, reps=5, warm=2):
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


def nll_case(n, d, B, nind=1):
    rng = np.random.default_rng(1)
    X = rng.uniform(-1, 1, (n, d))
    y = styb(X[:, :10])
    F = np.ones((n, 1)) if nind == 1 else np.hstack([np.ones((n, 1)), rng.normal(size=(n, nind - 1))])
    mdl = _lib.DeviceModel(X, y, F, ["gaussian"])
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    mdl.reserve_population(B)
    mdl.set_profiling(True)
    xpop = torch.tensor(np.random.default_rng(2).uniform(-0.3, 0.7, (B, d)), device="cuda")
    nll = torch.empty(B, dtype=torch.float64, device="cuda")
    ms = time_it(lambda: mdl.nll_batch_dev(xpop.data_ptr(), B, d, False, -6.0, nll.data_ptr()))
    npad = (n + 63) // 64 * 64
    flops = B * (npad ** 3 / 3.0)
    print(f"NLL n={n} d={d} B={B} nind={nind}: {ms:.3f} ms/pop  {B / ms * 1e3:.1f} evals/s  "
          f"chol {flops / ms / 1e9:.2f} TFLOP/s (n_pad^3/3)  stages(decode,build,chol,gls) ms "
          + " ".join("%.3f" % v for v in mdl.stage_ms(0)[:4]), flush=True)
    mdl.close()


def kpls_case(n, d, h, B):
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, (n, d))
    y = styb(X[:, :10]) + 0.1 * np.sum((5 * X[:, 10:]) ** 2, axis=1)
    coef = rng.normal(size=(d, h)) / np.sqrt(d)
    mdl = _lib.DeviceModel(X, y, np.ones((n, 1)), ["gaussian"], model_type="kpls", plscoeff=coef)
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    mdl.reserve_population(B)
    mdl.set_profiling(True)
    xpop = torch.tensor(np.random.default_rng(2).uniform(-0.3, 0.7, (B, h)), device="cuda")
    nll = torch.empty(B, dtype=torch.float64, device="cuda")
    ms = time_it(lambda: mdl.nll_batch_dev(xpop.data_ptr(), B, h, False, -6.0, nll.data_ptr()))
    print(f"KPLS NLL n={n} d={d} h={h} B={B}: {ms:.3f} ms/pop  {B / ms * 1e3:.1f} evals/s  stages(decode,build,chol,gls) ms "
          + " ".join("%.3f" % v for v in mdl.stage_ms(0)[:4]), flush=True)
    mdl.close()


def predict_case(n, d, m):
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, (n, d))
    y = styb(X[:, :10]) if d >= 2 else X[:, 0]
    F = np.ones((n, 1))
    mdl = _lib.DeviceModel(X, y, F, ["gaussian"])
    mdl.fit_state(np.full(d, 0.3), False, -6.0, want_U=False, want_Psi=False)
    mdl.set_norm(_lib.NORM_NONE)
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    xq = torch.rand(m, d, dtype=torch.float64, device="cuda") * 2 - 1
    res = torch.empty(8, dtype=torch.float64, device="cuda")
    ms = time_it(lambda: mdl.predict_dev(xq.data_ptr(), m, acq="ei", ybest=float(y.min()),
                                         result_ptr=res.data_ptr()), reps=3, warm=1)
    npad = (n + 63) // 64 * 64
    print(f"PREDICT n={n} d={d} m={m}: {ms:.2f} ms  {m / ms * 1e3:.3e} pts/s  "
          f"{m * npad * npad / ms / 1e9:.2f} TFLOP/s (n_pad^2 flop/pt)  {m * d * 8 / ms / 1e6:.1f} GB/s in",
          flush=True)
    mdl.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "nll2000":
        nll_case(2000, 10, 64)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "kpls":
        kpls_case(2000, 100, 3, 64)
        sys.exit(0)
    if len(sys.argv) > 3 and sys.argv[1] == "pred":          # pred <n> <d> [log2 sites]
        predict_case(int(sys.argv[2]), int(sys.argv[3]), 1 << (int(sys.argv[4]) if len(sys.argv) > 4 else 19))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c3":
        nll_case(4000, 20, 8, nind=231)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pred1000":
        predict_case(1000, 10, 1 << 20)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "nll1000b8":
        nll_case(1000, 10, 8)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "nll1000":
        nll_case(1000, 10, 64)
        sys.exit(0)
    nll_case(1000, 10, 64)
    nll_case(1000, 10, 8)
    nll_case(2000, 10, 64)
    nll_case(2000, 10, 16)
    nll_case(4000, 20, 8, nind=231)
    predict_case(1000, 10, 1 << 20)
    predict_case(2000, 10, 1 << 19)
    predict_case(32, 2, 1 << 24)
    predict_case(40, 2, 1 << 24)
    predict_case(64, 2, 1 << 23)
    predict_case(128, 2, 1 << 23)
