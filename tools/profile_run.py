"""Small fixed workloads for ncu captures: `python tools/profile_run.py nll|predict|small`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesianopttools_b200 import _lib

which = sys.argv[1] if len(sys.argv) > 1 else "nll"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
X, Xn, y, F = bench.make_design()
if which == "small":
    rng = np.random.default_rng(0)
    Xn = rng.uniform(-1, 1, (40, 2))
    y = np.sum(Xn ** 2, axis=1)
    F = np.ones((40, 1))
mdl = _lib.DeviceModel(Xn, y, F, ["gaussian"], device=0)
mdl.set_stream(torch.cuda.current_stream().cuda_stream)
d = Xn.shape[1]
if which == "nll":
    B = bench.POP
    xpop = torch.as_tensor(bench.make_population(), device="cuda")
    nll = torch.empty(B, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        mdl.nll_batch_dev(xpop.data_ptr(), B, d, False, -6.0, nll.data_ptr())
    torch.cuda.synchronize()
    print("nll ok", float(nll.min()))
else:
    mdl.fit_state(np.full(d, 0.3), False, -6.0, want_U=False, want_Psi=False)
    mdl.set_norm(_lib.NORM_NONE)
    m = 1 << 19 if which == "predict" else 1 << 22
    xq = torch.rand(m, d, dtype=torch.float64, device="cuda") * 2 - 1
    res = torch.zeros(8, dtype=torch.int64, device="cuda")
    for _ in range(reps):
        mdl.predict_dev(xq.data_ptr(), m, acq="ei", ybest=float(y.min()), result_ptr=res.data_ptr())
    torch.cuda.synchronize()
    print("predict ok", res.cpu().numpy()[:4])
mdl.close()
