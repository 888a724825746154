"""The SOBO acquisition sweep with its candidate set sharded over the GPUs of one box (BASELINE configs[1],
SURVEY.md section 8e): ordinary Kriging n=1000, d=10 at fixed hyper-parameters, EI over `ncand` device-generated
Sobol candidates through `run_single_opt(acquifuncopt='sweep', shard_candidates=True)`.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/sweep_multi.py [log10 ncand]

Rank 0 also runs the unsharded sweep over the same stretch of the Sobol sequence: winner, coordinates and EI must be
identical.  One JSON line from rank 0 (time = max over ranks, device-generated candidates, no polish).
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200.optim_tools import acquifunc_opt
from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
from bayesianopttools_b200.surrogate_models.supports.likelihood_func import likelihood


def main():
    order = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    rank, size = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    n, d = 1000, 10
    rng = np.random.default_rng(1)
    X = rng.uniform(-5, 5, (n, d))
    y = 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1).reshape(-1, 1)
    K = initkriginfo("single")
    K.update(X=X, y=y, lb=-5 * np.ones(d), ub=5 * np.ones(d), kernel=["gaussian"], nugget=-6, TrendOrder=0,
             optimizer="lbfgsb", nrestart=1)
    krig = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    krig.KrigInfo = likelihood(np.full(d, 0.2), krig.KrigInfo, mode="all", trainvar=False)
    info = dict(acquifuncopt="sweep", acquifunc="EI", nrestart=1, ncandidates=10 ** order, sweep_polish=False)

    def sweep(shard):
        acquifunc_opt._sweep_calls = 3                  # the same stretch of the sequence every time
        torch.cuda.synchronize()
        if shard and size > 1:
            dist.barrier()
        t0 = time.perf_counter()
        x, f = acquifunc_opt.run_single_opt(krig, dict(info, shard_candidates=shard))
        torch.cuda.synchronize()
        return x, f, time.perf_counter() - t0

    sweep(size > 1)                                      # warm-up
    x, f, t = sweep(size > 1)
    tt = torch.tensor([t], dtype=torch.float64, device="cuda")
    if size > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    line = dict(workload="run_single_opt sweep: EI over 1e%d device-generated Sobol candidates, Kriging n=1000 d=10" % order,
                n_gpus=size, sweep_s=float(tt.item()), candidates_per_s=10 ** order / float(tt.item()),
                best_ei=float(f), x_next=np.round(x, 6).tolist())
    ok = True
    if size > 1:
        if rank == 0:
            x1, f1, t1 = sweep(False)
            ok = bool(np.array_equal(x, x1) and f == f1)
            line.update(identical_to_one_gpu=ok, one_gpu_sweep_s=t1)
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if size > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
