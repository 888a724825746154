"""`Kriging.train` with every optimiser population (L-BFGS-B stencil of nbhyp+1 candidates) sharded over the GPUs
of one box: candidate b is evaluated on rank b mod G, the NLL slices are all-gathered over NCCL
(KrigInfo["shard_population"]).  BASELINE configs[2] shape by default: Styblinski-Tang d=20,
n=4000, TrendOrder 2 (231 regressors).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/train_multi.py [n] [d] [trend order] [restarts]

Every rank runs the same deterministic optimiser, so all ranks end with the same hyper-parameters; rank 0 prints one
JSON line (wall time = max over ranks).  Without torchrun: one rank, nothing sharded -- the line to compare with.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo


def train(n, d, order, nrestart, shard):
    rng = np.random.default_rng(1)
    X = rng.uniform(-5, 5, (n, d))
    y = 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1).reshape(-1, 1)
    K = initkriginfo("single")
    K.update(X=X, y=y, lb=-5 * np.ones(d), ub=5 * np.ones(d), kernel=["gaussian"], nugget=-6, TrendOrder=order,
             optimizer="lbfgsb", nrestart=nrestart, shard_population=shard)
    o = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    t0 = time.perf_counter()
    o.train(disp=False)
    torch.cuda.synchronize()
    return o, time.perf_counter() - t0


def main():
    a = [int(v) for v in sys.argv[1:]]
    n, d, order, nrestart = (a + [4000, 20, 2, 2][len(a):])[:4]
    rank, size = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    train(200, 4, 1, 1, size > 1)                      # warm-up: context, library, NCCL communicator
    o, t = train(n, d, order, nrestart, size > 1)
    theta = np.asarray(o.KrigInfo["Theta"], dtype=float).ravel()
    nll = float(np.ravel(o.KrigInfo["NegLnLike"])[0])
    same = True
    tmax = t
    if size > 1:
        buf = torch.tensor(np.concatenate([theta, [nll]]), device="cuda")
        outs = [torch.empty_like(buf) for _ in range(size)]
        dist.all_gather(outs, buf)
        same = all(torch.equal(outs[0], v) for v in outs)
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tmax = float(tt.item())
    if rank == 0:
        print(json.dumps(dict(workload="Kriging.train lbfgsb, n=%d d=%d TrendOrder=%d (nind=%d), %d restarts, stencil of %d "
                                       "candidates per call" % (n, d, order, o.KrigInfo["F"].shape[1], nrestart, d + 1),
                              n_gpus=size, train_s=tmax, NegLnLike=nll, theta_log10=theta.round(6).tolist(),
                              all_ranks_identical=bool(same))), flush=True)
    if size > 1:
        dist.destroy_process_group()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
