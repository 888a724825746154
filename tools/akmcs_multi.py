"""AK-MCS with the Monte-Carlo population sharded over the GPUs of one box (BASELINE configs[4], SURVEY.md
section 8e): four-branch limit state, d = 2, N ~ N(0,1)^2 points in total, 12 Halton points to start with.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/akmcs_multi.py [log10 N] [updates] [--check]

Every rank generates its contiguous share of ONE Philox stream on its own GPU and keeps it there; per update
the ranks exchange one 64-byte record (failure counters summed, min U merged) over NCCL and the owner of the
new site broadcasts its two coordinates.  `--check`: rank 0 then repeats the run alone over the whole
population and the two histories (sites, min U, Pf, stop criterion, failure counters) must be identical.
Also runs without torchrun (one rank).
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200.misc.sampling.samplingplan import sampling
from bayesianopttools_b200.reliability_analysis.akmcs import AKMCS, mcpopgen_device
from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
from bayesianopttools_b200.testcase.RA.testcase import evaluate


def run(n_order, maxupdate, shard, barrier=True):
    np.random.seed(0)
    lb, ub = -5.0 * np.ones(2), 5.0 * np.ones(2)
    _, X = sampling("halton", 2, 12, result="real", upbound=ub, lobound=lb)
    y = evaluate(X, type="fourbranches")
    K = initkriginfo("single")
    K.update(X=X, y=y, nvar=2, problem="fourbranches", nsamp=12, nrestart=5, ub=ub, lb=lb, optimizer="lbfgsb",
             kernel=["gaussian"], nugget=-6)
    krig = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    krig.train(disp=False)
    pop = mcpopgen_device(type="gaussian", ndim=2, n_order=n_order, n_coeff=1, stddev=1, mean=0, seed=1, shard=shard)
    ak = AKMCS(krig, dict(init_samp=pop, maxupdate=maxupdate, problem="fourbranches", keep_fields=False,
                          shard_population=shard))
    counters = []
    sweep = ak._sweep

    def recording_sweep():
        sweep()
        counters.append((ak._res["n_fail"], ak._res["n_fail_minus"], ak._res["n_fail_plus"], ak._res["min_u_idx"],
                         ak._res["n_points"]))
    ak._sweep = recording_sweep
    torch.cuda.synchronize()
    if barrier and dist.is_initialized():
        dist.barrier()
    t0 = time.perf_counter()
    ak.run(disp=False)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    hist = dict(updateX=ak.updateX.tolist(), minU=np.asarray(ak.minUiter).ravel().tolist(), Pf=ak.Pf, cov=ak.cov,
                stop=ak.stop_criteria, counters=counters, nsamp=ak.nsamp, local=ak.nsamp_local)
    return hist, t


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_order = int(args[0]) if args else 8
    maxupdate = int(args[1]) if len(args) > 1 else 20
    rank = int(os.environ.get("RANK", 0))
    size = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    run(5, 2, size > 1)                                   # warm-up: library load, pools, NCCL communicator
    hist, t = run(n_order, maxupdate, size > 1)
    tmax = torch.tensor([t], dtype=torch.float64, device="cuda")
    if size > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    nupd = len(hist["updateX"]) - 1
    if rank == 0:
        print("sharded run done: %.3f s, Pf %.6g" % (float(tmax.item()), hist["Pf"]), file=sys.stderr, flush=True)
    line = dict(workload="AK-MCS four-branch d=2, population sharded over the GPUs, %d updates incl. retraining" % nupd,
                n_gpus=size, population=hist["nsamp"], population_per_gpu=hist["local"],
                loop_s=float(tmax.item()), s_per_update=float(tmax.item()) / max(1, nupd),
                sites_per_s=hist["nsamp"] * (nupd + 1) / float(tmax.item()),
                Pf=hist["Pf"], cov=hist["cov"], stop_criterion=hist["stop"],
                exchange="one 64-byte record per rank and update (all-gather), 16-byte site broadcast")
    if "--check" in sys.argv and size > 1:
        same = None
        if rank == 0:
            ref, _ = run(n_order, maxupdate, False, barrier=False)   # the whole population on this GPU alone
            same = all(hist[k] == ref[k] for k in ("updateX", "minU", "Pf", "cov", "stop", "counters", "nsamp"))
            line["identical_to_one_gpu"] = bool(same)
            if not same:
                line["one_gpu"] = {k: ref[k] for k in ("Pf", "stop", "counters")}
                line["sharded"] = {k: hist[k] for k in ("Pf", "stop", "counters")}
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if size > 1:
        dist.destroy_process_group()
    if rank == 0 and line.get("identical_to_one_gpu") is False:
        sys.exit(1)


if __name__ == "__main__":
    main()
