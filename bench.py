#!/usr/bin/env python
"""bench.py -- BASELINE.json configs[1] on N GPUs of one node, plus the other configs and the sharded runs.

Headline workload ("SOBO d=10, n=1000 ordinary Kriging"): one STEP is one pass of the hot path over one
batch of synthetic input on every rank:
  phase A  population-batched concentrated ln-likelihood, B = 64 hyper-parameter candidates (a new population
           every step)  (decode -> correlation build -> DMMA Cholesky -> GLS reductions)
  phase B  fused predict + EI sweep with device arg-min over M = 1e7 candidate sites
`value` is BASELINE's first metric (likelihood evaluations / s) over phase A of the timed steps, device-timed
with the inputs resident in HBM; `e2e` is the same metric through the reference-facing API
(`likelihood_batch`, `Kriging.predict_sweep`) with pageable numpy inputs and numpy results.  Weak scaling for
`value`: every rank owns its own population and candidate shard.

Two more blocks ride in the same JSON line:
  `sharded`  STRONG scaling of the splits BASELINE names, at every N: one fixed 64-candidate population split
             b mod N (configs[1]); one fixed 1e8-site Monte-Carlo population split N ways and 20 AK-MCS updates
             through the AKMCS class (configs[4]); Kriging.train at the configs[2] shape (n=4000, d=20, 231
             regressors) with the lock-step stencil populations of its restarts spread over the ranks.  Results are compared with a one-GPU evaluation by rank 0 inside the run.
  `configs`  (N = 1 only) device-timed throughput + roofline + sampled CPU baseline at the other shapes:
             n=2000 B=64 (the north-star Cholesky target), C3, C4 (KPLS), C5 sweeps n=40 / n=120 / d=40.

Launch:  python bench.py [--gpus N --steps K --warmup W]            (N == 1)
         python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
         python bench.py --impl reference ...    (CPU: the oracle port of the reference path)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM, POP, SWEEP_PTS = 1000, 10, 64, 10_000_000
METRIC, UNIT = "kriging_nll_evals_per_s", "evals/s"


def make_design(seed=1):
    """X ~ U(-5,5)^(1000x10), y = Styblinski-Tang."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, (N_TRAIN, DIM))
    y = 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1)
    Xn = ((X + 5.0) / 10.0 - 0.5) * 2
    return X, Xn, y, np.ones((N_TRAIN, 1))


def make_population(seed=2, B=POP, d=DIM):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, (B, d))


def config_dict(n_gpus):
    return {"workload": "BASELINE configs[1]: ordinary Kriging n=1000 d=10 gaussian nugget=1e-6; "
                        "theta-population B=64 concentrated NLL + EI sweep over 1e7 candidates per step",
            "n": N_TRAIN, "d": DIM, "population_per_gpu": POP, "sweep_points_per_gpu": SWEEP_PTS,
            "parallelism": "candidates and sweep points sharded per GPU (x%d), no data-path collective" % n_gpus,
            "l2": "inputs larger than L2 (606 MB population workspace, 800 MB candidate set per step); a new "
                  "theta-population every step"}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path (the reference itself is Python and is not on
# the GPU box; oracle/kriging_oracle.py is pinned to it by tests/golden)
# ------------------------------------------------------------------------------------------
def cpu_sample(n_nll, n_pred):
    from oracle import kriging_oracle as ko
    _, Xn, y, F = make_design()
    xpop = make_population()
    t0 = time.perf_counter()
    for b in range(n_nll):
        ko.nll(xpop[b], Xn, y, F)
    t_nll = time.perf_counter() - t0
    fit = ko.nll(np.full(DIM, 0.3), Xn, y, F, full=True)
    xq = np.random.default_rng(3).uniform(-1, 1, (n_pred, DIM))
    t0 = time.perf_counter()
    ko.predict(xq, Xn, y, F, np.zeros((1, DIM), int), fit)
    t_pred = time.perf_counter() - t0
    return n_nll / t_nll, n_pred / t_pred


def cpu_reference_equivalent(n_nll=3):
    """Cost of the UNMODIFIED reference's own operation sequence (eigvals check + six LU solves + (n,n,d) distance
    stack, likelihood_func.py:168-213 / kernelfunc.py:36-47) on this host: `oracle.nll_reference_cost`."""
    from oracle import kriging_oracle as ko
    _, Xn, y, F = make_design()
    xpop = make_population()
    t0 = time.perf_counter()
    for b in range(n_nll):
        ko.nll_reference_cost(xpop[b], Xn, y, F)
    return n_nll / (time.perf_counter() - t0)


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms are meant to use the whole host, so the
    BLAS pools are raised back to the core count (rank 0 is the only rank doing CPU work at that point)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    t_all = []
    sample_nll, sample_pred = 16, 4000          # ~2.5 s of host work per step
    for _ in range(args.warmup and 1):
        cpu_sample(2, 200)
    vals, pvals = [], []
    for _ in range(max(1, min(args.steps, 3))):
        t0 = time.perf_counter()
        v, p = cpu_sample(sample_nll, sample_pred)
        t_all.append(time.perf_counter() - t0)
        vals.append(v)
        pvals.append(p)
    v, p = float(np.mean(vals)), float(np.mean(pvals))
    cores = blas_threads()
    v_unmod = cpu_reference_equivalent(3)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(t_all)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d of the 64 candidates (n=1000,d=10) per step, numpy/scipy oracle port "
                                       "of likelihood_func.py; sweep sample %d points" % (sample_nll, sample_pred),
                             "unmodified_reference_equivalent": {
                                 "value": v_unmod, "unit": UNIT,
                                 "what": "the reference's own operation sequence (eigvals check, six LU solves on "
                                         "the triangular factors, (n,n,d) distance stack), 3 candidates"}},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sweep": {"value": p, "unit": "points/s"}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = [r.split(",") for r in self.f.read().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.f.name)
        if not rows:
            return None
        sm = [float(r[0]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any("Active" in r[3 + i] and "Not" not in r[3 + i] for r in rows)]
        busy = [s for s, r in zip(sm, rows) if float(r[2]) > 200.0] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows)}


class Ctx:
    """Rank / device plumbing shared by the blocks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)"
                             % (args.gpus, self.world))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]


def c2_kriging():
    """The configs[1] design behind the reference-facing classes (what a KADAL user constructs)."""
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
    X, _, y, _ = make_design()
    K = initkriginfo("single")
    K.update(X=X, y=y.reshape(-1, 1), lb=-5.0 * np.ones(DIM), ub=5.0 * np.ones(DIM), nrestart=1, kernel=["gaussian"],
             nugget=-6, TrendOrder=0, optimizer="lbfgsb", nvar=DIM, nsamp=N_TRAIN, problem="styblinski")
    return Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)


# ------------------------------------------------------------------------------------------
# headline: configs[1], device-timed value + e2e through the reference-facing API
# ------------------------------------------------------------------------------------------
def run_headline(cx, args):
    torch = cx.torch
    from bayesianopttools_b200 import _lib, parallel
    from bayesianopttools_b200.surrogate_models.supports.likelihood_func import likelihood, likelihood_batch
    from bayesianopttools_b200.surrogate_models.supports.device_state import get_model

    X, Xn, y, F = make_design()
    krig = c2_kriging()
    K = krig.KrigInfo
    mdl = get_model(K).model                     # the same handle the reference-facing calls use
    assert np.array_equal(np.asarray(K["X_norm"]), Xn)
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    mdl.reserve_population(POP)
    mdl.set_profiling(True)
    ybest = float(y.min())
    nsteps = args.warmup + args.steps
    # a different population every step (theta ~ U(-1,1)^10 in log10), resident in HBM
    pops_h = np.stack([make_population(seed=1000 * cx.rank + 2 + k) for k in range(nsteps)])
    nll0, info0 = mdl.nll_batch(pops_h[0], False, -6.0, return_info=True)
    pops_d = torch.as_tensor(pops_h, device=cx.dev)
    nll_d = torch.empty(nsteps, POP, dtype=torch.float64, device=cx.dev)
    info_d = torch.empty(POP, dtype=torch.int32, device=cx.dev)
    # predictor state for the sweep: fitted at the best candidate of the first population (through `likelihood`)
    likelihood(pops_h[0][int(np.argmin(nll0))], K, mode="all", trainvar=False)
    krig.predict(X[:4], ["pred"])                # configures normalisation / trend on the handle
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device=cx.dev)
    gen.manual_seed(3 + cx.rank)
    cand_d = torch.rand(SWEEP_PTS, DIM, dtype=torch.float64, device=cx.dev, generator=gen) * 10.0 - 5.0
    res_d = torch.zeros(8, dtype=torch.int64, device=cx.dev)
    offset = cx.rank * SWEEP_PTS

    def step_device(k, ev=None):
        if ev:
            ev[0].record()
        mdl.nll_batch_dev(pops_d[k].data_ptr(), POP, DIM, False, -6.0, nll_d[k].data_ptr(), info_d.data_ptr())
        if ev:
            ev[1].record()
        mdl.predict_dev(cand_d.data_ptr(), SWEEP_PTS, idx_offset=offset, acq="ei", ybest=ybest,
                        result_ptr=res_d.data_ptr())
        if ev:
            ev[2].record()

    for k in range(args.warmup):
        step_device(k)
    cx.barrier()
    sampler = ClockSampler(cx.local) if cx.rank == 0 else None
    launches0 = _lib.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        step_device(args.warmup + k, evs[k])
    cx.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    t_nll = sum(e[0].elapsed_time(e[1]) for e in evs) * 1e-3
    t_swp = sum(e[1].elapsed_time(e[2]) for e in evs) * 1e-3
    t_tot = evs[0][0].elapsed_time(evs[-1][2]) * 1e-3
    # stage times of the timed launches themselves (event ring inside the library)
    hist = [mdl.stage_ms(b) for b in range(min(args.steps, 32))]
    chol_ms = float(np.mean([h[2] for h in hist]))
    build_ms = float(np.mean([h[1] for h in hist]))
    gls_ms = float(np.mean([h[3] for h in hist]))
    pred_ms = float(np.mean([h[4] for h in hist]))
    clocks = sampler.stop() if sampler else None
    t_nll_m, t_swp_m, t_tot_m = cx.max_over_ranks([t_nll, t_swp, t_tot])
    merged = parallel.gather_sweep_result(res_d)
    best_nll = -cx.max_over_ranks([-float(nll_d[args.warmup:].min())])[0]

    # ---- e2e: the calls a KADAL user makes, pageable numpy in, numpy out, copies inside the timed region ----
    mdl.set_stream(None)                           # the handle's own stream: the host API path a user gets
    e2e_steps = max(10, args.steps)
    cand_h = cand_d.cpu().numpy()                  # pageable, 800 MB
    e_pops = [make_population(seed=5000 + 100 * cx.rank + k) for k in range(e2e_steps + 1)]
    likelihood_batch(e_pops[-1], K, trainvar=False)
    r_host, _ = krig.predict_sweep(cand_h, acq="EI")
    cx.barrier()
    e_nll = e_swp = 0.0
    for k in range(e2e_steps):
        t0 = time.perf_counter()
        f = likelihood_batch(e_pops[k], K, trainvar=False)
        t1 = time.perf_counter()
        r_host, _ = krig.predict_sweep(cand_h, acq="EI")
        t2 = time.perf_counter()
        e_nll += t1 - t0
        e_swp += t2 - t1
    cx.barrier()
    assert isinstance(f, np.ndarray) and f.shape == (POP,)
    e_nll, e_swp = cx.max_over_ranks([e_nll, e_swp])
    same_winner = (r_host["best_idx"] + offset == int(res_d[1].item()))

    out = None
    if cx.rank == 0:
        dmma_peak = _lib.measure_peak(0, device=cx.local)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            traffic = {}
        chol_flops = POP * (N_TRAIN ** 3) / 3.0
        pred_flops = float(SWEEP_PTS) * N_TRAIN * N_TRAIN           # n^2/2 FMA = n^2 flop per site
        use_all_host_threads()
        cpu_v, cpu_p = cpu_sample(48, 16000)        # ~7 s of host work (48 of the 64 candidates, 16000 sites)
        cpu_unmod = cpu_reference_equivalent(3)
        cores = blas_threads()
        world = cx.world
        out = {
            "metric": METRIC, "value": world * POP * args.steps / t_nll_m, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot_m / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world),
            "phases_ms_per_step": {"nll_population": 1e3 * t_nll_m / args.steps, "ei_sweep": 1e3 * t_swp_m / args.steps,
                                   "corr_build": build_ms, "cholesky": chol_ms, "gls": gls_ms,
                                   "wall_clock_all_steps_s": t_wall},
            "roofline": {"kernel": "kb_chol_kernel", "bound": "tensor", "achieved": chol_flops / (chol_ms * 1e-3) / 1e12,
                         "peak": dmma_peak, "unit": "TFLOP/s", "frac": chol_flops / (chol_ms * 1e-3) / 1e12 / dmma_peak,
                         "traffic": traffic.get("kb_chol_kernel"),
                         "traffic_source": "dram__bytes_read+write of one `ncu --set full` capture of this command "
                                           "(profiles/, committed); not re-measured inside this run",
                         "peak_source": "FP64 tensor (DMMA, mma.sync m8n8k4) issue-bound loop measured in this run "
                                        "by kb_measure_peak; MEASURED_PEAKS.json carries no FP64 figure "
                                        "(its hbm_gbs=%s)" % peaks.get("hbm_gbs"),
                         "algorithmic": "64 candidates x n^3/3 flop (n=1000) per launch"},
            "sweep": {"metric": "predict_ei_points_per_s", "value": world * SWEEP_PTS * args.steps / t_swp_m,
                      "unit": "points/s", "points_per_step_per_gpu": SWEEP_PTS,
                      "roofline": {"kernel": "kb_predict_kernel", "bound": "tensor",
                                   "achieved": pred_flops / (pred_ms * 1e-3) / 1e12, "peak": dmma_peak,
                                   "unit": "TFLOP/s", "frac": pred_flops / (pred_ms * 1e-3) / 1e12 / dmma_peak,
                                   "traffic": (traffic.get("kb_predict_kernel_bytes_per_site") or 0) * SWEEP_PTS or None,
                                   "algorithmic": "1e7 sites x n^2 flop (n=1000); HBM input 8*d B/site = %.1f GB/s"
                                                  % (SWEEP_PTS * DIM * 8 / (pred_ms * 1e-3) / 1e9)},
                      "e2e": {"value": world * SWEEP_PTS * e2e_steps / e_swp, "unit": "points/s",
                              "api": "Kriging.predict_sweep(numpy (1e7,10) pageable, acq='EI')",
                              "h2d_bytes_per_step": SWEEP_PTS * DIM * 8, "d2h_bytes_per_step": 64,
                              "same_winner_as_device_arm": bool(same_winner)},
                      "cpu_baseline": {"value": cpu_p, "unit": "points/s", "cores": cores, "kind": "port",
                                       "sample": "16000 sites, n=1000"},
                      "best_idx": merged["best_idx"], "best_val": merged["best_val"]},
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "48 of the 64 candidates (n=1000, d=10), numpy/scipy oracle port of "
                                       "likelihood_func.py on this box's host cores",
                             "unmodified_reference_equivalent": {
                                 "value": cpu_unmod, "unit": UNIT,
                                 "what": "the reference's own operation sequence (eigvals check, six LU solves on "
                                         "the triangular factors, (n,n,d) distance stack), 3 candidates"}},
            "e2e": {"value": world * POP * e2e_steps / e_nll, "unit": UNIT, "steps": e2e_steps,
                    "api": "likelihood_batch(numpy (64,10) pageable, KrigInfo) -> numpy (64,)",
                    "h2d_bytes_per_step": POP * DIM * 8, "d2h_bytes_per_step": POP * 12},
            "gpu_launches": int(launches), "clocks": clocks,
            "not_pd_candidates": int(np.count_nonzero(info0)), "best_nll": best_nll,
        }
    mdl.set_stream(None)
    K["_kb"].drop()
    del cand_d, cand_h
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------
# sharded: strong scaling of the three splits BASELINE names (same total work at every N)
# ------------------------------------------------------------------------------------------
def sharded_population(cx, reps=20):
    """configs[1]: ONE 64-candidate population, candidate b on rank b mod N, NLL slices all-gathered (NCCL)."""
    from bayesianopttools_b200.surrogate_models.supports.likelihood_func import likelihood_batch
    krig = c2_kriging()
    K = krig.KrigInfo
    K["shard_population"] = True
    pops = [make_population(seed=7000 + k) for k in range(reps + 3)]          # the same on every rank
    for k in range(3):
        likelihood_batch(pops[reps + k], K, trainvar=False)
    cx.barrier()
    t0 = time.perf_counter()
    for k in range(reps):
        f = likelihood_batch(pops[k], K, trainvar=False)
    cx.barrier()
    t = cx.max_over_ranks([time.perf_counter() - t0])[0]
    same = None
    if cx.rank == 0:
        K["shard_population"] = False
        same = bool(np.array_equal(likelihood_batch(pops[reps - 1], K, trainvar=False), f))
    K["_kb"].drop()
    return {"workload": "one 64-candidate theta-population (n=1000, d=10) per call through likelihood_batch, "
                        "candidate b on rank b mod N, all-gather of the NLL slices",
            "candidates_per_gpu": -(-POP // cx.world), "calls": reps, "ms_per_population": 1e3 * t / reps,
            "evals_per_s": POP * reps / t, "identical_to_one_gpu": same}


def akmcs_run(n_order, maxupdate, shard, cx):
    from bayesianopttools_b200.misc.sampling.samplingplan import sampling
    from bayesianopttools_b200.reliability_analysis.akmcs import AKMCS, mcpopgen_device
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
    from bayesianopttools_b200.testcase.RA.testcase import evaluate
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
    counters, phases = [], {"sweep": 0.0}
    sweep = ak._sweep

    def recording_sweep():
        t0 = time.perf_counter()
        sweep()
        phases["sweep"] += time.perf_counter() - t0
        counters.append((ak._res["n_fail"], ak._res["n_fail_minus"], ak._res["n_fail_plus"], ak._res["min_u_idx"],
                         ak._res["n_points"]))
    ak._sweep = recording_sweep
    cx.torch.cuda.synchronize()
    if shard:
        cx.barrier()
    t0 = time.perf_counter()
    ak.run(disp=False)
    cx.torch.cuda.synchronize()
    t = time.perf_counter() - t0
    hist = dict(updateX=ak.updateX.tolist(), minU=np.asarray(ak.minUiter).ravel().tolist(), Pf=ak.Pf, cov=ak.cov,
                stop=ak.stop_criteria, counters=counters, nsamp=ak.nsamp)
    K["_kb"].drop()
    pop.close()
    return hist, t, phases["sweep"]


def sharded_akmcs(cx, n_order=8, updates=20):
    """configs[4]: ONE 1e8-site Philox population split N ways, resident per GPU; per update one all-reduce of the
    int64 failure counters + all-gather of the (min U, index) records, the new site broadcast by its owner.  Every
    rank retrains the 12..32-point model from identical inputs (the five L-BFGS-B restarts advance in lock-step, one
    15-candidate population per round): a round is one ~0.1 ms latency-bound device call whatever the population, so
    spreading the restarts over the ranks (KrigInfo['shard_restarts']) would add a collective and save nothing."""
    shard = cx.world > 1
    akmcs_run(5, 2, shard, cx)                    # warm-up: pools, graphs, communicator
    hist, t, t_sweep = akmcs_run(n_order, updates, shard, cx)
    t, t_sweep = cx.max_over_ranks([t, t_sweep])
    nupd = len(hist["updateX"]) - 1
    out = {"workload": "AK-MCS four-branch d=2 through the AKMCS class: 1e%d-site population (one Philox stream) split "
                       "over the GPUs, %d updates incl. retraining (5 L-BFGS-B restarts)" % (n_order, nupd),
           "population": hist["nsamp"], "updates": nupd, "loop_s": t, "ms_per_update": 1e3 * t / max(1, nupd),
           "sweep_ms_per_update": 1e3 * t_sweep / (nupd + 1), "retrain_and_host_ms_per_update": 1e3 * (t - t_sweep) / max(1, nupd),
           "sites_per_s": hist["nsamp"] * (nupd + 1) / t, "Pf": hist["Pf"], "cov": hist["cov"],
           "stop_criterion": hist["stop"], "fail_counters_last": list(hist["counters"][-1][:3]),
           "exchange": "per update: all_reduce(int64[4]) of the failure counters + ONE all_gather of the 32-byte "
                       "(min, index) records with each rank's best site attached (no separate broadcast); the "
                       "retraining is replicated (serial part of the loop: ~35 dependent L-BFGS-B rounds of one "
                       "latency-bound device call each)"}
    if shard:
        same = None
        if cx.rank == 0:
            ref, _, _ = akmcs_run(n_order, updates, False, cx)      # the whole population on this GPU alone
            same = all(hist[k] == ref[k] for k in ("updateX", "minU", "Pf", "cov", "stop", "counters", "nsamp"))
        cx.barrier()
        out["identical_to_one_gpu"] = same
    return out


def c3_design(n=4000, d=20, seed=11):
    """configs[2]: Styblinski-Tang-like response on U(-1,1)^20, second-order trend (231 regressors).  The response
    varies on the scale of the box, so the likelihood optimum is INTERIOR to the log10(theta) box."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (n, d))
    w = np.linspace(0.8, 1.25, d)
    y = np.sum(np.sin(4.0 * w * X) + 0.15 * (w * X) ** 3, axis=1) + 0.5 * np.sin(X[:, 0] * X[:, 1] * 3)
    return X, y.reshape(-1, 1)


def sharded_train_c3(cx, n=4000, d=20):
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
    X, y = c3_design(n, d)
    K = initkriginfo("single")
    K.update(X=X, y=y, lb=-np.ones(d), ub=np.ones(d), nrestart=2,
             kernel=["gaussian"], nugget=-6, TrendOrder=2, optimizer="lbfgsb", nvar=d, nsamp=n, problem="c3")
    if cx.world > 1:
        K["shard_population"] = True             # the 2 x 21-candidate lock-step stencil is split b mod N
    krig = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False, lb=-2, ub=2)
    from bayesianopttools_b200.surrogate_models.supports.likelihood_func import likelihood_batch
    likelihood_batch(np.zeros((2 * (d + 1), d)), K, trainvar=False)      # warm-up: workspace, task queue
    cx.barrier()
    t0 = time.perf_counter()
    krig.train(disp=False)
    cx.barrier()
    t = cx.max_over_ranks([time.perf_counter() - t0])[0]
    th = np.asarray(K["Theta"], dtype=float)
    out = {"workload": "Kriging.train at the configs[2] shape: n=%d, d=%d, TrendOrder 2 (%d regressors), L-BFGS-B x 2 "
                       "restarts in lock-step (42-candidate stencils), candidates b mod N" % (n, d, K["F"].shape[1]),
           "train_s": t, "device_calls": int(getattr(krig, "lockstep_calls", 0)), "nll": float(K["NegLnLike"]),
           "log10_theta_min_max": [float(th.min()), float(th.max())],
           "interior_optimum": bool(th.min() > -2 + 1e-6 and th.max() < 2 - 1e-6),
           "theta_checksum": float(np.dot(th, np.cos(np.arange(d))))}
    K["_kb"].drop()
    return out


def run_sharded(cx):
    out = {"scaling": "strong", "n_gpus": cx.world}
    for name, fn in (("population", sharded_population), ("akmcs", sharded_akmcs), ("train_c3", sharded_train_c3)):
        try:
            out[name] = fn(cx)
        except Exception as e:                       # a failed block is reported, not hidden
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            if cx.world > 1:
                raise
        cx.torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------
# configs: the other BASELINE shapes, device-timed, each with its roofline and a sampled CPU baseline
# ------------------------------------------------------------------------------------------
def timed(torch, fn, reps, warm=2):
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


def styb(X5):
    return 0.5 * np.sum(X5 ** 4 - 16 * X5 ** 2 + 5 * X5, axis=1)


def cfg_nll(cx, peak, name, n, d, B, trend=0, kpls=0, cpu_n=2, reps=5):
    torch = cx.torch
    from bayesianopttools_b200 import _lib
    from oracle import kriging_oracle as ko
    rng = np.random.default_rng(1)
    X = rng.uniform(-1, 1, (n, d))
    y = styb(5 * X[:, :10]) + (0.1 * np.sum((5 * X[:, 10:]) ** 2, axis=1) if d > 10 else 0.0)
    F = np.ones((n, 1))
    if trend:
        from bayesianopttools_b200.surrogate_models.supports.trendfunction import compute_regression_mat, polytruncation
        idx = polytruncation(trend, d, 1)
        F = compute_regression_mat(idx, X, np.vstack((-np.ones(d), np.ones(d))), np.ones(d))
    kw, nth, pls = {}, d, None
    if kpls:
        pls = rng.normal(size=(d, kpls)) / np.sqrt(d)
        kw, nth = dict(model_type="kpls", plscoeff=pls), kpls
    mdl = _lib.DeviceModel(X, y, F, ["gaussian"], **kw)
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    mdl.reserve_population(B)
    mdl.set_profiling(True)
    xpop_h = np.random.default_rng(2).uniform(-0.3, 0.7, (B, nth))
    xpop = torch.tensor(xpop_h, device=cx.dev)
    nll = torch.empty(B, dtype=torch.float64, device=cx.dev)
    ms = timed(torch, lambda: mdl.nll_batch_dev(xpop.data_ptr(), B, nth, False, -6.0, nll.data_ptr()), reps)
    st = mdl.stage_ms(0)
    nll_h = nll.cpu().numpy()
    mdl.close()
    t0 = time.perf_counter()
    ref = [ko.nll(xpop_h[b], X, y, F, plscoeff=pls, n_princomp=kpls or None) for b in range(cpu_n)]
    cpu_s = (time.perf_counter() - t0) / cpu_n
    err = float(np.max(np.abs(nll_h[:cpu_n] - np.array(ref)) / np.maximum(1.0, np.abs(ref))))
    flop = B * n ** 3 / 3.0
    return {"case": name, "n": n, "d": d, "population": B, "regressors": int(F.shape[1]),
            "ms_per_population": ms, "evals_per_s": B / ms * 1e3,
            "stage_ms": {"decode": st[0], "corr_build": st[1], "cholesky": st[2], "gls": st[3]},
            "roofline": {"kernel": "kb_chol_kernel", "bound": "tensor", "achieved": flop / (st[2] * 1e-3) / 1e12,
                         "peak": peak, "unit": "TFLOP/s", "frac": flop / (st[2] * 1e-3) / 1e12 / peak,
                         "algorithmic": "%d candidates x n^3/3 flop" % B, "traffic": None},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                             "sample": "%d candidates" % cpu_n},
            "nll_rel_err_vs_oracle_on_sample": err}


def cfg_sweep(cx, peak, hbm, fp64_peak, name, n, d, m, acq, cpu_m=20000, reps=3, trend=0):
    torch = cx.torch
    from bayesianopttools_b200 import _lib
    from oracle import kriging_oracle as ko
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, (n, d))
    if d >= 10 and n >= 500:
        y = styb(5 * X[:, :10])
        th = np.full(d, 0.3)
    elif d > 4:                                     # hidimenRA (testcase/RA/testcase.py:81-88): g = d + 0.6 sqrt(d) - sum x
        X = rng.uniform(0.4, 1.6, (n, d))
        y = d + 3 * 0.2 * np.sqrt(d) - np.sum(X, axis=1)
        th = np.full(d, 0.9)
    else:                                           # AK-MCS shapes: a limit state that changes sign over the population
        y = np.sum(X ** 2, axis=1) - 0.5
        th = np.full(d, -0.2)
    F = np.ones((n, 1))
    mdl = _lib.DeviceModel(X, y, F, ["gaussian"])
    mdl.fit_state(th, False, -6.0, want_U=False, want_Psi=False)
    mdl.set_norm(_lib.NORM_NONE)
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    if acq == "u" and d > 4:
        pts = _lib.DevicePoints.generate(m, d, dist="lognormal", mean=1.0, stddev=0.2, seed=9)
    elif acq == "u":
        pts = _lib.DevicePoints.generate(m, d, dist="gaussian", mean=0.0, stddev=0.4, seed=9)
    else:
        pts = _lib.DevicePoints.generate(m, d, dist="random", lb=-np.ones(d), ub=np.ones(d), seed=9)
    res = {}

    def go():
        res["r"] = mdl.predict_points(pts, (), acq=acq, ybest=float(y.min()))[1]
    ms = timed(torch, go, reps, warm=1)
    r = res["r"]
    # an ill-conditioned fit (pivots at the nugget) sweeps through the corrected paths by default: time the plain
    # explicit-inverse product too, so that the cost of the correction is on the record
    refined, ms_plain = bool(mdl.predict_refined), None
    if refined:
        mdl.set_predict_refine(False)
        mdl.fit_state(th, False, -6.0, want_U=False, want_Psi=False)
        ms_plain = timed(torch, go, reps, warm=1)
        mdl.set_predict_refine("auto")
        mdl.fit_state(th, False, -6.0, want_U=False, want_Psi=False)
    # sampled check + CPU baseline: the first cpu_m sites of the same Philox stream
    xs = pts.rows(0, cpu_m)
    fit = ko.nll(th, X, y, F, full=True)
    t0 = time.perf_counter()
    pr = ko.predict(xs, X, y, F, np.zeros((1, d), int), fit)
    cpu_s = time.perf_counter() - t0
    sub, _ = None, None
    outs, rsub = mdl.predict(xs, ("pred", "s"), acq=acq, ybest=float(y.min()))
    pred_o, s_o = np.asarray(pr["pred"]).ravel(), np.asarray(pr["s"]).ravel()
    chk = {"pred_rel_err": float(np.max(np.abs(outs["pred"].ravel() - pred_o)) / np.max(np.abs(pred_o))),
           "s_rel_err": float(np.max(np.abs(outs["s"].ravel() - s_o)) / np.max(np.abs(s_o)))}
    if acq == "u":
        chk["fail_counters_device"] = [int(rsub["n_fail"]), int(rsub["n_fail_minus"]), int(rsub["n_fail_plus"])]
        chk["fail_counters_oracle"] = [int(np.sum(pred_o <= 0)), int(np.sum(pred_o - 1.96 * s_o <= 0)),
                                       int(np.sum(pred_o + 1.96 * s_o <= 0))]
    mdl.close()
    pts.close()
    flop = float(m) * n * n
    inb = m * d * 8 / (ms * 1e-3) / 1e9
    # FP64 instructions the path must issue per site: psi n(3d+25)/... counted as n*(2d+18) scalar FP64 ops + n^2/2 FMA
    # on the DMMA pipe; both share ONE pipe, so the issue roof is peak_fp64_flops / flop_per_site
    flop_issue = n * n + 2.0 * n * (2 * d + 18)
    out = {"case": name, "n": n, "d": d, "sites": m, "ms": ms, "sites_per_s": m / ms * 1e3, "acq": acq,
           "roofline": ({"kernel": "kb_predict_kernel", "bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12,
                         "peak": peak, "unit": "TFLOP/s", "frac": flop / (ms * 1e-3) / 1e12 / peak,
                         "algorithmic": "n^2 flop per site", "traffic": None} if n > 128 else
                        {"kernel": "kb_predict_warp_kernel", "bound": "fp64_issue",
                         "achieved": m * flop_issue / (ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": m * flop_issue / (ms * 1e-3) / 1e12 / fp64_peak,
                         "algorithmic": "per site n^2 flop on the DMMA pipe + 2n(2d+18) scalar FP64 flop for the "
                                        "correlations (one shared FP64 pipe: the roof is its issue rate, not HBM)",
                         "traffic": None}),
           "hbm": {"input_GBps": inb, "frac_of_measured_hbm": inb / hbm},
           "cpu_baseline": {"value": cpu_m / cpu_s, "unit": "points/s", "cores": blas_threads(), "kind": "port",
                            "sample": "%d sites" % cpu_m},
           "check_on_sample": chk, "n_fail": int(r["n_fail"]), "min_idx": int(r["min_u_idx"] if acq == "u" else r["best_idx"]),
           "ill_conditioned_fit": refined,
           "explicit_inverse_product": ({"ms": ms_plain, "sites_per_s": m / ms_plain * 1e3,
                                         "what": "the same sweep with the corrected paths switched off "
                                                 "(KrigInfo['predict_refine'] = False): faster, 5-30x further from an "
                                                 "exact evaluation of the variance on such fits (DESIGN section 2)"}
                                        if refined else None)}
    return out


def run_configs(cx):
    from bayesianopttools_b200 import _lib
    peak = _lib.measure_peak(0, device=cx.local)
    fp64 = _lib.measure_peak(1, device=cx.local)
    hbm = 6554.2
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    use_all_host_threads()
    out = {"dmma_peak_tflops": peak, "fp64_scalar_peak_tflops": fp64, "hbm_peak_GBps": hbm, "cases": []}
    jobs = [
        lambda: cfg_nll(cx, peak, "north-star Cholesky target: n=2000 d=10 B=64", 2000, 10, 64, cpu_n=3),
        lambda: cfg_nll(cx, peak, "C2 one GPU's share of an 8-GPU population: n=1000 d=10 B=8", 1000, 10, 8, cpu_n=4),
        lambda: cfg_nll(cx, peak, "C3 trend n=4000 d=20 order 2 (231 regressors) B=8", 4000, 20, 8, trend=2, cpu_n=2, reps=3),
        lambda: cfg_nll(cx, peak, "C4 KPLS n=2000 d=100 h=3 B=64", 2000, 100, 64, kpls=3, cpu_n=3),
        lambda: cfg_sweep(cx, peak, hbm, fp64, "sweep n=2000 d=10, 1e6 sites (EI)", 2000, 10, 1_000_000, "ei", cpu_m=4000),
        lambda: cfg_sweep(cx, peak, hbm, fp64, "C5 AK-MCS U sweep n=40 d=2, 1e8 sites", 40, 2, 100_000_000, "u"),
        lambda: cfg_sweep(cx, peak, hbm, fp64, "C5 AK-MCS late model n=120 d=2, 1e8 sites", 120, 2, 100_000_000, "u"),
        lambda: cfg_sweep(cx, peak, hbm, fp64, "C5 hidimenRA n=100 d=40 (lognormal population), 2e7 sites", 100, 40, 20_000_000, "u"),
    ]
    for j in jobs:
        try:
            out["cases"].append(j())
        except Exception as e:
            out["cases"].append({"error": "%s: %s" % (type(e).__name__, e)})
        cx.torch.cuda.empty_cache()
    return out


def run_gpu(args):
    from bayesianopttools_b200 import _lib
    _lib.require_gpu()
    cx = Ctx(args)
    line = run_headline(cx, args)
    sharded = run_sharded(cx) if not args.no_sharded else None
    configs = run_configs(cx) if (cx.world == 1 and cx.rank == 0 and not args.no_configs) else None
    if cx.rank == 0:
        line["sharded"] = sharded
        line["configs"] = configs
        emit(line)
    if cx.world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()


_RESULT_OUT = None


def emit(line):
    """The one JSON line of the contract, written to the process's ORIGINAL stdout."""
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries write to stdout behind Python's back (NCCL prints its version banner there when the communicator
    # comes up): file descriptor 1 is pointed at stderr for the whole run and only emit() uses the real stdout.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-sharded", action="store_true", help="skip the strong-scaling block")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE shapes (N = 1 block)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
