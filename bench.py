#!/usr/bin/env python
"""bench.py -- BASELINE.json configs[1] on N GPUs of one node.

Workload ("SOBO d=10, n=1000 ordinary Kriging"): one STEP is one pass of the hot path over one
batch of synthetic input on every rank:
  phase A  population-batched concentrated ln-likelihood, B = 64 hyper-parameter candidates
           (decode -> correlation build -> DMMA Cholesky -> GLS reductions)
  phase B  fused predict + EI sweep with device arg-min over M = 1e7 candidate sites
`value` is BASELINE's first metric (likelihood evaluations / s) over phase A of the timed
steps; phase B's throughput (points / s) is reported in `sweep`.  Weak scaling: every rank owns
its own population and candidate shard (global population 64 N, global candidates 1e7 N); the
only communication is the final (best NLL, index) / sweep-record gather (parallel.py).

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


def make_population(seed=2, B=POP):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, (B, DIM))


def config_dict(n_gpus):
    return {"workload": "BASELINE configs[1]: ordinary Kriging n=1000 d=10 gaussian nugget=1e-6; "
                        "theta-population B=64 concentrated NLL + EI sweep over 1e7 candidates per step",
            "n": N_TRAIN, "d": DIM, "population_per_gpu": POP, "sweep_points_per_gpu": SWEEP_PTS,
            "parallelism": "candidates and sweep points sharded per GPU (x%d), no data-path collective" % n_gpus,
            "l2": "inputs larger than L2 (606 MB population workspace, 800 MB candidate set per step)"}


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
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(t_all)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d of the 64 candidates (n=1000,d=10) per step, numpy/scipy oracle port "
                                       "of likelihood_func.py; sweep sample %d points" % (sample_nll, sample_pred)},
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


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from bayesianopttools_b200 import _lib, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    _lib.require_gpu()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    X, Xn, y, F = make_design()
    mdl = _lib.DeviceModel(Xn, y, F, ["gaussian"], device=local)
    mdl.set_stream(torch.cuda.current_stream().cuda_stream)
    mdl.reserve_population(POP)
    mdl.set_profiling(True)
    lb, ub = -5.0 * np.ones(DIM), 5.0 * np.ones(DIM)
    mdl.set_norm(_lib.NORM_DEFAULT, lb, ub)
    ybest = float(y.min())
    # predictor state for the sweep: fitted at the best candidate of a first population pass
    xpop_h = make_population(seed=2 + rank)
    nll0, info0 = mdl.nll_batch(xpop_h, False, -6.0, return_info=True)
    mdl.fit_state(xpop_h[int(np.argmin(nll0))], False, -6.0, want_U=False, want_Psi=False)

    xpop_d = torch.as_tensor(xpop_h, device=dev)
    nll_d = torch.empty(POP, dtype=torch.float64, device=dev)
    info_d = torch.empty(POP, dtype=torch.int32, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(3 + rank)
    cand_d = torch.rand(SWEEP_PTS, DIM, dtype=torch.float64, device=dev, generator=gen) * 10.0 - 5.0
    res_d = torch.zeros(8, dtype=torch.int64, device=dev)
    offset = rank * SWEEP_PTS

    def step_device(ev=None):
        if ev:
            ev[0].record()
        mdl.nll_batch_dev(xpop_d.data_ptr(), POP, DIM, False, -6.0, nll_d.data_ptr(), info_d.data_ptr())
        if ev:
            ev[1].record()
        mdl.predict_dev(cand_d.data_ptr(), SWEEP_PTS, idx_offset=offset, acq="ei", ybest=ybest,
                        result_ptr=res_d.data_ptr())
        if ev:
            ev[2].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        step_device(evs[k])
    barrier()
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

    times = torch.tensor([t_nll, t_swp, t_tot], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_nll_m, t_swp_m, t_tot_m = [float(v) for v in times.cpu()]

    # the one real exchange: per-rank best NLL and the 64-byte sweep records (NCCL all-gather)
    merged = parallel.gather_sweep_result(res_d)
    best_nll = float(nll_d.min())
    if world > 1:
        bl = torch.tensor([best_nll], dtype=torch.float64, device=dev)
        outs = [torch.empty_like(bl) for _ in range(world)]
        dist.all_gather(outs, bl)
        best_nll = min(float(o) for o in outs)

    # ---- e2e through the host C-ABI (pinned host buffers, copies inside the timed region) ----
    xpop_pin = torch.as_tensor(xpop_h).pin_memory()
    nll_pin = torch.empty(POP, dtype=torch.float64).pin_memory()
    cand_pin = torch.empty(SWEEP_PTS, DIM, dtype=torch.float64).pin_memory()
    cand_pin.copy_(cand_d)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 3))

    def step_host():
        a = mdl.nll_batch(xpop_pin.numpy(), False, -6.0)
        nll_pin.numpy()[:] = a
        t1 = time.perf_counter()
        _, r = mdl.predict(cand_pin.numpy(), (), acq="ei", ybest=ybest)
        return t1, r

    mdl.set_stream(None)          # the handle's own stream: the host API path a user gets
    step_host()
    barrier()
    e_nll = e_swp = 0.0
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        t1, r_host = step_host()
        t2 = time.perf_counter()
        e_nll += t1 - t0
        e_swp += t2 - t1
    barrier()
    et = torch.tensor([e_nll, e_swp], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e_nll, e_swp = [float(v) for v in et.cpu()]

    if rank == 0:
        dmma_peak = _lib.measure_peak(0, device=local)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            traffic = {}
        chol_flops = POP * (N_TRAIN ** 3) / 3.0
        pred_flops = float(SWEEP_PTS) * N_TRAIN * N_TRAIN           # n^2/2 FMA = n^2 flop per site
        use_all_host_threads()
        cpu_v, cpu_p = cpu_sample(48, 16000)        # ~7 s of host work (48 of the 64 candidates, 16000 sites)
        cores = blas_threads()
        line = {
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
                              "h2d_bytes_per_step": SWEEP_PTS * DIM * 8, "d2h_bytes_per_step": 64},
                      "cpu_baseline": {"value": cpu_p, "unit": "points/s", "cores": cores, "kind": "port",
                                       "sample": "16000 sites, n=1000"},
                      "best_idx": merged["best_idx"], "best_val": merged["best_val"]},
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "48 of the 64 candidates (n=1000, d=10), numpy/scipy oracle port of "
                                       "likelihood_func.py on this box's host cores"},
            "e2e": {"value": world * POP * e2e_steps / e_nll, "unit": UNIT,
                    "h2d_bytes_per_step": POP * DIM * 8, "d2h_bytes_per_step": POP * 12},
            "gpu_launches": int(launches), "clocks": clocks,
            "not_pd_candidates": int(np.count_nonzero(info0)), "best_nll": best_nll,
        }
        emit(line)
    mdl.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
