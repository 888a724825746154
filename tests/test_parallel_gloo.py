"""CPU, world_size=2 over gloo: the sharding + exchange logic of bayesianopttools_b200.parallel."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesianopttools_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        B = 7
        allf = np.array([3.0, 1.5, np.inf, 1.5, 9.0, 2.0, 4.0])        # tie between 1 and 3
        mine = parallel.candidate_slice(B, rank, size)
        full, best = parallel.gather_population_nll(allf[mine], B)
        assert np.array_equal(full, allf) and best == 1
        # sweep: shard 11 points, fabricate per-shard records from a global table
        m = 11
        vals = np.array([5., 2., 7., 2., 9., 8., 2., 6., 5., 4., 3.])
        u = np.array([1., .5, .7, .9, .5, .8, .6, .5, .9, 1., 2.])
        g = np.array([1., -1., 0., 2., -3., 1., 1., 0., 5., -2., 1.])
        lo, hi = parallel.point_shard(m, rank, size)
        sl = slice(lo, hi)
        local = dict(best_val=float(vals[sl].min()), best_idx=int(lo + np.argmin(vals[sl])),
                     min_u=float(u[sl].min()), min_u_idx=int(lo + np.argmin(u[sl])),
                     n_fail=int((g[sl] <= 0).sum()), n_fail_minus=int((g[sl] - 1 <= 0).sum()),
                     n_fail_plus=int((g[sl] + 1 <= 0).sum()), n_points=hi - lo)
        res = parallel.gather_sweep_result(local)
        assert res["best_idx"] == int(np.argmin(vals)) and res["best_val"] == 2.0
        assert res["min_u_idx"] == int(np.argmin(u))
        assert res["n_fail"] == int((g <= 0).sum()) and res["n_points"] == m
        assert res["n_fail_minus"] == int((g - 1 <= 0).sum()) and res["n_fail_plus"] == int((g + 1 <= 0).sum())
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


class _FakeKrig:
    """Stands in for the fitted Kriging model (the device is not available on CPU): mean = the limit
    state plus a bias that shrinks with every added sample, s = distance to the nearest sample.  Only
    the reductions matter here -- they follow csrc/kb_predict.cu (first index on ties)."""

    def __init__(self, X, y):
        self.KrigInfo = dict(X=np.array(X, dtype=float), y=np.array(y, dtype=float).reshape(-1, 1), nvar=2,
                             nsamp=len(X))

    def standardize(self):
        pass

    def train(self, disp=False):
        pass

    def predict_sweep(self, x, acq="u", want=()):
        x = np.asarray(x, dtype=float)
        X = self.KrigInfo["X"]
        g = _limit_state(x) + 2.0 / len(X) * np.sin(3 * x[:, 0])
        s = np.sqrt(((x[:, None, :] - X[None, :, :]) ** 2).sum(-1)).min(1) + 0.01
        u = np.abs(g) / s
        res = dict(best_val=float(u.min()), best_idx=int(np.argmin(u)), min_u=float(u.min()),
                   min_u_idx=int(np.argmin(u)), n_fail=int((g <= 0).sum()),
                   n_fail_minus=int((g - 1.96 * s <= 0).sum()), n_fail_plus=int((g + 1.96 * s <= 0).sum()),
                   n_points=len(x))
        return res, dict(pred=g, s=s)


def _limit_state(x):
    x = np.atleast_2d(x)
    return 2.5 - x[:, 0] ** 2 - 0.5 * x[:, 1]


def _akmcs_history(pop, shard):
    from bayesianopttools_b200.reliability_analysis.akmcs import AKMCS
    X0 = np.array([[0.0, 0.0], [1.0, -1.0], [-2.0, 1.0], [2.0, 2.0]])
    krig = _FakeKrig(X0, _limit_state(X0))
    info = dict(init_samp=pop, maxupdate=4, evaluate=_limit_state, resident_population=False,
                shard_population=shard)
    ak = AKMCS(krig, info)
    ak.run(disp=False)
    return ak, np.hstack([ak.updateX, ak.minUiter.reshape(-1, 1)]), (ak.Pf, ak.stop_criteria, ak.cov, ak.nsamp)


def _akmcs_worker(rank, size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        pop = np.random.default_rng(11).normal(size=(1001, 2))         # uneven shards: 501 + 500
        pop[700] = pop[3]                                              # a tie across the shard boundary
        lo, hi = parallel.point_shard(len(pop), rank, size)
        ak, hist, tail = _akmcs_history(pop[lo:hi], True)
        assert ak.sharded and ak.nsamp == 1001 and ak.nsamp_local == hi - lo and ak.Gx.shape == (hi - lo, 1)
        _, hist1, tail1 = _akmcs_history(pop, False)                   # the whole population on one rank
        assert np.array_equal(hist, hist1), (hist, hist1)
        assert tail == tail1, (tail, tail1)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_akmcs_sharded_over_two_ranks_equals_one_rank():
    """AKMCS.run with the Monte-Carlo population split over 2 ranks (counter sum, min-U merge, owner
    broadcast of the new site) walks exactly the path of the unsharded run."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_akmcs_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(out) == [(0, "ok"), (1, "ok")], out


class _FakeHolder:
    """likelihood_batch only needs holder.model.nll_batch: a deterministic stand-in for the device."""

    class model:
        calls = []

        @classmethod
        def nll_batch(cls, xpop, trainvar, nugget, return_info=False):
            cls.calls.append(len(xpop))
            f = np.sum((xpop - 0.25) ** 2, axis=1) + 3.0
            info = (xpop[:, 0] > 0.9).astype(np.int32)
            f = np.where(info != 0, np.inf, f)
            return (f, info) if return_info else f


def _likelihood_worker(rank, size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from bayesianopttools_b200.surrogate_models.supports import likelihood_func as lf
        lf.get_model = lambda K: _FakeHolder
        K = dict(kernel=["gaussian"], nvar=3, nugget=-6, shard_population=True)
        xpop = np.random.default_rng(5).uniform(0, 1, (7, 3))
        want, want_info = _FakeHolder.model.nll_batch(xpop, False, -6.0, return_info=True)
        _FakeHolder.model.calls.clear()
        got, info = lf.likelihood_batch(xpop, K, trainvar=False, return_info=True)
        assert np.array_equal(got, want) and np.array_equal(info, want_info) and info.dtype == want_info.dtype
        assert _FakeHolder.model.calls == [4 - rank]                    # 7 candidates: ranks own 4 and 3
        assert np.array_equal(lf.likelihood_batch(xpop, K, trainvar=False), want)
        # fewer candidates than ranks, or the switch off: evaluated locally, no exchange
        _FakeHolder.model.calls.clear()
        assert np.array_equal(lf.likelihood_batch(xpop[:1], K, trainvar=False), want[:1])
        K["shard_population"] = False
        assert np.array_equal(lf.likelihood_batch(xpop, K, trainvar=False), want)
        assert _FakeHolder.model.calls == [1, 7]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_likelihood_population_sharded_over_two_ranks():
    """likelihood_batch with KrigInfo['shard_population']: rank r evaluates candidates b = r mod 2 and the
    all-gathered population equals the unsharded evaluation, flags included."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_likelihood_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(out) == [(0, "ok"), (1, "ok")], out


def _sweep_worker(rank, size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from bayesianopttools_b200.optim_tools.acquifunc_opt import run_single_opt
        X0 = np.array([[0.0, 0.0], [1.0, -1.0], [-2.0, 1.0], [2.0, 2.0]])
        krig = _FakeKrig(X0, _limit_state(X0))
        krig.KrigInfo.update(lb=-3 * np.ones(2), ub=3 * np.ones(2))
        info = dict(acquifuncopt="sweep", acquifunc="EI", nrestart=1, ncandidates=777, sweep_polish=False)
        x1, f1 = run_single_opt(krig, dict(info, rng=np.random.default_rng(9)))
        x2, f2 = run_single_opt(krig, dict(info, rng=np.random.default_rng(9), shard_candidates=True))
        assert np.array_equal(x1, x2) and f1 == f2, (x1, x2, f1, f2)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _mobo_sweep_worker(rank, size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from bayesianopttools_b200.optim_tools import acquifunc_opt as ao

        def fake_ehvi_best(cand, ypar, moboInfo, kriglist):        # stands in for the device EHVI arg-max
            v = np.cos(3 * cand[:, 0]) * np.sin(2 * cand[:, 1]) + 1.5
            i = int(np.argmax(v))
            return i, float(v[i])
        ao.ehvi_best = fake_ehvi_best
        X0 = np.zeros((3, 2))
        krig = _FakeKrig(X0, np.zeros(3))
        krig.KrigInfo.update(lb=-3 * np.ones(2), ub=3 * np.ones(2))
        info = dict(acquifuncopt="sweep", acquifunc="ehvi", nrestart=1, ncandidates=1001, sweep_polish=False,
                    refpoint=np.array([1.0, 1.0]))
        ypar = np.array([[0.0, 1.0], [1.0, 0.0]])
        x1, f1 = ao.run_multi_opt([krig, krig], dict(info, rng=np.random.default_rng(4)), ypar)
        x2, f2 = ao.run_multi_opt([krig, krig], dict(info, rng=np.random.default_rng(4), shard_candidates=True), ypar)
        assert np.array_equal(x1, x2) and f1 == f2, (x1, x2, f1, f2)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_mobo_sweep_sharded_over_two_ranks_equals_one_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mobo_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(out) == [(0, "ok"), (1, "ok")], out


def test_sobo_sweep_sharded_over_two_ranks_equals_one_rank():
    """run_single_opt(acquifuncopt='sweep', shard_candidates=True): 777 candidates split 389 + 388, the merged
    winner and its coordinates equal the unsharded sweep on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(out) == [(0, "ok"), (1, "ok")], out


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(out) == [(0, "ok"), (1, "ok")], out


def test_shards_cover_everything():
    for m, size in ((10, 3), (7, 8), (100000007, 8), (1, 2)):
        spans = [parallel.point_shard(m, r, size) for r in range(size)]
        assert spans[0][0] == 0 and spans[-1][1] == m
        assert all(spans[i][1] == spans[i + 1][0] for i in range(size - 1))
    for B, size in ((64, 8), (7, 2), (3, 4)):
        idx = np.sort(np.concatenate([parallel.candidate_slice(B, r, size) for r in range(size)]))
        assert np.array_equal(idx, np.arange(B))


def test_owner_lookup_and_single_rank_helpers():
    counts = [5, 0, 3, 4]
    assert [parallel.owner_of(i, counts) for i in (0, 4, 5, 7, 8, 11)] == [(0, 0), (0, 4), (2, 0), (2, 2), (3, 0), (3, 3)]
    try:
        parallel.owner_of(12, counts)
        raise AssertionError("expected IndexError")
    except IndexError:
        pass
    # without an initialised process group every helper degenerates to the local answer
    assert parallel.world() == (0, 1)
    assert parallel.gather_counts(17) == [17]
    assert np.array_equal(parallel.broadcast_row([1.5, -2.0], 0, 2), np.array([1.5, -2.0]))
    full, best = parallel.gather_population_nll([3.0, 1.0, 1.0], 3)
    assert best == 1 and np.array_equal(full, [3.0, 1.0, 1.0])
