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
