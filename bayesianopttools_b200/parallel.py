"""Multi-GPU sharding of the Kriging hot path (one process per GPU, torch.distributed).

The path partitions into independent units:
  * hyper-parameter candidates: candidate b goes to rank b mod G; the design (X, y, F) is
    replicated.  The only exchange is an all-gather of the per-rank NLL slices (8 B per
    candidate) followed by a local arg-min with lowest-index tie-break.
  * prediction / Monte-Carlo sites: contiguous shards of m/G points; the fit state is
    replicated.  The exchange is an all-gather of one 64-byte reduction record per rank
    (min acquisition + index, min U + index, three failure counters), merged locally and
    deterministically -- NCCL has no arg-reduce, and summing int64 counters after a gather is
    exact and order independent.
No data-path collective exists: NCCL (NVLink) carries only these records.  The same code runs
over gloo on CPU tensors (tests/test_parallel_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as dist

RESULT_FIELDS = ("best_val", "best_idx", "min_u", "min_u_idx", "n_fail", "n_fail_minus", "n_fail_plus",
                 "n_points")


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _comm_device():
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def candidate_slice(B, rank, size):
    """Indices of the candidates owned by `rank` (b mod G == rank)."""
    return np.arange(rank, B, size)


def point_shard(m, rank, size):
    """Contiguous [start, stop) shard of m points for `rank`."""
    base, rem = divmod(int(m), size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_population_nll(local_nll, B):
    """All-gather the per-rank NLL slices back into candidate order; returns (nll[B], argmin)."""
    rank, size = world()
    local = np.asarray(local_nll, dtype=np.float64)
    if size == 1:
        full = local
    else:
        per = (B + size - 1) // size
        buf = torch.full((per,), float("inf"), dtype=torch.float64, device=_comm_device())
        buf[:len(local)] = torch.as_tensor(local, dtype=torch.float64)
        outs = [torch.empty_like(buf) for _ in range(size)]
        dist.all_gather(outs, buf)
        full = np.full(B, np.inf)
        for r in range(size):
            idx = candidate_slice(B, r, size)
            full[idx] = outs[r].cpu().numpy()[:len(idx)]
    best = int(np.flatnonzero(full == np.min(full))[0]) if np.isfinite(np.min(full)) else int(np.argmin(full))
    return full, best


def merge_results(records):
    """Deterministic merge of per-shard sweep records (dicts with RESULT_FIELDS)."""
    out = dict(records[0])
    for q in records[1:]:
        if q["best_val"] < out["best_val"] or (q["best_val"] == out["best_val"] and q["best_idx"] < out["best_idx"]):
            out["best_val"], out["best_idx"] = q["best_val"], q["best_idx"]
        if q["min_u"] < out["min_u"] or (q["min_u"] == out["min_u"] and q["min_u_idx"] < out["min_u_idx"]):
            out["min_u"], out["min_u_idx"] = q["min_u"], q["min_u_idx"]
        for k in ("n_fail", "n_fail_minus", "n_fail_plus", "n_points"):
            out[k] += q[k]
    return out


def gather_sweep_result(local):
    """All-gather one reduction record per rank and merge.  `local` is a dict with
    RESULT_FIELDS, or a (8,) int64-viewable device tensor laid out like KbSweepResult."""
    rank, size = world()
    if isinstance(local, torch.Tensor):
        raw = local.view(torch.int64).reshape(8)
    else:
        vals = np.zeros(8, dtype=np.int64)
        vals[[0, 2]] = np.array([local["best_val"], local["min_u"]], dtype=np.float64).view(np.int64)
        vals[[1, 3, 4, 5, 6, 7]] = [local[k] for k in ("best_idx", "min_u_idx", "n_fail", "n_fail_minus",
                                                       "n_fail_plus", "n_points")]
        raw = torch.as_tensor(vals)
    if size > 1:
        raw = raw.to(_comm_device())
        outs = [torch.empty_like(raw) for _ in range(size)]
        dist.all_gather(outs, raw)
    else:
        outs = [raw]
    recs = []
    for o in outs:
        a = o.cpu().numpy()
        f = a.view(np.float64)
        recs.append(dict(best_val=float(f[0]), best_idx=int(a[1]), min_u=float(f[2]), min_u_idx=int(a[3]),
                         n_fail=int(a[4]), n_fail_minus=int(a[5]), n_fail_plus=int(a[6]), n_points=int(a[7])))
    return merge_results(recs)


def gather_counts(n_local):
    """Sizes of every rank's shard (all-gather of one int64), e.g. to turn local site indices into
    indices of the whole Monte-Carlo population."""
    rank, size = world()
    if size == 1:
        return [int(n_local)]
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=_comm_device())
    outs = [torch.empty_like(mine) for _ in range(size)]
    dist.all_gather(outs, mine)
    return [int(o.item()) for o in outs]


def owner_of(index, counts):
    """Rank whose contiguous shard holds global site `index`, and the index inside that shard."""
    start = 0
    for r, c in enumerate(counts):
        if index < start + c:
            return r, int(index - start)
        start += c
    raise IndexError("site %d is outside the %d-site population" % (index, start))


def broadcast_row(row, src, d):
    """The d coordinates of one site, sent by rank `src` to every rank (`row` is ignored elsewhere)."""
    rank, size = world()
    if size == 1:
        return np.asarray(row, dtype=np.float64).reshape(-1)
    buf = torch.zeros(d, dtype=torch.float64)
    if rank == src:
        buf = torch.as_tensor(np.asarray(row, dtype=np.float64).reshape(-1).copy())
    buf = buf.to(_comm_device())
    dist.broadcast(buf, src=src)
    return buf.cpu().numpy()
