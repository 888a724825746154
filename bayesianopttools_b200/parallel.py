"""Multi-GPU sharding of the Kriging hot path (one process per GPU, torch.distributed).

The path partitions into independent units:
  * hyper-parameter candidates: candidate b goes to rank b mod G; the design (X, y, F) is
    replicated.  The only exchange is an all-gather of the per-rank NLL slices (8 B per
    candidate) followed by a local arg-min with lowest-index tie-break.
  * prediction / Monte-Carlo sites: contiguous shards of m/G points; the fit state is
    replicated.  The exchange is an all-reduce (sum) of the int64 failure counters (akmcs.py:208-238)
    and an all-gather of one 32-byte (min acquisition, index, min U, index) record per rank, merged
    locally and deterministically (lowest index on ties) -- NCCL has no arg-reduce.
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


def gather_sharded_nll(local_nll, local_info, xpop):
    """ONE all-gather per sharded likelihood call: every rank contributes its NLL slice, its info flags and a
    (population size, checksum) pair of the population it was handed; the slices are put back into candidate order
    and the pairs are compared -- every rank must have submitted the same rows (same optimiser state, same seed),
    otherwise the slices of different populations would be interleaved silently.  Returns (nll[B], info[B])."""
    rank, size = world()
    B = len(xpop)
    local = np.asarray(local_nll, dtype=np.float64)
    info = np.asarray(local_info)
    if size == 1:
        return local, info
    x = np.ascontiguousarray(xpop, dtype=np.float64)
    w = np.cos(np.arange(x.size, dtype=np.float64) * 0.7310585786) + 1.5      # position-dependent weights
    chk = float(np.dot(x.ravel(), w)) if x.size else 0.0
    if not np.isfinite(chk):
        chk = float(np.count_nonzero(~np.isfinite(x)))
    per = (B + size - 1) // size
    row = np.full(2 * per + 2, np.inf)
    row[:len(local)] = local
    row[per:per + len(info)] = info
    row[2 * per], row[2 * per + 1] = float(B), chk
    mine = torch.as_tensor(row).to(_comm_device())
    out = torch.empty(size * len(row), dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    rows = out.cpu().numpy().reshape(size, len(row))
    if np.any(rows[:, 2 * per] != float(B)) or np.any(rows[:, 2 * per + 1] != chk):
        raise ValueError("shard_population: the ranks submitted different populations (sizes %s); every rank must run "
                         "the same deterministic optimiser with the same seed" % sorted(set(rows[:, 2 * per].tolist())))
    full, flags = np.full(B, np.inf), np.zeros(B, dtype=info.dtype if info.size else np.int32)
    for r in range(size):
        idx = candidate_slice(B, r, size)
        full[idx] = rows[r, :len(idx)]
        flags[idx] = rows[r, per:per + len(idx)].astype(flags.dtype)
    return full, flags


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


def gather_sweep_result(local, payload=None):
    """Merge one reduction record per rank.  `local` is a dict with RESULT_FIELDS, or a (8,) int64-viewable
    device tensor laid out like KbSweepResult.  The failure counters and the site count are summed with
    `all_reduce` (int64, exact); the two (min, index) pairs are all-gathered and merged with the lowest index
    winning ties, which is what np.argmin over the whole population returns.
    `payload`: a float64 row every rank attaches to its (min U, index) record -- the coordinates of its own best
    site -- so that the winner's row arrives with the same all-gather (no separate broadcast); the merged dict
    then carries it as "min_u_payload"."""
    rank, size = world()
    if isinstance(local, torch.Tensor):
        raw = local.view(torch.int64).reshape(8)
    else:
        vals = np.zeros(8, dtype=np.int64)
        vals[[0, 2]] = np.array([local["best_val"], local["min_u"]], dtype=np.float64).view(np.int64)
        vals[[1, 3, 4, 5, 6, 7]] = [local[k] for k in ("best_idx", "min_u_idx", "n_fail", "n_fail_minus",
                                                       "n_fail_plus", "n_points")]
        raw = torch.as_tensor(vals)
    npay = 0 if payload is None else int(np.size(payload))
    pays = None
    if size > 1:
        raw = raw.to(_comm_device())
        counts = raw[4:8].clone()
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)               # n_fail, n_fail_minus, n_fail_plus, n_points
        mins = raw[0:4].contiguous()
        if npay:
            pay = torch.as_tensor(np.ascontiguousarray(payload, dtype=np.float64).reshape(-1).view(np.int64).copy())
            mins = torch.cat([mins, pay.to(mins.device)])
        out = torch.empty(size * mins.numel(), dtype=torch.int64, device=mins.device)
        dist.all_gather_into_tensor(out, mins)
        counts = counts.cpu().numpy()
        rows = out.cpu().numpy().reshape(size, -1)
        outs = [rows[r, 0:4] for r in range(size)]
        if npay:
            pays = [rows[r, 4:].view(np.float64) for r in range(size)]
    else:
        a = raw.cpu().numpy()
        counts, outs = a[4:8], [a[0:4]]
        if npay:
            pays = [np.asarray(payload, dtype=np.float64).reshape(-1)]
    recs = []
    for r, a in enumerate(outs):
        f = np.ascontiguousarray(a).view(np.float64)
        recs.append(dict(best_val=float(f[0]), best_idx=int(a[1]), min_u=float(f[2]), min_u_idx=int(a[3]),
                         n_fail=0, n_fail_minus=0, n_fail_plus=0, n_points=0))
    out = merge_results(recs)
    out.update(n_fail=int(counts[0]), n_fail_minus=int(counts[1]), n_fail_plus=int(counts[2]), n_points=int(counts[3]))
    if pays is not None:
        win = [r for r, q in enumerate(recs) if q["min_u_idx"] == out["min_u_idx"] and q["min_u"] == out["min_u"]][0]
        out["min_u_payload"] = np.array(pays[win], dtype=np.float64)
    return out


def assert_same_population(xpop):
    """Sharded likelihood evaluation slices ONE population over the ranks: every rank must have submitted the same
    rows (same optimiser state, same seed).  One all-reduce of (B, -B, checksum, -checksum) with MAX catches a
    rank that drew different random numbers before the slices are silently interleaved."""
    rank, size = world()
    if size == 1:
        return
    x = np.ascontiguousarray(xpop, dtype=np.float64)
    w = np.cos(np.arange(x.size, dtype=np.float64) * 0.7310585786) + 1.5      # position-dependent weights
    chk = float(np.dot(x.ravel(), w)) if x.size else 0.0
    if not np.isfinite(chk):
        chk = float(np.count_nonzero(~np.isfinite(x)))
    v = torch.tensor([float(len(x)), -float(len(x)), chk, -chk], dtype=torch.float64, device=_comm_device())
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    v = v.cpu().numpy()
    if v[0] != -v[1] or v[2] != -v[3]:
        raise ValueError("shard_population: the ranks submitted different populations (sizes %g..%g); every rank "
                         "must run the same deterministic optimiser with the same seed" % (-v[1], v[0]))


def broadcast_array(a, src=0):
    """A float64 array of a shape every rank knows, sent by rank `src`."""
    rank, size = world()
    a = np.ascontiguousarray(a, dtype=np.float64)
    if size == 1:
        return a
    buf = torch.as_tensor(a.copy()).to(_comm_device())
    dist.broadcast(buf, src=src)
    return buf.cpu().numpy().reshape(a.shape)


def allgather_rows(row):
    """Every rank contributes one float64 row of the same length; returns the (size, len) matrix."""
    rank, size = world()
    row = np.ascontiguousarray(row, dtype=np.float64).reshape(-1)
    if size == 1:
        return row[None, :]
    mine = torch.as_tensor(row.copy()).to(_comm_device())
    outs = [torch.empty_like(mine) for _ in range(size)]
    dist.all_gather(outs, mine)
    return np.stack([o.cpu().numpy() for o in outs])


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
