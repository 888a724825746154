"""Analyse a KB_CHOL_TRACE dump: per-task phases (dependency wait / GEMM / diag wait / epilogue)
of the bulk tile tasks and the chain steps (kb_chol.cu)."""
import sys
import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8)
raw = raw[raw[:, 0] > 0]
t0 = raw[:, 0].min()
T = (raw[:, :5].astype(np.int64) - int(t0)) / 1e3          # us
T[:, 3] = np.where(raw[:, 3] > 0, T[:, 3], 0)
meta = raw[:, 5]
b, ti, tk, ty = (meta >> 40).astype(int), ((meta >> 24) & 0xffff).astype(int), ((meta >> 8) & 0xffff).astype(int), (meta & 0xff).astype(int)
end = T[:, 4].max()
print("records", len(raw), "span %.1f us" % end)
names = {0: "CHAIN", 1: "OFF", 2: "SCHUR", 3: "PRE", 4: "THIN", 5: "SCHURT", 6: "OFF2", 7: "PRE_A", 8: "SCHURTA", 9: "OFFS"}
for code, name in names.items():
    m = ty == code
    if not m.any():
        continue
    dep = T[m, 1] - T[m, 0]
    gemm = T[m, 2] - T[m, 1]
    dw = np.where(T[m, 3] > 0, T[m, 3] - T[m, 2], 0)
    epi = T[m, 4] - np.where(T[m, 3] > 0, T[m, 3], T[m, 2])
    tot = T[m, 4] - T[m, 0]
    pollus = (raw[m, 6] >> np.uint64(16)).astype(np.int64) / 1e3 if code != 0 else np.zeros(m.sum())
    print(f"{name:6s} n={m.sum():5d} total {tot.sum()/1e3:8.2f} ms-CTA | mean us: depwait {dep.mean():6.2f} gemm {gemm.mean():6.2f} "
          f"diagwait {dw.mean():6.2f} epilogue {epi.mean():6.2f} | sums ms: dep {dep.sum()/1e3:.2f} gemm {gemm.sum()/1e3:.2f} "
          f"dw {dw.sum()/1e3:.2f} epi {epi.sum()/1e3:.2f} | polling inside the K loop: {pollus.sum()/1e3:.2f} ms, mean {pollus.mean():.2f} us, tasks with > 5 us: {(pollus > 5).sum()}")
bulk = ty != 0
slots = np.unique(raw[bulk, 7])
raw6 = raw[:, 6] & np.uint64(0xffff)
busy = (T[bulk, 4] - T[bulk, 0]).sum()
print("bulk CTAs", len(slots), "busy fraction %.2f" % (busy / (end * len(slots))))
cb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m0 = (b == cb) & (ty == 0)
print("candidate %d chain steps: k, start, end, dur (us) | PRE(k) end | gap to previous step end" % cb)
prev = 0.0
for r in np.flatnonzero(m0)[np.argsort(tk[m0])]:
    k = tk[r]
    pm = (b == cb) & (ty == 3) & (ti == k)
    pre_end = T[pm, 4][0] if pm.any() else float("nan")
    pre_start = T[pm, 0][0] if pm.any() else float("nan")
    ph = (raw[r, [1, 2, 3, 6, 4]].astype(np.int64) - int(raw[r, 0])) / 1e3 if raw[r, 2] > 0 else None
    print("  k=%2d start %8.1f end %8.1f dur %5.1f | PRE start %8.1f end %8.1f | idle before %6.1f | phases(load,trsm,prod,factor,store) %s" % (k, T[r, 0], T[r, 4], T[r, 4] - T[r, 0], pre_start, pre_end, T[r, 0] - prev, "" if ph is None else " ".join("%.1f" % v for v in np.diff(np.concatenate([[0], ph])))))
    prev = T[r, 4]
# timeline of utilisation: number of bulk tasks in flight per 50 us bin
bins = np.arange(0, end + 50, 50)
occ = np.zeros(len(bins) - 1)
for s, e in zip(T[bulk, 0], T[bulk, 4]):
    i0, i1 = int(s // 50), int(e // 50)
    for i in range(i0, min(i1, len(occ) - 1) + 1):
        lo, hi = max(s, bins[i]), min(e, bins[i + 1])
        occ[i] += max(0.0, hi - lo) / 50
print("bulk tasks in flight per 50us bin:", " ".join("%d" % round(o) for o in occ))
