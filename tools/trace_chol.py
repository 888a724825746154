"""Analyse a KB_CHOL_TRACE dump: per-task phases (dependency wait / GEMM / diag wait / epilogue)."""
import sys
import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8)
raw = raw[raw[:, 0] > 0]
t0 = raw[:, 0].min()
T = (raw[:, :5].astype(np.int64) - int(t0)) / 1e3          # us
meta = raw[:, 5]
b, ti, tk, ty = (meta >> 40).astype(int), ((meta >> 24) & 0xffff).astype(int), ((meta >> 8) & 0xffff).astype(int), (meta & 0xff).astype(int)
end = T[:, 4].max()
print("tasks", len(raw), "span %.1f us" % end)
for name, code in (("DIAG", 0), ("OFF", 1), ("SCHUR", 2)):
    m = ty == code
    if not m.any():
        continue
    dep = T[m, 1] - T[m, 0]
    gemm = T[m, 2] - T[m, 1]
    dw = np.where(T[m, 3] > 0, T[m, 3] - T[m, 2], 0)
    epi = T[m, 4] - np.where(T[m, 3] > 0, T[m, 3], T[m, 2])
    tot = T[m, 4] - T[m, 0]
    print(f"{name:5s} n={m.sum():5d} total {tot.sum()/1e3:8.2f} ms-CTA | mean us: depwait {dep.mean():6.2f} gemm {gemm.mean():6.2f} "
          f"diagwait {dw.mean():6.2f} epilogue {epi.mean():6.2f} | sums ms: dep {dep.sum()/1e3:.2f} gemm {gemm.sum()/1e3:.2f} "
          f"dw {dw.sum()/1e3:.2f} epi {epi.sum()/1e3:.2f}")
busy = (T[:, 4] - T[:, 0]).sum()
nslots = len(np.unique(raw[:, 7]))
print("CTA slots", nslots, "busy fraction %.2f" % (busy / (end * nslots)))
# critical path of matrix 0: diag tasks timeline
m0 = (b == 0) & (ty == 0)
order = np.argsort(tk[m0])
print("matrix 0 diag tasks: k, start, depend, gemmend, end (us)")
for r in np.flatnonzero(m0)[order]:
    print("  k=%2d start %8.1f dep %8.1f gemm %8.1f end %8.1f" % (tk[r], T[r, 0], T[r, 1], T[r, 2], T[r, 4]))
m1 = (b == 0) & (ty == 1) & (ti == tk + 1)
print("matrix 0 sub-diagonal OFF tasks: k, start, dep, gemm, diagready, end")
for r in np.flatnonzero(m1)[np.argsort(tk[m1])]:
    print("  k=%2d start %8.1f dep %8.1f gemm %8.1f dr %8.1f end %8.1f" % (tk[r], T[r, 0], T[r, 1], T[r, 2], T[r, 3], T[r, 4]))
