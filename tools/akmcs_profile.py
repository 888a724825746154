"""Where one AK-MCS update spends its host time (cProfile over the bench's AK-MCS loop, one GPU)."""
import cProfile
import os
import pstats
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

cx = types.SimpleNamespace(torch=torch, world=1, rank=0, barrier=lambda: torch.cuda.synchronize())
n_order = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bench.akmcs_run(5, 2, False, cx)
pr = cProfile.Profile()
pr.enable()
hist, t, ts = bench.akmcs_run(n_order, 20, False, cx)
pr.disable()
print("loop %.4f s, sweeps %.4f s, updates %d" % (t, ts, len(hist["updateX"]) - 1))
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
