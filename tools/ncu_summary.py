"""Summarise an .ncu-rep (read here, not on the GPU box): key launch / pipe / memory / stall metrics as text,
and DRAM traffic per launch as JSON for bench.py's roofline.traffic.   python tools/ncu_summary.py rep [title]"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"]
print(title)
name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
print("kernel:", name)
out = {}
for i, h in enumerate(hdr):
    if h in keys or ("issue_stalled" in h and "per_issue_active" in h):
        print(f"{h:92s} {units[i]:16s} {vals[i]}")
    if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
        out[h] = float(vals[i]) * mult
print("dram bytes per launch (read+write): %.4g" % sum(out.values()))
print("JSON " + json.dumps({"dram_bytes_per_launch": sum(out.values())}))
