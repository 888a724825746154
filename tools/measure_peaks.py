"""Measure the roofline denominators MEASURED_PEAKS.json lacks (FP64 DMMA / DFMA) on this GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianopttools_b200 import _lib

out = {"dmma_tflops": _lib.measure_peak(0), "dfma_tflops": _lib.measure_peak(1),
       "d2d_copy_gbs": _lib.measure_peak(2)}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peaks.json", "w"))
