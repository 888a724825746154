timeout -k 5 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2_21.log 2>&1; echo "pytest rc=$?"
timeout -k 5 200 python tools/quick_perf.py > gpurun_out/quick_perf_s2_22.log 2>&1; echo "perf rc=$?"
