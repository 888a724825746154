timeout -k 5 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2_9.log 2>&1; echo "pytest rc=$?"
timeout -k 5 200 python tools/quick_perf.py > gpurun_out/quick_perf_s2_10.log 2>&1; echo "perf rc=$?"
KB_CHOL_TRACE=gpurun_out/trace_s2_7.bin timeout -k 5 120 python tools/profile_run.py nll 1 > gpurun_out/tr_s2_7.log 2>&1
