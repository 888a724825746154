timeout -k 5 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2_24.log 2>&1; echo "pytest rc=$?"
timeout -k 5 200 python tools/quick_perf.py pred1000 > gpurun_out/quick_perf_s2_24.log 2>&1; echo "perf rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:kb_predict_kernel -s 1 -c 1 python tools/profile_run.py predict 2 > gpurun_out/ncu_pred_traffic.log 2>&1
