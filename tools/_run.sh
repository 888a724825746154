timeout -k 5 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2_20.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 5 > gpurun_out/bench_s2_4.json 2> gpurun_out/bench_s2_4.err; echo "bench rc=$?"
