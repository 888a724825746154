timeout -k 5 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2_23.log 2>&1; echo "pytest rc=$?"
