cd /root/repo
timeout 300 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for big in 1 0; do
  echo "== KB_PW_BIG=$big"
  for shape in "120 2 23" "100 2 23" "72 2 23" "128 6 22" "96 4 23"; do KB_PW_BIG=$big timeout 60 python tools/quick_perf.py pred $shape 2>&1 | tail -1; done
done
