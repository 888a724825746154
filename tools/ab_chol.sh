#!/bin/bash
# A/B of library variants on the Cholesky-dominated shapes: tools/ab_chol.sh <variant> [<variant> ...]
out=gpurun_out/ab_chol.log; : > $out
for rep in 1 2; do
  for v in "" "$@"; do
    lib=bayesianopttools_b200/csrc/libkrigb200${v:+_$v}.so
    echo "== ${v:-product} rep $rep" >> $out
    KRIGB200_LIB=$PWD/$lib timeout 120 python tools/quick_perf.py nll1000 >> $out 2>&1
    KRIGB200_LIB=$PWD/$lib timeout 120 python tools/quick_perf.py nll2000 >> $out 2>&1
  done
done
