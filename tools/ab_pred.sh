#!/bin/bash
# A/B of library variants on the sweep shapes: tools/ab_pred.sh <variant> [<variant> ...]
out=gpurun_out/ab_pred.log; : > $out
for rep in 1 2; do
  for v in "" "$@"; do
    lib=bayesianopttools_b200/csrc/libkrigb200${v:+_$v}.so
    echo "== ${v:-product} rep $rep" >> $out
    KRIGB200_LIB=$PWD/$lib timeout 120 python tools/quick_perf.py pred 1000 10 20 >> $out 2>&1
    KRIGB200_LIB=$PWD/$lib timeout 120 python tools/quick_perf.py pred 2000 10 19 >> $out 2>&1
    KRIGB200_LIB=$PWD/$lib timeout 120 python tools/quick_perf.py pred 600 10 20 >> $out 2>&1
  done
done
