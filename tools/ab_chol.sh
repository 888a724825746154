#!/bin/bash
# A/B of library variants on the Cholesky-dominated shapes: tools/ab_chol.sh <variant> [<variant> ...]
out=gpurun_out/ab_chol.log; : > $out
for rep in 1 2; do
  for v in "" "$@"; do
    lib=bayesianopttools_b200/csrc/libkrigb200${v:+_$v}.so
    echo "== ${v:-product} rep $rep" >> $out
    for c in nll1000 nll2000 nll1000b8 c3; do
      KRIGB200_LIB=$PWD/$lib timeout 120 python tools/quick_perf.py $c >> $out 2>&1
    done
  done
done
