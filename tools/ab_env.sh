#!/bin/bash
# A/B of a run-time switch on the Cholesky-dominated shapes: tools/ab_env.sh VAR=VALUE [cases...]
out=gpurun_out/ab_env.log; : > $out
sw=$1; shift
cases=${@:-nll1000 nll2000 nll1000b8 c3}
for rep in 1 2; do
  for v in "" "$sw"; do
    echo "== ${v:-default} rep $rep" >> $out
    for c in $cases; do
      env $v timeout 120 python tools/quick_perf.py $c >> $out 2>&1
    done
  done
done
