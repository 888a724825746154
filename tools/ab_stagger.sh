#!/bin/bash
# Staggered-queue sweep of the Cholesky (KB_CHOL_GROUPS x KB_CHOL_STAGGER): tools/ab_stagger.sh
out=gpurun_out/ab_stagger.log; : > $out
for rep in 1 2; do
  for gs in "1 0" "2 2" "2 4" "2 6" "2 8" "4 1" "4 2" "4 3" "4 4" "8 1" "8 2"; do
    set -- $gs
    echo "== groups $1 stagger $2 rep $rep" >> $out
    for c in nll1000 nll2000; do
      KB_CHOL_GROUPS=$1 KB_CHOL_STAGGER=$2 timeout 120 python tools/quick_perf.py $c >> $out 2>&1
    done
  done
done
