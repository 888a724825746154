ncu --set full --clock-control none --import-source on -k regex:bench -s 3 -c 1 -o gpurun_out/prof_diagf ./tools/ubench/diagf > gpurun_out/ncu_diagf.log 2>&1
