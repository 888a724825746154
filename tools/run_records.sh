#!/bin/bash
# How the round-2 records under profiles/ were produced on one B200 box (run from the repository root through gpurun;
# the ncu reports are summarised on the box with tools/ncu_summary.py because they are too large to travel back).
# --- single-GPU records ---
python -m pytest tests -m gpu -q > gpurun_out/r02_gputests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_gputests.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-sharded --no-configs > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/parity_report.py > gpurun_out/r02_parity_report.txt 2>&1
(for e in 0 1; do KB_PRED_REFINE=$e python tools/illcond_check.py; done) > gpurun_out/r02_illcond_check.txt 2>&1
timeout 900 python tools/fuzz_parity.py 320 11 > gpurun_out/r02_fuzz.txt 2>&1; tail -4 gpurun_out/r02_fuzz.txt
timeout 600 python tools/accuracy_vs_longdouble.py > gpurun_out/r02_accuracy_vs_longdouble.txt 2>&1; tail -1 gpurun_out/r02_accuracy_vs_longdouble.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?"
# --- ncu summaries of the hot kernels ---
cap() {  # name kernel-regex title command...
  local name=$1 rx=$2 title=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -o /tmp/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep "$title" > gpurun_out/r02_ncu_$name.txt 2>&1
}
H="ncu --set full --clock-control none --import-source on (round 2, final build)"
cap chol kb_chol_kernel "$H -- kb_chol_kernel, n=1000 d=10 B=64 (python tools/profile_run.py nll 2)" python tools/profile_run.py nll 2
cap chol_n2000 kb_chol_kernel "$H -- kb_chol_kernel, n=2000 d=10 B=64 (python tools/quick_perf.py nll2000)" python tools/quick_perf.py nll2000
cap pred kb_predict_kernel "$H -- kb_predict_kernel, n=1000 d=10, 2^19 sites (python tools/profile_run.py predict 1)" python tools/profile_run.py predict 1
KB_PRED_REFINE=0 cap warp kb_predict_warp "$H -- kb_predict_warp_kernel (plain product), n=40 d=2, 2^22 sites (KB_PRED_REFINE=0 python tools/profile_run.py small 1)" python tools/profile_run.py small 1
cap warp_subst kb_predict_warp "$H -- kb_predict_warp_kernel (blocked substitution: the fit is flagged), n=40 d=2, 2^22 sites (python tools/profile_run.py small 1)" python tools/profile_run.py small 1
cap warp_d40 kb_predict_warp "$H -- kb_predict_warp_kernel, n=100 d=40, 2^21 sites (python tools/quick_perf.py pred 100 40 21)" python tools/quick_perf.py pred 100 40 21
cap build kb_build_aug "$H -- kb_build_aug_kernel, n=1000 d=10 B=64" python tools/profile_run.py nll 1
# --- multi-GPU records (gpurun --gpus N) ---
# python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N > gpurun_out/r02_bench_nN.json
# --- further fuzz seeds ---
# python tools/fuzz_parity.py 320 5 ; python tools/fuzz_parity.py 320 23
