#!/bin/bash
# The round's evidence run on the GPU box: the default bench line (with cpu_baseline), the ncu launch
# list of a three-step run, and one full ncu capture of every pipeline kernel.  Outputs in gpurun_out/.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_err.log
tail -c 600 gpurun_out/bench_${TAG}_n1.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python tools/one_step.py 180 3 > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/launches_${TAG}.csv
timeout 900 ncu --set full --clock-control none --import-source on \
  -k "regex:analyze_kernel|spectrum_mean|smooth_operator_kernel|design_kernel|convolve|clip_sumsq|correction_final|limiter_kernel" \
  --launch-count 30 -f -o gpurun_out/prof_${TAG} python tools/one_step.py 180 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_${TAG}.ncu-rep
