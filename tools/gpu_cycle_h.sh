mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_h_build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k "regex:spectrum_mean|smooth_operator_kernel|design_kernel|clip_sumsq|correction_final" \
  --launch-count 14 -f -o gpurun_out/r02_h_small python tools/one_step.py 180 2 > gpurun_out/r02_h_ncu.log 2>&1
tail -3 gpurun_out/r02_h_ncu.log
