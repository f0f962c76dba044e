mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_b_build.log 2>&1
(timeout 600 python -m pytest tests -m gpu -x -q -k "smoothing or golden or frame_lengths or half_minute" 2>&1 | tail -5) > gpurun_out/r02_b_tests.log
python bench.py --steps 20 --warmup 5 --no-files > gpurun_out/r02_b_bench_c2.json 2> gpurun_out/r02_b_bench_c2.err
python bench.py --steps 20 --warmup 5 --no-files --no-cpu-baseline --opt conv_frame=2 > gpurun_out/r02_b_bench_c2_frame2.json 2>> gpurun_out/r02_b_bench_c2.err
python tools/seam_sweep.py > gpurun_out/r02_b_seam_sweep.log 2>&1
python tools/process_profile.py > gpurun_out/r02_b_process_profile.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_b_launches.csv python bench.py --steps 2 --warmup 1 --no-files --no-cpu-baseline > gpurun_out/r02_b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"convolve_fused_kernel|limiter_kernel|analyze_kernel" -s 4 -c 4 -o gpurun_out/r02_b_full python bench.py --steps 2 --warmup 1 --no-files --no-cpu-baseline --lanes 1 > gpurun_out/r02_b_ncu_full.log 2>&1
cat gpurun_out/r02_b_tests.log
