mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_i_build.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q -x -k "golden or fir or other_configs or half_minute or process or lowess or direct_smoothing" 2>&1 | tail -8) > gpurun_out/r02_i_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_i_bench_c2.json 2> gpurun_out/r02_i_bench_c2.err
python tools/process_profile.py 2>&1 | head -30 > gpurun_out/r02_i_process_profile.log
cat gpurun_out/r02_i_tests.log
