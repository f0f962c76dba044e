mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_g_build.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q -k "second_order or limiter or resampl or golden or config5" 2>&1 | tail -15) > gpurun_out/r02_g_tests.log
python bench.py --steps 20 --warmup 5 --no-files --no-cpu-baseline > gpurun_out/r02_g_bench_c2.json 2> gpurun_out/r02_g_bench_c2.err
for lanes in 2 4 6; do python bench.py --steps 24 --warmup 6 --no-files --no-cpu-baseline --lanes $lanes > gpurun_out/r02_g_bench_c2_lanes$lanes.json 2>> gpurun_out/r02_g_bench_c2.err; done
cat gpurun_out/r02_g_tests.log
