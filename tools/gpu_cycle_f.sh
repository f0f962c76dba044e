mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_f_build.log 2>&1
(timeout 600 python -m pytest tests -m gpu -q -x -k "second_order or resamples or golden or gain_envelopes" 2>&1 | tail -40) > gpurun_out/r02_f_tests.log
python bench.py --steps 20 --warmup 5 --no-files --no-cpu-baseline > gpurun_out/r02_f_bench_c2.json 2> gpurun_out/r02_f_bench_c2.err
cat gpurun_out/r02_f_tests.log
