mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_e_build.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r02_e_tests.log
python bench.py --steps 20 --warmup 5 --no-files --no-cpu-baseline > gpurun_out/r02_e_bench_c2.json 2> gpurun_out/r02_e_bench_c2.err
cat gpurun_out/r02_e_tests.log
