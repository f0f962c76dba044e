mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_d_build.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r02_d_tests.log
python tools/seam_sweep2.py > gpurun_out/r02_d_seam_sweep2.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_d_bench_c2.json 2> gpurun_out/r02_d_bench_c2.err
python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_d_bench_c5.json 2> gpurun_out/r02_d_bench_c5.err
cat gpurun_out/r02_d_tests.log
