mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_k_build.log 2>&1
(timeout 600 python -m pytest tests -m gpu -q -x -k "host_seam or golden or process" 2>&1 | tail -4) > gpurun_out/r02_k_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files > gpurun_out/r02_k_bench_c2.json 2> gpurun_out/r02_k_bench.err
python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_k_bench_c5.json 2>> gpurun_out/r02_k_bench.err
MGB_HOST_CHUNK=262144 MGB_HOST_RING=12 python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_k_bench_c5_bigring.json 2>> gpurun_out/r02_k_bench.err
cat gpurun_out/r02_k_tests.log
