mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_j_build.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q -x -k "limiter or config5 or host_seam or golden" 2>&1 | tail -6) > gpurun_out/r02_j_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_j_bench_c2.json 2> gpurun_out/r02_j_bench_c2.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files --opt limiter_ticket=1 > gpurun_out/r02_j_bench_c2_ticket.json 2>> gpurun_out/r02_j_bench_c2.err
python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_j_bench_c5.json 2>> gpurun_out/r02_j_bench_c2.err
python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline --opt limiter_ticket=1 > gpurun_out/r02_j_bench_c5_ticket.json 2>> gpurun_out/r02_j_bench_c2.err
cat gpurun_out/r02_j_tests.log
