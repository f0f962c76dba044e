mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_c_build.log 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q -k "lowess or host_seam or process or checker or pcm or golden" 2>&1 | tail -5) > gpurun_out/r02_c_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c_bench_c2.json 2> gpurun_out/r02_c_bench_c2.err
python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c_bench_c5.json 2> gpurun_out/r02_c_bench_c5.err
python tools/process_profile.py > gpurun_out/r02_c_process_profile.log 2>&1
cat gpurun_out/r02_c_tests.log
