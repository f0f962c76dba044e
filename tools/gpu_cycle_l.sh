mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us 2>/dev/null
python -m matchering_b200.build > gpurun_out/r02_l_build.log 2>&1
(timeout 600 python -m pytest tests -m gpu -q -x -k "golden or host_seam or other_configs" 2>&1 | tail -4) > gpurun_out/r02_l_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files > gpurun_out/r02_l_bench_c2.json 2> gpurun_out/r02_l_bench.err
for t in 8 12 24; do MGB_HOST_THREADS=$t python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files > gpurun_out/r02_l_bench_c2_t$t.json 2>> gpurun_out/r02_l_bench.err; done
cat gpurun_out/r02_l_tests.log
