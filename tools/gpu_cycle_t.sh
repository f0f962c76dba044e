#!/bin/bash
# A/B on the GPU box: FFT addressing change (bench kernels), analyze_chain, host transport variants (CLDEMOTE ring,
# chunk geometry of the ring download, flush).
mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_t_build.log 2>&1
(grep -m1 "model name" /proc/cpuinfo; grep -o -m1 -w "cldemote" /proc/cpuinfo; lscpu | grep -E "L3|L2|Socket|NUMA|Thread|Core" ; cat /sys/fs/cgroup/cpu.max) > gpurun_out/r02_t_host.txt 2>&1
(timeout 600 python -m pytest tests -m gpu -q -x -k "fft or golden or frame_lengths" 2>&1 | tail -4) > gpurun_out/r02_t_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files > gpurun_out/r02_t_bench_default.json 2> gpurun_out/r02_t_bench_default.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files --opt analyze_chain=1 > gpurun_out/r02_t_bench_achain.json 2> gpurun_out/r02_t_bench_achain.err
timeout 300 python tools/seam_ab.py 12 > gpurun_out/r02_t_seam_ab.txt 2>&1
cat gpurun_out/r02_t_host.txt; tail -2 gpurun_out/r02_t_tests.log
python - <<'PY'
import json
for n in ("default","achain"):
    try:
        d=json.loads(open(f"gpurun_out/r02_t_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), round(d["single_track_latency"]["ms_per_step"]*1e3,1), {k:round(v["avg_ms"]*1e3,1) for k,v in d["kernels"].items()}, round(d["e2e"]["ms_per_step"],2))
    except Exception as e: print(n, "failed", e)
PY
cat gpurun_out/r02_t_seam_ab.txt
