#!/bin/bash
# A/B on the GPU box: persistent convolution, chained analysis twiddles (twice: its 6-lane figure was odd once),
# the new transport defaults through bench.py.
mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_u_build.log 2>&1
(timeout 600 python -m pytest tests -m gpu -q -x -k "kernel_variants or host_seam or frame_lengths" 2>&1 | tail -4) > gpurun_out/r02_u_tests.log
run() { tag=$1; shift; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files "$@" > gpurun_out/r02_u_bench_$tag.json 2> gpurun_out/r02_u_bench_$tag.err; }
run default
run persist --opt conv_persistent=1
run achain --opt analyze_chain=1
run achain2 --opt analyze_chain=1
run both --opt analyze_chain=1 --opt conv_persistent=1
tail -2 gpurun_out/r02_u_tests.log
python - <<'PY'
import json
for n in ("default","persist","achain","achain2","both"):
    try:
        d=json.loads(open(f"gpurun_out/r02_u_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), round(d["single_track_latency"]["ms_per_step"]*1e3,1), {k:round(v["avg_ms"]*1e3,1) for k,v in d["kernels"].items()}, round(d["e2e"]["ms_per_step"],2), d["e2e"]["d2h_bytes_per_step"])
    except Exception as e: print(n, "failed", e)
PY
