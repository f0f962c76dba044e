#!/bin/bash
# Validation of the round's final host transport on the GPU box: GPU tests, the config-2 bench line, the
# in-process A/B of the transport geometries, one call with transfer statistics.
mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_s_build.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r02_s_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_s_bench_c2.json 2> gpurun_out/r02_s_bench_c2.err
timeout 300 python tools/seam_ab.py 15 > gpurun_out/r02_s_seam_ab.txt 2>&1
(MGB_HOST_STATS=1 timeout 200 python tools/seam_ab.py 1 "t12 512K x8 nt512  dma" "t12 512K x8 nt512  ring" 2>&1 | tail -30) > gpurun_out/r02_s_seam_stats.txt
tail -3 gpurun_out/r02_s_tests.log; tail -c 1500 gpurun_out/r02_s_bench_c2.json | head -c 600; echo; cat gpurun_out/r02_s_seam_ab.txt
