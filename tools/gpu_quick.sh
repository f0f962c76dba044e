#!/bin/bash
# one test selection on the GPU box:  gpurun -- 'bash tools/gpu_quick.sh "<pytest -k expression>"'
mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/quick_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "$1" 2>&1 | tail -25 | tee gpurun_out/quick_tests.log
