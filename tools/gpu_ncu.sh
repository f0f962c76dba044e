#!/bin/bash
# Full ncu capture of one launch of the kernels matching a regex (after one warm-up step), into
# gpurun_out/prof_r01_TAG.ncu-rep.  usage: tools/gpu_ncu.sh TAG 'convolve|limiter' [seconds]
TAG=${1:-x}
REGEX=${2:-.}
SECONDS_=${3:-180}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$REGEX" \
  --launch-skip-before-match 0 --launch-skip ${4:-1} --launch-count ${5:-1} -f -o gpurun_out/prof_r01_$TAG \
  python tools/one_step.py $SECONDS_ 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/prof_r01_$TAG.ncu-rep
