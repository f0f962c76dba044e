#!/bin/bash
mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_w_build.log 2>&1
timeout 300 python tools/seam_ab.py 15 > gpurun_out/r02_w_seam_ab.txt 2>&1
cat gpurun_out/r02_w_seam_ab.txt
