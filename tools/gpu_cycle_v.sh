#!/bin/bash
mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_v_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_v_tests.log
tail -30 gpurun_out/r02_v_tests.log
