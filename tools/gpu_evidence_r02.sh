#!/bin/bash
# Round 2's evidence run on the GPU box (one call): the GPU test suite, the bench lines of BASELINE configs 2, 3
# and 5, the reference arm, the ncu launch list of a three-step run, one full ncu capture of every pipeline kernel
# (config 2, one step) and of the limiter on config 5's buffer, compute-sanitizer over the new code paths.
# The ncu reports are summarised ON THE BOX (tools/r02_summary.py, tools/ncu_lines.py) and dropped if the
# outputs would exceed what gpurun copies back (64 MiB); text outputs land in gpurun_out/, the generated
# profiles/ files in gpurun_out/profiles_out/.
TAG=${1:-r02_final}
mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/${TAG}_build.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --no-files > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-files > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_ncu_launches.csv python tools/one_step.py 180 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k "regex:analyze_kernel|spectrum_mean|smooth_operator_kernel|design_kernel|convolve|clip_sumsq|correction_final|limiter_kernel" \
  --launch-skip 11 --launch-count 11 -f -o gpurun_out/${TAG}_prof python tools/one_step.py 180 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:limiter_kernel" --launch-count 1 -f -o gpurun_out/${TAG}_prof_c5 \
  python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/${TAG}_ncu_c5.log 2>&1
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x \
  -k "second_order or resampl or lowess or frame_lengths or kernel_variants or pipeline_matches_golden or fft_size_16384" 2>&1 | tail -6) > gpurun_out/${TAG}_sanitizer.txt
(timeout 600 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x \
  -k "second_order or pipeline_matches_golden or kernel_variants" 2>&1 | tail -6) >> gpurun_out/${TAG}_sanitizer.txt
(timeout 300 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "host_seam_results" 2>&1 | tail -60) > gpurun_out/${TAG}_sanitizer_host_seam.txt
# ---- summaries on the box
(
  for k in limiter_kernel convolve analyze_kernel spectrum_mean design_kernel; do
    echo "## ${k} (stall samples / executed instructions by source line)"
    python tools/ncu_lines.py gpurun_out/${TAG}_prof.ncu-rep $k 14
    echo
  done
) > gpurun_out/${TAG}_hot_lines.txt 2>&1
python tools/r02_summary.py ${TAG} > gpurun_out/${TAG}_summary.log 2>&1
mkdir -p gpurun_out/profiles_out
cp profiles/r02_summary.md profiles/r02_ncu_full_metrics.txt profiles/traffic.json gpurun_out/profiles_out/ 2>/dev/null
ls -la gpurun_out > gpurun_out/${TAG}_files.txt
total_mb=$(du -sm gpurun_out | cut -f1)
if [ "$total_mb" -gt 58 ]; then rm -f gpurun_out/${TAG}_prof_c5.ncu-rep; fi
total_mb=$(du -sm gpurun_out | cut -f1)
if [ "$total_mb" -gt 58 ]; then rm -f gpurun_out/${TAG}_prof.ncu-rep; fi
du -sm gpurun_out
tail -3 gpurun_out/${TAG}_tests.log; tail -c 300 gpurun_out/${TAG}_bench_c2.json; echo; cat gpurun_out/${TAG}_sanitizer.txt; head -30 gpurun_out/${TAG}_hot_lines.txt
