mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_q_build.log 2>&1
(timeout 600 python -m pytest tests -m gpu -q -x -k "host_seam or host or seam" 2>&1 | tail -4) > gpurun_out/r02_q_tests.log
for r in 0 1; do
  MGB_DOWNLOAD_RING=$r python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/r02_q_bench_c5_ring$r.json 2>> gpurun_out/r02_q_bench.err
  python - gpurun_out/r02_q_bench_c5_ring$r.json "c5 ring_download=$r" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], 'e2e ms', round(d['e2e']['ms_per_step'], 2), 'threads', d['e2e']['host_threads'], 'value ms', round(d['ms_per_step'], 3))
PY
done
MGB_HOST_STATS=1 python tools/seam_ab.py 2 "t12 512K x8 nt512  ring" 2>&1 | tail -8
python tools/seam_ab.py 25 "64K x16 plain" "512K x8" "1M x6" "t13" > gpurun_out/r02_q_seam_ab.txt 2>&1
cat gpurun_out/r02_q_seam_ab.txt gpurun_out/r02_q_tests.log
