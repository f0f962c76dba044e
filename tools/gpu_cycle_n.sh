mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_n_build.log 2>&1
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files > gpurun_out/r02_n_bench_$name.json 2>> gpurun_out/r02_n_bench.err
  python - gpurun_out/r02_n_bench_$name.json "$name" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], 'e2e ms', round(d['e2e']['ms_per_step'], 2), 'threads', d['e2e']['host_threads'], 'value ms', round(d['ms_per_step'], 3))
PY
}
run base MGB_X=0
run nt MGB_HOST_NT=1
run nt_c128k MGB_HOST_NT=1 MGB_HOST_CHUNK=131072 MGB_HOST_RING=8
run nt_c256k MGB_HOST_NT=1 MGB_HOST_CHUNK=262144 MGB_HOST_RING=8
run nt_c1m MGB_HOST_NT=1 MGB_HOST_CHUNK=1048576 MGB_HOST_RING=6
run nt_t14 MGB_HOST_NT=1 MGB_HOST_THREADS=14
run nt_chunk MGB_HOST_NT=1 MGB_HOST_SPLIT=chunk MGB_HOST_RING=32
run base2 MGB_X=0
MGB_HOST_NT=1 python tools/seam_stats.py > gpurun_out/r02_n_seam_stats_nt.txt 2>&1
