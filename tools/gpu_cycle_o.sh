mkdir -p gpurun_out
lscpu | head -30 > gpurun_out/r02_o_lscpu.txt
python -m matchering_b200.build > gpurun_out/r02_o_build.log 2>&1
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files > gpurun_out/r02_o_bench_$name.json 2>> gpurun_out/r02_o_bench.err
  python - gpurun_out/r02_o_bench_$name.json "$name" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], 'e2e ms', round(d['e2e']['ms_per_step'], 2), 'threads', d['e2e']['host_threads'], 'value ms', round(d['ms_per_step'], 3))
PY
}
C="MGB_HOST_CHUNK=262144 MGB_HOST_RING=8"
run nt2 $C
run nt1 $C MGB_HOST_NT=1
run nt2_pf $C MGB_HOST_PREFETCH=1024
run nt2_t14 $C MGB_HOST_THREADS=14
run nt2_t15 $C MGB_HOST_THREADS=15
run nt2_dring $C MGB_DOWNLOAD_RING=1
run nt2_dring_t14 $C MGB_DOWNLOAD_RING=1 MGB_HOST_THREADS=14
run nt2_c1m MGB_HOST_CHUNK=1048576 MGB_HOST_RING=6
run nt2_c1m_dring MGB_HOST_CHUNK=1048576 MGB_HOST_RING=6 MGB_DOWNLOAD_RING=1
run nt2_small MGB_HOST_CHUNK=65536 MGB_HOST_RING=16
MGB_HOST_CHUNK=262144 MGB_HOST_RING=8 MGB_HOST_STATS=1 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-files 2>&1 >/dev/null | grep "mgb upload" | tail -4
