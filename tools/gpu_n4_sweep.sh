mkdir -p gpurun_out
for cfg in "65536 16 16" "65536 16 8" "16384 32 16" "262144 4 16"; do
  set -- $cfg
  MGB_HOST_CHUNK=$1 MGB_HOST_RING=$2 MGB_HOST_THREADS=$3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 12 --warmup 4 --no-files 2> gpurun_out/r02_n4_sweep.err | grep '^{' | python -c "
import sys, json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('chunk $1 ring $2 threads $3: seam', round(e['value']), 'x', round(e['ms_per_step'],2), 'ms; value', round(d['value']))
" >> gpurun_out/r02_n4_sweep.log
done
cat gpurun_out/r02_n4_sweep.log
