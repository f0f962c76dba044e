#!/bin/bash
# Four ranks on one socket (gpurun --gpus 4): which host-transport mode should ranks that share a socket use?
mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_n4_build.log 2>&1
: > gpurun_out/r02_n4_transport.log
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 --no-files --no-cpu-baseline 2> gpurun_out/r02_n4_$tag.err | grep '^{' > gpurun_out/r02_n4_$tag.json
  python - "$tag" >> gpurun_out/r02_n4_transport.log <<'PY'
import sys, json
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02_n4_{tag}.json").read()); e = d["e2e"]
    print(f"{tag}: stages.main {e['value']:.0f}x ({e['ms_per_step']:.2f} ms per call, {e['host_threads']} workers per rank, d2h {e['d2h_bytes_per_step']}); device-resident {d['value']:.0f}x")
except Exception as ex:
    print(tag, "failed:", ex)
PY
}
run streaming_ring MGB_BENCH_SHARED_SOCKET_MODE=streaming
run cached_dma MGB_BENCH_SHARED_SOCKET_MODE=cached
cat gpurun_out/r02_n4_transport.log
