#!/bin/bash
# One measurement cycle on the GPU box: parity tests, then a bench line into gpurun_out/.
# usage: tools/gpu_cycle.sh TAG [pytest -k expression]
TAG=${1:-x}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread -p no:cacheprovider -k "$KEXPR" 2>&1 | tail -3
else
  timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread -p no:cacheprovider 2>&1 | tail -3
fi
timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/bench_r01_$TAG.json 2> gpurun_out/bench_err.log
python - "$TAG" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_r01_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"], 4), round(d["config"]["one_track_at_a_time"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "pcm", round(d["e2e"].get("pcm16", {}).get("value", 0)))
print({k: round(v["ms_per_step"] * 1e3, 1) for k, v in d.get("kernels", {}).items()})
PY
tail -3 gpurun_out/bench_err.log
