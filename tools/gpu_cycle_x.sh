#!/bin/bash
# tracks in flight for the device-resident figure, with the round's final kernels
mkdir -p gpurun_out
python -m matchering_b200.build > gpurun_out/r02_x_build.log 2>&1
for lanes in 4 6 8 12; do
  timeout 200 python bench.py --steps 24 --warmup 6 --lanes $lanes --no-cpu-baseline --no-files > gpurun_out/r02_x_lanes$lanes.json 2> gpurun_out/r02_x_lanes$lanes.err
done
MGB_HOST_THREADS=15 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-files > gpurun_out/r02_x_t15.json 2> gpurun_out/r02_x_t15.err
python - <<'PY'
import json
for n in ("lanes4","lanes6","lanes8","lanes12","t15"):
    try:
        d=json.loads(open(f"gpurun_out/r02_x_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), round(d["ms_per_step"]*1e3,1), "e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"]["host_threads"])
    except Exception as e: print(n, "failed", e)
PY
