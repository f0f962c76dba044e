"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in an .ncu-rep, averaged over
its launches, merged into profiles/traffic.json under a workload key (bench.py copies the dominant kernel's figure into
`roofline.traffic`).   python tools/traffic_from_ncu.py report.ncu-rep c2"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0]
    name = {"convolve_fused_kernel": "convolve_kernel"}.get(name, name)  # bench.py's label for both convolution kernels
    total = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        total += float(r[ix[key]]) * scale.get(units[ix[key]], 1.0)
    acc.setdefault(name, []).append(total)
path = os.path.join(ROOT, "profiles", "traffic.json")
table = json.load(open(path)) if os.path.exists(path) else {}
if table and not all(isinstance(v, dict) for v in table.values()):
    table = {"c2_round1": table}  # round 1 kept a flat table for config 2
table[workload] = {k: round(sum(v) / len(v)) for k, v in acc.items()}
table[workload]["_source"] = os.path.basename(rep)
json.dump(table, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(table[workload], indent=1))
