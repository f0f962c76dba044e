"""Top source lines of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, cur_file, hdr = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        src_col = 1
    elif hdr and len(r) > 8 and r[2] == "-":  # per-source-line aggregate
        def f(name):
            try:
                return float(r[hdr[name]])
            except (ValueError, KeyError):
                return 0.0
        stalls = {k[6:]: f(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        lines.append((cur_file, r[0], r[src_col].strip(), f("# Samples"), f("Instructions Executed"), stalls))
ts = sum(l[3] for l in lines) or 1
ti = sum(l[4] for l in lines) or 1
print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
agg = {}
for l in lines:
    for k, v in l[5].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join(f"{k} {v / ts * 100:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for l in sorted(lines, key=lambda l: -l[3])[:top_n]:
    top = max(l[5].items(), key=lambda kv: kv[1])[0] if l[5] else ""
    print(f"{l[3] / ts * 100:5.1f}% smp {l[4] / ti * 100:5.1f}% ins {top:>14} | {l[0]}:{l[1]} {l[2][:100]}")
