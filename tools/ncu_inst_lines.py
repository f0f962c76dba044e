"""Source lines of one kernel ranked by executed warp instructions (needs -lineinfo + --import-source on).
usage: python tools/ncu_inst_lines.py report.ncu-rep kernel_regex [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
cur, hdr = None, None
agg, txt = collections.defaultdict(float), {}
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
    elif hdr and len(r) > 8 and r[2] == "-":
        try:
            v = float(r[hdr["Instructions Executed"]])
        except (ValueError, KeyError):
            v = 0.0
        agg[(cur, int(r[0]))] += v
        txt[(cur, int(r[0]))] = r[1].strip()
tot = sum(agg.values()) or 1.0
print(f"warp instructions {tot:.0f}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1])[:top_n]:
    print(f"{v / tot * 100:5.1f}% {f}:{l} {txt[(f, l)][:110]}")
