"""One line of key metrics per kernel launch of an .ncu-rep (for profiles/*.txt).
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"), ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts")]
print("# " + sys.argv[1])
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
    parts = []
    for key, label in cols:
        if key in ix:
            v = r[ix[key]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            parts.append(f"{label}={v}{units[ix[key]] if label.startswith('dram_r') or label.startswith('dram_w') else ''}")
    print(f"{name:28s} " + " ".join(parts))
