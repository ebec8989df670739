#!/usr/bin/env python
"""Key metrics per profiled launch of an .ncu-rep:  python tools/ncu_summary.py prof.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "no_instruction", "barrier", "branch_resolving", "math_pipe_throttle",
          "mio_throttle", "lg_throttle", "not_selected", "dispatch_stall", "membar", "drain", "imc_miss", "sleeping", "tex_throttle"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("##", r[hdr.index("Kernel Name")][:100])
    for k in KEYS:
        if k in hdr:
            print(f"   {k:62s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = []
    for s in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if k in hdr:
            st.append((float(r[hdr.index(k)]), s))
    print("   stall cycles per issued instruction:", ", ".join(f"{s} {v:.2f}" for v, s in sorted(st, reverse=True) if v >= 0.05))
