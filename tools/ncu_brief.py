#!/usr/bin/env python
"""Brief per-kernel view of an ncu report: python tools/ncu_brief.py gpurun_out/prof_X.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:60])
    for w in want:
        if w in h:
            print(f"   {w:70s} {r[h.index(w)]}")
    st = []
    for i, n in enumerate(h):
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.05:
                st.append((v, n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    print("   stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)))
