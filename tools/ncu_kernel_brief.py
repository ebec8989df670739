#!/usr/bin/env python
"""Key metrics and stall breakdown per kernel of an ncu report.   python tools/ncu_kernel_brief.py prof.ncu-rep"""
import csv, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:70])
    for w in want:
        if w in d: print(f"  {w:80s} {d[w]} {units[hdr.index(w)]}")
    st = {}
    for h in hdr:
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(.*)\.ratio", h)
        if m and d[h] not in ("", "n/a"):
            try: st[m.group(1) or m.group(2)] = float(d[h])
            except ValueError: pass
    print("  stalls/issue:", ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
