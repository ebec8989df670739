#!/usr/bin/env python
"""Rank CUDA source lines of an ncu `--page source --csv --print-source cuda,sass` export.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python profiles/ncu_source_summary.py src.csv [top_n]

Prints, per source line, its share of warp-stall samples and of executed warp instructions.
"""
import csv
import os
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = "?"
lines = []
hdr = None
for r in csv.reader(open(path)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples")
        i_i = hdr.index("Instructions Executed")
        i_t = hdr.index("Thread Instructions Executed")
        continue
    if hdr and r and r[0].isdigit():
        try:
            lines.append((cur_file, int(r[0]), r[1].strip()[:100], float(r[i_s]), float(r[i_i]), float(r[i_t])))
        except ValueError:
            pass
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print(f"total: {tot_i:.3e} warp-instructions, {tot_s:.0f} samples")
for f, ln, src, s, i, t in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% inst  thr/inst {t/max(i,1):4.1f}  {f}:{ln:<4} {src}")
