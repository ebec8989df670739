#!/usr/bin/env python
"""Per-source-line instruction/sample shares of one kernel from `ncu --page source --csv` output.
   python tools/ncu_lines2.py src.csv kernel_substr [top]"""
import csv, os, sys
path, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cur_fun = cur_file = ""; hdr = None; rows = []
for r in csv.reader(open(path)):
    if len(r) >= 2 and r[0] == "Function Name": cur_fun = r[1]; continue
    if len(r) >= 2 and r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit() and kern in cur_fun:
        d = dict(zip(hdr, r))
        try: rows.append((cur_file, int(r[0]), r[1].strip()[:86], float(d["# Samples"]), float(d["Instructions Executed"])))
        except Exception: pass
ti = sum(x[4] for x in rows); ts = sum(x[3] for x in rows)
print(f"{kern}: {ti:.4g} warp-instr, {ts:.0f} samples")
for f, ln, src, smp, ins in sorted(rows, key=lambda x: -x[4])[:top]:
    print(f"{100*ins/ti:5.1f}% inst {100*smp/ts:5.1f}% smp {f}:{ln} {src}")
