#!/usr/bin/env python
"""Hot source lines of ONE kernel of an ncu report (stall samples by reason).
    python tools/ncu_lines.py prof.ncu-rep stats_warp [top_n] [file_filter]"""
import csv, os, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
filt = sys.argv[4] if len(sys.argv) > 4 else ".cu"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_fun, cur_file, hdr = "", "", None
lines = []
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "Function Name":
        cur_fun = r[1]; continue
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and r and r[0].isdigit() and kern in cur_fun and filt in cur_file:
        d = dict(zip(hdr, r))
        try:
            smp = float(d["# Samples"]); ins = float(d["Instructions Executed"])
        except (ValueError, KeyError):
            continue
        st = {k[6:]: float(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "-") and float(v) > 0}
        lines.append((int(r[0]), r[1].strip()[:70], smp, ins, st))
ts = sum(l[2] for l in lines) or 1; ti = sum(l[3] for l in lines) or 1
print(f"{kern}: {ti:.3e} warp-instr, {ts:.0f} samples in {filt}")
for ln, src, smp, ins, st in sorted(lines, key=lambda l: -l[2])[:top]:
    top3 = ", ".join(f"{k} {v:.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*smp/ts:5.1f}% smp {100*ins/ti:5.1f}% inst  {ln:4d} {src:70s} | {top3}")
