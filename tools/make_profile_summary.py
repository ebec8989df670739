#!/usr/bin/env python
"""Turn the scratch outputs of `tools/gpu_profile.sh TAG` (gpurun_out/) into the tracked evidence under profiles/:

    python tools/make_profile_summary.py TAG

writes profiles/TAG_summary.md (launch shares, ncu key metrics, stall reasons, hottest source lines), copies the
launch list to profiles/TAG_launches.csv and refreshes profiles/traffic.json (dram bytes per launch of each hot
kernel, read by bench.py for `roofline.traffic`).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
rep = os.path.join(OUT, f"prof_{tag}.ncu-rep")
launches = os.path.join(OUT, f"launches_{tag}.csv")
bench = os.path.join(OUT, f"bench_{tag}.json")
STAGE_OF = {"label_scan_kernel": "label_scan", "object_sweep": "object_stats", "object_stats_warp": "object_stats",
            "object_edt_grid": "object_edt", "finalize_kernel": "finalize"}

md = [f"# ncu summary {tag}", "",
      "Produced by `tools/gpu_profile.sh` on a B200 (sm_100a) and `tools/make_profile_summary.py`; command profiled: "
      "`python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs` (C2: 32 fields of 5ch x 2160^2 per step; "
      "the step's two chains run in line under the stage events / under ncu's serialisation).", ""]

if os.path.exists(bench):
    try:
        line = [ln for ln in open(bench) if ln.startswith("{")][-1]
        b = json.loads(line)
        md += ["## bench line of the same build (not under ncu)", "", "```json", json.dumps(b, indent=1), "```", ""]
    except Exception as e:  # noqa: BLE001
        md += [f"(bench line unreadable: {e})", ""]


def short(name):
    n = name.split("(")[0]
    n = n.replace("void ", "").replace("<unnamed>::", "")
    return n.split("<")[0].strip()


if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        k = short(r[4])
        per.setdefault(k, []).append(float(r[-1]))
    tot = sum(sum(v) for v in per.values())
    md += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised)", "",
           f"{len(rows)} launches captured. Shares of the summed kernel time:", "",
           "| kernel | launches | mean us | share |", "|---|---|---|---|"]
    for k, v in per.items():
        md.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f} % |")
    md.append("")
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum [ns]"])
        for r in rows:
            w.writerow([r[0], short(r[4]), r[8], r[7], r[-1]])

traffic = {}
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "launch__shared_mem_per_block_static", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
    stalls = ["long_scoreboard", "short_scoreboard", "wait", "no_instruction", "barrier", "branch_resolving",
              "math_pipe_throttle", "mio_throttle", "lg_throttle", "not_selected", "dispatch_stall", "membar", "drain",
              "imc_miss", "sleeping", "tex_throttle"]
    md += ["## `ncu --set full --clock-control none --import-source on` (one launch of each hot kernel)", ""]

    def to_bytes(val, unit):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return float(val) * mult.get(unit, 1)

    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        md += [f"### `{name}`", "", "| metric | value |", "|---|---|"]
        for k in keys:
            if k in hdr:
                md.append(f"| {k} | {r[hdr.index(k)]} {units[hdr.index(k)]} |")
        st = []
        for s_ in stalls:
            k = f"smsp__average_warps_issue_stalled_{s_}_per_issue_active.ratio"
            if k in hdr:
                st.append((float(r[hdr.index(k)]), s_))
        md.append("| stall cycles per issued instruction | " + ", ".join(f"{s_} {v:.2f}" for v, s_ in sorted(st, reverse=True) if v >= 0.05) + " |")
        md.append("")
        try:
            i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            stage = STAGE_OF.get(name)
            if stage:
                traffic[stage] = to_bytes(r[i_r], units[i_r]) + to_bytes(r[i_w], units[i_w])
        except ValueError:
            pass
    # hottest source lines per kernel
    kerns = []
    for r in rows[2:]:
        k = short(r[hdr.index("Kernel Name")])
        if k not in kerns:
            kerns.append(k)
    for kern in kerns:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                              f"regex:{kern}"], capture_output=True, text=True).stdout
        tmp = os.path.join(OUT, f"src_{tag}_{kern}.csv")
        open(tmp, "w").write(src)
        top = subprocess.run([sys.executable, os.path.join(PROF, "ncu_source_summary.py"), tmp, "14"], capture_output=True,
                             text=True).stdout
        md += [f"### hottest source lines, `{kern}` (share of warp-stall samples / of executed instructions)", "", "```",
               top.rstrip(), "```", ""]

if traffic:
    traffic["_source"] = f"profiles/{tag}_summary.md (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full launch each)"
    json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
open(os.path.join(PROF, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
print("wrote", os.path.join(PROF, f"{tag}_summary.md"), "traffic:", traffic)
