#!/usr/bin/env python
"""Static SASS instruction count per source line: nvdisasm --print-line-info <cubin> | this script [kernel-substr] [top]."""
import collections
import re
import sys

want = sys.argv[1] if len(sys.argv) > 1 else ""
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_fn, cur_line = None, "?"
counts = collections.defaultdict(collections.Counter)
for ln in sys.stdin:
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = f"{m.group(1).split('/')[-1]}:{m.group(2)}"
        continue
    if cur_fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        counts[cur_fn][cur_line] += 1
for fn, c in counts.items():
    if want not in fn:
        continue
    tot = sum(c.values())
    print(f"== {fn[:90]}: {tot} instr = {tot * 16 / 1024:.1f} KB")
    for line, n in c.most_common(top):
        print(f"   {n:5d}  {line}")
