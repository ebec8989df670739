#!/usr/bin/env python
"""Blackwell-specific SASS mnemonics per kernel of libaliby_b200.so (nvdisasm of the embedded cubins), for profiles/sass_TAG.txt:

    python tools/sass_evidence.py > profiles/sass_r02.txt

TMA (UTMALDG / UTMAPF = cp.async.bulk.tensor / .prefetch.tensor), mbarrier transactions (SYNCS), tensor memory
(LDTM / STTM = tcgen05.ld / tcgen05.st; UTCATOMSWS = tcgen05.alloc), shared-memory histogram atomics (ATOMS.POPC.INC),
packed DPX min/add (VIADDMNMX / VIMNMX / VIMNMX3 .U16x2), global reductions (REDG)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "aliby_b200", "csrc", "libaliby_b200.so")
pat = re.compile(r"\b(UTMALDG[.\w]*|UTMAPF[.\w]*|SYNCS[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTC\w+[.\w]*|ATOMS[.\w]*|VIADDMNMX[.\w]*|VIMNMX3?[.\w]*|REDG[.\w]*|REDUX[.\w]*|LDGSTS[.\w]*)")
# cuobjdump -sass cuts long functions off at 4096 instructions: extract the cubins and disassemble them with nvdisasm
import glob
import tempfile

tmp = tempfile.mkdtemp(prefix="abx_sass_")
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True, text=True)
out = ""
for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
    out += subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout
fn = None
counts = collections.OrderedDict()
sizes = collections.Counter()
for ln in out.splitlines():
    m = re.search(r"^\s*\.text\.(\S+):", ln) or re.search(r"Function : (\S+)", ln)
    if m:
        dem = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        fn = re.sub(r"\(.*", "", dem).replace("void ", "")
        fn = re.sub(r"(\(anonymous namespace\)|<unnamed>)::", "", fn)
        counts.setdefault(fn, collections.Counter())
        continue
    if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        sizes[fn] += 1
        m = pat.search(ln)
        if m:
            counts[fn][m.group(1)] += 1
print(f"# {os.path.relpath(lib, ROOT)}: nvdisasm of the embedded sm_100a cubins; instruction counts are static (per kernel image)")
for fn, c in counts.items():
    print(f"\n{fn}   [{sizes[fn]} instructions]")
    for k, v in sorted(c.items()):
        print(f"    {v:5d}  {k}")
