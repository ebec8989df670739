"""In-tree build of ``libaliby_b200.so`` (sm_100a) with nvcc.

    python -m aliby_b200.build [--force] [--verbose]

The shared library is written next to the sources (``aliby_b200/csrc/``) so that it
travels with a snapshot of the repository; it is git-ignored.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(CSRC, "libaliby_b200.so")
SOURCES = ["abi.cu", "label_scan.cu", "object_warp.cu", "object_sweep.cu", "object_edt.cu", "zreduce.cu", "object_stats.cu", "object_float.cu", "object_pair.cu", "background.cu", "shape_edt.cu", "finalize.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "warp_common.cuh"), os.path.join(CSRC, "tma.cuh"), os.path.join(CSRC, "edt_phases.cuh"), os.path.join(INCLUDE, "aliby_b200.h")]

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "-fmad=false",  # fp64 finalisation must round like NumPy: no FMA contraction
    "-Xcompiler",
    "-fPIC",
    "-I",
    INCLUDE,
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libaliby_b200.so cannot be built")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = find_nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []
    for flag in ("ABX_NO_PREFETCH",):  # experiment switches: ABX_NO_PREFETCH=1 python -m aliby_b200.build --force
        if os.environ.get(flag):
            extra.append(f"-D{flag}")
    objs = []
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + HEADERS):
            jobs.append([nvcc, *NVCC_FLAGS, *extra, "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return r.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for log in ex.map(run, jobs):
            if verbose and log:
                print(log, file=sys.stderr)
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
