#!/usr/bin/env python
"""Host-side cost of one engine.run_planes call (enqueue only): the GPU step of the C2 bench is ~0.55 ms, so the host
has to enqueue a step faster than that.  Small planes, so that the GPU never pushes back.

    python tools/host_overhead_probe.py [--profile]
"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aliby_b200 import engine, synth  # noqa: E402
from bench import c2_tree  # noqa: E402

F, H, W, C = 8, 256, 256, 5
dev = torch.device("cuda:0")
plan = engine.compile_tree(c2_tree())
fields = [synth.make_field(100 + i, (H, W), C, 20) for i in range(F)]
px = torch.from_numpy(np.concatenate([f[0] for f in fields])).to(dev)
lab = torch.from_numpy(np.stack([f[1] for f in fields])).to(dev)
n_labels = np.array([int(f[1].max()) for f in fields], dtype=np.int64)
offs = np.arange(F, dtype=np.int64) * (C * H * W)
plane_tile = np.arange(F, dtype=np.int32)
out = torch.empty((int(n_labels.sum()), plan.n_columns), dtype=torch.float64, device=dev)


def step():
    engine.run_planes(plan, lab, plane_tile, n_labels, px, offs, H * W, H * W, W, C, 1, out=out)


for _ in range(20):
    step()
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue: {1e6 * (t1 - t0) / N:.1f} us per call; with the GPU drained: {1e6 * (t2 - t0) / N:.1f} us per call")
if "--profile" in sys.argv:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(N):
        step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
