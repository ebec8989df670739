#!/usr/bin/env python
"""C3 (yeast time-lapse, 40 tiles of 96^2 fused out of 1200^2 frames, 5 channels): resident time of one call per
time point (what the reference pipeline does) and of one call per batch of time points (the same planes, batched)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aliby_b200 import engine, synth  # noqa: E402

dev = torch.device("cuda:0")
T, C, TILE, NT = 25, 5, 96, 40
frames, centres, labels = synth.make_trap_position(1003, n_tp=T, n_channels=C, frame=(1200, 1200), n_tiles=NT, tile_size=TILE)
H, W = frames.shape[-2:]
INT = ["mean", "median", "std", "max5px_median", "imBackground"]
tree = {"None": {"None": ["area", "volume", "eccentricity", "centroid_x", "centroid_y"]}}
for ch in range(C):
    tree[ch] = {"max": list(INT)}
plan = engine.compile_tree(tree)
fr = torch.from_numpy(frames).to(dev)                       # (T, C, 1, H, W)
lab = torch.from_numpy(labels.reshape(T * NT, TILE, TILE)).to(dev)
n_labels = labels.reshape(T * NT, -1).max(axis=1).astype(np.int64)
org = centres - TILE // 2
tile_off1 = (org[:, 0] * W + org[:, 1]).astype(np.int64)     # inside one frame
algo = T * NT * TILE * TILE * (C * 2 + 2)


def per_tp():
    for t in range(T):
        engine.run_planes(plan, lab[t * NT:(t + 1) * NT], np.arange(NT, dtype=np.int32), n_labels[t * NT:(t + 1) * NT], fr[t],
                          tile_off1, H * W, H * W, W, C, 1)


def batched():
    offs = (np.arange(T, dtype=np.int64)[:, None] * (C * H * W) + tile_off1[None, :]).reshape(-1)
    engine.run_planes(plan, lab, np.arange(T * NT, dtype=np.int32), n_labels, fr, offs, H * W, H * W, W, C, 1)


for name, fn in (("one call per time point", per_tp), ("one call per 25 time points", batched)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {1e3 * dt / T:.3f} ms per time point, {int(n_labels.sum())} cells in {T} time points, "
          f"{algo / 1e9 / dt:.1f} GB/s of the {algo / T / 1e6:.2f} MB per time point the tiles hold")

# stage breakdown of one per-time-point call
import ctypes as C_  # noqa: E402

from aliby_b200 import _native as nat  # noqa: E402

lib = nat.lib()
evs = []
for _ in range(6):
    h = C_.c_void_p()
    nat.check(lib.abx_event_create(C_.byref(h)), "ev")
    evs.append(h)
acc = np.zeros(5)
for t in range(T):
    engine.run_planes(plan, lab[t * NT:(t + 1) * NT], np.arange(NT, dtype=np.int32), n_labels[t * NT:(t + 1) * NT], fr[t],
                      tile_off1, H * W, H * W, W, C, 1, stage_events=evs)
    torch.cuda.synchronize()
    for i in range(5):
        ms = C_.c_float()
        nat.check(lib.abx_event_elapsed_ms(evs[i], evs[i + 1], C_.byref(ms)), "el")
        acc[i] += ms.value
print("per-time-point stages ms (scan, stats, edt, large, finalize):", [round(x / T, 3) for x in acc])
