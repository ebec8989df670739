#!/usr/bin/env python
"""Stage times of one C4 field (5 channels x 16 z x 2048^2 uint16, ~1500 cells, Z-max on four channels, Z-add on one)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aliby_b200 import _native as nat  # noqa: E402
from aliby_b200 import engine, synth  # noqa: E402

dev = torch.device("cuda:0")
lib = nat.lib()
H = W = 2048
CH, Z = 5, int(os.environ.get("Z", "16"))
INT = ["mean", "std", "median", "total", "max2p5pc", "max5px_median"]
tree = {"None": {"None": ["area", "centroid_x", "centroid_y", "eccentricity", "volume"]}}
for ch in range(CH):
    tree[ch] = {("add" if ch == 4 and os.environ.get("ADD", "1") == "1" else "max"): list(INT)}
plan = engine.compile_tree(tree)
rng = np.random.default_rng(4)
labels_np = synth.ellipse_labels(rng, (H, W), 1500)
labels = torch.from_numpy(labels_np[None]).to(dev)
pix = torch.randint(100, 5000, (1, CH, Z, H, W), dtype=torch.int32, device=dev).to(torch.uint16)
n_labels = np.array([int(labels_np.max())], dtype=np.int64)
evs = []
for _ in range(6):
    h = C.c_void_p()
    nat.check(lib.abx_event_create(C.byref(h)), "ev")
    evs.append(h)
best = None
for _ in range(4):
    engine.run_planes(plan, labels, np.zeros(1, np.int32), n_labels, pix, np.zeros(1, np.int64), Z * H * W, H * W, W, CH, Z,
                      stage_events=evs)
    torch.cuda.synchronize()
    st = []
    for i in range(5):
        ms = C.c_float()
        nat.check(lib.abx_event_elapsed_ms(evs[i], evs[i + 1], C.byref(ms)), "el")
        st.append(ms.value)
    if best is None or sum(st) < sum(best):
        best = st
algo = pix.numel() * 2 + labels.numel() * 2
print("stages ms (scan, stats, edt, large, finalize):", [round(x, 3) for x in best], "total", round(sum(best), 3),
      f"-> {algo / 1e9 / (sum(best) / 1e3):.0f} GB/s of {algo / 1e6:.0f} MB algorithmic")
