#!/usr/bin/env python
"""Find the object that makes the shape kernel slow: EDT stage time per field, then bisection over label ranges."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from aliby_b200 import _native as nat  # noqa: E402
from aliby_b200 import engine  # noqa: E402

dev = torch.device("cuda:0")
lib = nat.lib()
plan = engine.compile_tree({"None": {"None": ["eccentricity", "volume", "conical_volume"]}})
F = 8
px, lab = bench.make_fields(F, int(os.environ.get("ABX_SEED_BASE", "5100")))
H, W = lab.shape[1:]


def edt_ms(labels_np):
    labels = torch.from_numpy(labels_np[None]).to(dev)
    n_labels = np.array([int(labels_np.max())], dtype=np.int64)
    pix = torch.zeros((1, 1, 1, H, W), dtype=torch.uint16, device=dev)
    evs = []
    for _ in range(6):
        h = C.c_void_p()
        nat.check(lib.abx_event_create(C.byref(h)), "ev")
        evs.append(h)
    best = 1e9
    for _ in range(3):
        engine.run_planes(plan, labels, np.zeros(1, np.int32), n_labels, pix, np.zeros(1, np.int64), H * W, H * W, W, 1, 1,
                          stage_events=evs)
        torch.cuda.synchronize()
        ms = C.c_float()
        nat.check(lib.abx_event_elapsed_ms(evs[2], evs[3], C.byref(ms)), "el")
        best = min(best, ms.value)
    return best


times = [edt_ms(lab[f]) for f in range(F)]
print("EDT stage ms per field:", [round(t, 3) for t in times])
f = int(np.argmax(times))
L = lab[f]
lo, hi = 1, int(L.max()) + 1
while hi - lo > 1:
    mid = (lo + hi) // 2
    sub = np.where((L >= lo) & (L < mid), L, 0).astype(np.uint16)
    t = edt_ms(sub)
    print(f"labels [{lo},{mid}): {t:.3f} ms")
    if t > 0.5 * max(times):
        hi = mid
    else:
        lo = mid
obj = lo
ys, xs = np.nonzero(L == obj)
print("slow object", obj, "n", len(ys), "bbox rows", ys.min(), ys.max(), "cols", xs.min(), xs.max(), "cmin&7", xs.min() & 7)
m = (L[ys.min():ys.max() + 1, xs.min():xs.max() + 1] == obj)
for row in m[:70]:
    print("".join("#" if v else "." for v in row))
