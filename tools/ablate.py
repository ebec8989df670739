"""Stage timing of abx_extract for different feature trees (which feature costs what)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from aliby_b200 import engine, _native as nat

lib = nat.lib()
F = int(os.environ.get("F", 8))
import os
px, lab = bench.make_fields(F, int(os.environ.get("ABX_SEED_BASE", "5000")))
dev = torch.device("cuda")
pxd = torch.from_numpy(px).to(dev); labd = torch.from_numpy(lab).to(dev)
nl = lab.reshape(F, -1).max(axis=1).astype(np.int64)
H, W = bench.FIELD
offs = np.arange(F, dtype=np.int64) * (5 * H * W)

def run(tree, reps=5):
    plan = engine.compile_tree(tree)
    evs = []
    for _ in range(reps):
        e = []
        for _ in range(6):
            h = C.c_void_p(); lib.abx_event_create(C.byref(h)); e.append(h)
        evs.append(e)
    for _ in range(3):
        engine.run_planes(plan, labd, np.arange(F, dtype=np.int32), nl, pxd, offs, H * W, H * W, W, 5, 1)
    for e in evs:
        engine.run_planes(plan, labd, np.arange(F, dtype=np.int32), nl, pxd, offs, H * W, H * W, W, 5, 1, stage_events=e)
    torch.cuda.synchronize()
    ms = np.zeros(5)
    for e in evs:
        for i in range(5):
            t = C.c_float(); lib.abx_event_elapsed_ms(e[i], e[i + 1], C.byref(t)); ms[i] += t.value
    return ms / reps

ch = range(5)
trees = {
    "shape: area+centroid only": {"None": {"None": ["area", "centroid_x", "centroid_y"]}},
    "mean,std,total (5ch)": {c: {"max": ["mean", "std", "total"]} for c in ch},
    "+moment_of_inertia": {c: {"max": ["mean", "std", "total", "moment_of_inertia"]} for c in ch},
    "mean,std,median (5ch)": {c: {"max": ["mean", "std", "median"]} for c in ch},
    "all intensity (5ch)": {c: {"max": bench.INTENSITY_FEATURES} for c in ch},
    "edt axes only (ecc, volume)": {"None": {"None": ["eccentricity", "volume"]}},
    "edt + conical": {"None": {"None": ["eccentricity", "volume", "conical_volume"]}},
    "full C2 tree": bench.c2_tree(),
    "1 channel all intensity": {0: {"max": bench.INTENSITY_FEATURES}},
}
print(f"{F} fields, {int(nl.sum())} objects; stage ms: scan | stats | edt | large | finalize")
for name, tree in trees.items():
    ms = run(tree)
    print(f"{name:32s} " + " ".join(f"{v:7.3f}" for v in ms) + f"   total {ms.sum():7.3f}")
