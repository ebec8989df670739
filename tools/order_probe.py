"""Does the spatial order of the label ids matter?  Same fields, ids relabelled in raster order of the first pixel."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from aliby_b200 import engine, _native as nat
lib = nat.lib()
F = 8
px, lab = bench.make_fields(F, 5000)
def raster(lab):
    out = np.zeros_like(lab)
    for f in range(lab.shape[0]):
        flat = lab[f].ravel()
        ids, first = np.unique(flat, return_index=True)
        keep = ids > 0
        order = ids[keep][np.argsort(first[keep])]
        lut = np.zeros(int(flat.max()) + 1, np.uint16)
        lut[order] = np.arange(1, len(order) + 1)
        out[f] = lut[lab[f]]
    return out
dev = torch.device("cuda")
pxd = torch.from_numpy(px).to(dev)
H, W = bench.FIELD
offs = np.arange(F, dtype=np.int64) * (5 * H * W)
plan = engine.compile_tree(bench.c2_tree())
for name, L in (("random ids", lab), ("raster ids", raster(lab))):
    labd = torch.from_numpy(L).to(dev)
    nl = L.reshape(F, -1).max(axis=1).astype(np.int64)
    evs = []
    for _ in range(5):
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
    print(name, int(nl.sum()), "objects; scan|stats|edt|large|finalize ms:", np.round(ms / 5, 3), "total", round(ms.sum() / 5, 3))
