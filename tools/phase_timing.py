import ctypes as C, sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from aliby_b200 import engine, _native as nat
lib = nat.lib()
F=4
px, lab = bench.make_fields(F, 5000)
plan = engine.compile_tree(bench.c2_tree())
dev = torch.device('cuda')
pxd = torch.from_numpy(px).to(dev); labd = torch.from_numpy(lab).to(dev)
nl = lab.reshape(F,-1).max(axis=1).astype(np.int64)
H,W = bench.FIELD
offs = np.arange(F, dtype=np.int64)*(5*H*W)
def step(): return engine.run_planes(plan, labd, np.arange(F,dtype=np.int32), nl, pxd, offs, H*W, H*W, W, 5, 1)
for _ in range(3): step()
out = (C.c_ulonglong*8)()
lib.abx_debug_phase_cycles(out, 1)
step(); torch.cuda.synchronize()
lib.abx_debug_phase_cycles(out, 1)
v = np.array(list(out), dtype=float)
names = ['M mask','S pass1','S hist','S find','S refine+sums','E total','E g+colpass(in E)','E topdetect(in E)']
tot = v[0]+v[1]+v[2]+v[3]+v[4]+v[5]
for n_,x in zip(names, v): print(f"{n_:22s} {x/1e6:10.1f} Mcycles  {100*x/tot:5.1f}%")
print('objects', nl.sum(), 'cycles/object', tot/nl.sum())
