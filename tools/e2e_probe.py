"""e2e step time of extract_table for several upload chunk sizes (C2, 8 fields, pinned host arrays)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from aliby_b200 import engine, extract
F = 8
px, lab = bench.make_fields(F, 5000)
px_pin = torch.from_numpy(px).pin_memory(); lab_pin = torch.from_numpy(lab).pin_memory()
masks = [lab_pin[i].numpy() for i in range(F)]; pxh = px_pin.numpy()
tree = bench.c2_tree(); plan = engine.compile_tree(tree)
for mb in (48, 100, 160, 256, 1024):
    for _ in range(2): extract.extract_table(tree, masks, pxh, plan=plan, chunk_bytes=mb << 20)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): tab = extract.extract_table(tree, masks, pxh, plan=plan, chunk_bytes=mb << 20)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"chunk {mb:5d} MB: {dt*1e3:7.2f} ms/step  {tab.values.shape}")
