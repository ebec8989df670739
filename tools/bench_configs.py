"""Per-configuration measurements of ``bench.py`` (BASELINE.json configs C1, C2 one field per call, C3, C4).

Every configuration is measured twice:

* ``device_ms``  inputs resident in HBM, CUDA events around each call on the launch stream, an L2 flush (a 256 MB
  write) between the timed calls because these inputs are smaller than the 126 MB L2;
* ``e2e_ms``     the public API with HOST arrays (``extract.extract_table``; the fused tiler for C3): host-to-device
  copies of the inputs and the device-to-host copy of the table inside the timed region.

``image_gbs`` = algorithmic bytes (one read of every extracted pixel, one read of the label planes,
one write of the table) / ``device_ms``.  ``reference_configs`` times the reference's CPU algorithm (``oracle.port``)
on C1 and on one C3 time point IN FULL (these two are small enough not to be sampled).
"""

from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

STAGES = ["label_scan", "object_stats", "object_edt", "large_objects", "finalize"]

C1_TREE = {
    "None": {"None": ["area", "centroid_x", "centroid_y", "eccentricity", "volume"]},
    0: {"max": ["mean", "std", "median", "total", "max2p5pc", "max5px_median", "max"]},
    1: {"max": ["mean", "std", "median", "total", "max2p5pc", "max5px_median", "max"]},
}
C3_INTENSITY = ["mean", "median", "std", "imBackground", "max5px_median"]  # global_settings.py:45-53 fluorescence_functions
C3_SHAPE = ["area", "volume", "eccentricity", "centroid_x", "centroid_y"]  # global_settings.py:37-43 outline_functions
C4_INTENSITY = ["mean", "std", "median", "total", "max2p5pc", "max5px_median"]


def c3_tree(n_channels=5):
    tree = {"None": {"None": list(C3_SHAPE)}}
    for ch in range(n_channels):
        tree[ch] = {"max": list(C3_INTENSITY)}
    return tree


def c4_tree(n_channels=5):
    tree = {"None": {"None": ["area", "centroid_x", "centroid_y", "eccentricity", "volume"]}}
    for ch in range(n_channels):
        tree[ch] = {("add" if ch == n_channels - 1 else "max"): list(C4_INTENSITY)}
    return tree


class _Events:
    def __init__(self, lib, nat):
        self.lib, self.nat = lib, nat
        self.h = []
        for _ in range(len(STAGES) + 1):
            e = C.c_void_p()
            nat.check(lib.abx_event_create(C.byref(e)), "abx_event_create")
            self.h.append(e)

    def stage_ms(self):
        out = []
        for i in range(len(STAGES)):
            ms = C.c_float()
            self.nat.check(self.lib.abx_event_elapsed_ms(self.h[i], self.h[i + 1], C.byref(ms)), "abx_event_elapsed_ms")
            out.append(ms.value)
        return out

    def close(self):
        for e in self.h:
            self.lib.abx_event_destroy(e)


def _timed_calls(torch, fn, iters, flush):
    """Per-call device time (ms): L2 flush, event, call, event — the flush stays outside the bracket."""
    fn(None)
    fn(None)
    torch.cuda.synchronize()
    pairs = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(None)
        b.record()
        pairs.append((a, b))
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in pairs]))


def _stage_profile(torch, fn, ev, flush, iters=3):
    acc = np.zeros(len(STAGES))
    for _ in range(iters):
        flush.zero_()
        fn([e for e in ev.h])
        torch.cuda.synchronize()
        acc += np.asarray(ev.stage_ms())
    return {n: float(v / iters) for n, v in zip(STAGES, acc)}


def _e2e(torch, fn, iters):
    fn()
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / iters


def _entry(workload, n_objects, n_features, algo_bytes, device_ms, e2e_ms, stage_ms, peak, extra=None):
    d = {
        "workload": workload,
        "objects": int(n_objects),
        "features_per_object": int(n_features),
        "algorithmic_mb": algo_bytes / 1e6,
        "device_ms": device_ms,
        "e2e_ms": e2e_ms,
        "object_features_per_s": n_objects * n_features / (device_ms / 1e3),
        "e2e_object_features_per_s": n_objects * n_features / (e2e_ms / 1e3) if e2e_ms else None,
        "image_gbs": algo_bytes / 1e9 / (device_ms / 1e3),
        "image_gbs_frac_of_hbm_peak": algo_bytes / 1e9 / (device_ms / 1e3) / peak,
        "stage_ms": stage_ms,
        "l2_policy": "256 MB write between the timed calls (inputs smaller than L2)",
    }
    if extra:
        d.update(extra)
    return d


def measure_configs(device, peak_gbs, iters=10, which=("C1", "C2_single_field", "C2_cp_measure", "C3", "C4")):
    import torch

    from aliby_b200 import _native as nat
    from aliby_b200 import engine, extract, synth
    from aliby_b200.tile import TileView

    lib = nat.lib()
    ev = _Events(lib, nat)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    out = {}

    def dense_case(name, workload, pixels, labels, tree, cp_kwargs=None):
        """One field: pixels (1, C, Z, Y, X), labels (Y, X)."""
        plan = engine.compile_tree(tree, cp_kwargs)
        assert plan.error is None, plan.error
        _, C_, Z_, Y, X = pixels.shape
        lab_dev = torch.from_numpy(labels[None]).to(device)
        px_dev = torch.from_numpy(pixels).to(device)
        n_labels = np.array([int(labels.max())], dtype=np.int64)
        table = torch.empty((int(n_labels[0]), plan.n_columns), dtype=torch.float64, device=device)

        def call(events):
            engine.run_planes(plan, lab_dev, np.zeros(1, np.int32), n_labels, px_dev, np.zeros(1, np.int64), Z_ * Y * X, Y * X, X,
                              C_, Z_, out=table, stage_events=events)

        dms = _timed_calls(torch, call, iters, flush)
        st = _stage_profile(torch, call, ev, flush)
        px_pin = torch.from_numpy(pixels).pin_memory().numpy()
        lab_pin = torch.from_numpy(labels).pin_memory().numpy()
        ems = _e2e(torch, lambda: extract.extract_table(tree, lab_pin, px_pin, device=device, plan=plan), max(3, iters // 2))
        algo = pixels.nbytes + labels.nbytes + int(n_labels[0]) * plan.n_columns * 8
        n_feat = sum(len(c) for c in plan.inst_cols)  # a dict-valued (cp_measure) instruction counts once per key
        extra = None
        if int(n_labels[0]) <= 8192:
            # the same call replayed from a CUDA graph (what the drop-in functions do from the third call of a shape on):
            # an eager call of this size is bound by the host's launch calls, and its figure follows the box's CPU
            cap = max(16, -(-2 * int(n_labels[0]) // 8) * 8)
            g = engine.GraphedExtract(plan, 1, Y, X, np.zeros(1, np.int32), tuple(pixels.shape), px_dev.dtype, np.zeros(1, np.int64),
                                      Z_ * Y * X, Y * X, X, C_, Z_, cap, device)
            g.run(lab_dev, px_dev)
            extra = {"replay_only_device_ms": _timed_calls(torch, lambda events: g.run(), iters, flush)}
            del g
        out[name] = _entry(workload, n_labels[0], n_feat, algo, dms, ems, st, peak_gbs, extra)
        del lab_dev, px_dev, table

    if "C1" in which:
        px, lab = synth.make_field(synth.CONFIG_SEEDS["C1"], (1080, 1080), 2, 300)
        dense_case("C1", "2 channels x 1080^2 uint16, ~300 nuclei, intensity + sizeshape-like tree, one field per call", px, lab,
                   C1_TREE)
    if "C2_single_field" in which:
        import bench

        px, lab = synth.make_field(synth.CONFIG_SEEDS["C2"], (2160, 2160), 5, 2000)
        dense_case("C2_single_field", "5 channels x 2160^2 uint16, ~2k cells, full cell-function set, ONE field per call "
                   "(what a pipeline step does)", px, lab, bench.c2_tree())
    if "C2_cp_measure" in which:
        # what the reference's stock builder asks for (pipe_builder.py:19-43,115-120), restricted to the cp_measure features
        # with a kernel: sizeshape on the masks, intensity per channel, and the two-image features of every channel pair
        from itertools import combinations

        px, lab = synth.make_field(synth.CONFIG_SEEDS["C2"] + 1, (2160, 2160), 5, 2000)
        kw = {"intensity": {"edge_measurements": False}}
        tree = {"None": {"None": ["sizeshape"]}}
        for ch in range(5):
            tree[ch] = {"max": ["intensity"]}
        dense_case("C2_cp_measure", "C2 field, cp_measure features with a kernel: sizeshape (15 keys) + intensity on 5 channels "
                   "(16 keys each, edge features off), one field per call", px, lab, tree, kw)
        multi = {pair: {"None": {"max": ["pearson", "manders_fold", "rwc"]}} for pair in combinations(range(5), r=2)}
        dense_case("C2_extractmulti", "C2 field, extractmulti tree of the stock builder without costes: pearson, manders_fold, "
                   "rwc on all 10 channel pairs (5 keys each), one field per call", px, lab, multi)
    if "C4" in which:
        rng = np.random.default_rng(synth.CONFIG_SEEDS["C4"])
        lab = synth.ellipse_labels(rng, (2048, 2048), 1500)
        px = rng.integers(100, 5000, size=(1, 5, 16, 2048, 2048), dtype=np.uint16)
        dense_case("C4", "5 channels x 16 z x 2048^2 uint16, ~1.5k cells, Z-max on four channels and Z-add on one, one field "
                   "per call", px, lab, c4_tree())
        del px
    if "C3" in which:
        T, NC, TILE, NT = 25, 5, 96, 40
        frames, centres, labels = synth.make_trap_position(synth.CONFIG_SEEDS["C3"], n_tp=T, n_channels=NC, frame=(1200, 1200),
                                                          n_tiles=NT, tile_size=TILE)
        H, W = frames.shape[-2:]
        tree = c3_tree(NC)
        plan = engine.compile_tree(tree)
        fr = torch.from_numpy(frames).to(device)
        lab = torch.from_numpy(labels.reshape(T * NT, TILE, TILE)).to(device)
        n_labels = labels.reshape(T * NT, -1).max(axis=1).astype(np.int64)
        org = centres - TILE // 2
        off1 = (org[:, 0] * W + org[:, 1]).astype(np.int64)
        tiles = np.arange(NT, dtype=np.int32)
        n_obj = int(n_labels.sum())

        def per_tp(events):
            for t in range(T):
                engine.run_planes(plan, lab[t * NT:(t + 1) * NT], tiles, n_labels[t * NT:(t + 1) * NT], fr[t], off1, H * W, H * W,
                                  W, NC, 1, stage_events=events if t == T - 1 else None)

        offs_all = (np.arange(T, dtype=np.int64)[:, None] * (NC * H * W) + off1[None, :]).reshape(-1)
        tiles_all = np.arange(T * NT, dtype=np.int32)

        def batched(events):
            engine.run_planes(plan, lab, tiles_all, n_labels, fr, offs_all, H * W, H * W, W, NC, 1, stage_events=events)

        algo_tp = NT * TILE * TILE * (NC * 2 + 2) + (n_obj / T) * plan.n_columns * 8
        d1 = _timed_calls(torch, per_tp, max(3, iters // 2), flush) / T
        s1 = _stage_profile(torch, per_tp, ev, flush)
        d2 = _timed_calls(torch, batched, max(3, iters // 2), flush) / T
        s2 = {k: v / T for k, v in _stage_profile(torch, batched, ev, flush).items()}
        # e2e per time point through the fused tiler: host frame in, table out
        frames_pin = torch.from_numpy(frames).pin_memory().numpy()
        masks = [[labels[t, i] for i in range(NT)] for t in range(T)]

        def e2e_tp():
            for t in range(T):
                view = TileView(frames_pin[t], org, TILE)
                extract.extract_table(tree, masks[t], view, device=device, plan=plan)

        e1 = _e2e(torch, e2e_tp, 3) / T
        note = {"time_points": T, "tiles": NT, "tile": TILE, "frame": [H, W],
                "algorithmic_mb_whole_frames": NC * H * W * 2 / 1e6}
        out["C3_per_timepoint"] = _entry(
            "yeast position: 5 channels, 40 tiles of 96^2 fused out of 1200^2 frames, cell + per-tile background metrics, "
            "one call per time point (numbers per time point)", n_obj / T, len(plan.instructions), algo_tp, d1, e1, s1,
            peak_gbs, note)
        # the same per-time-point call captured once in a CUDA graph and replayed (engine.GraphedExtract: what the drop-in
        # does from the third call with identical shapes on); every plane owns `cap` table rows
        g = engine.GraphedExtract(plan, NT, TILE, TILE, tiles, (NC, 1, H, W), torch.uint16, off1, H * W, H * W, W, NC, 1, 16,
                                  device)

        def replay(events):
            for t in range(T):
                g.run(lab[t * NT:(t + 1) * NT], fr[t])  # device-to-device refresh of the static inputs + replay

        d3 = _timed_calls(torch, replay, max(3, iters // 2), flush) / T

        def replay_only(events):
            for t in range(T):
                g.run()

        d3b = _timed_calls(torch, replay_only, max(3, iters // 2), flush) / T

        def dropin_tp():  # process_tree_masks on the fused tile view, host frame in, item lists out (graph path inside)
            for t in range(T):
                extract.process_tree_masks(tree, masks[t], TileView(frames_pin[t], org, TILE), extract.extract_tree)

        e3 = _e2e(torch, dropin_tp, 3) / T
        out["C3_per_timepoint_graph"] = _entry(
            "one call per time point replayed from a CUDA graph (static input buffers refreshed by device copies; "
            "e2e_ms: process_tree_masks + extract_tree on the fused tile view, host frame in, Python item lists out)",
            n_obj / T, len(plan.instructions), algo_tp, d3, e3, s1, peak_gbs, {**note, "replay_only_device_ms": d3b})
        del g
        out["C3_batched"] = _entry(
            "the same 25 time points in ONE call (tile offsets into the (T, C, Z, Y, X) stack; numbers per time point)",
            n_obj / T, len(plan.instructions), algo_tp, d2, None, s2, peak_gbs, note)
    ev.close()
    return out


def dropin_call_ms(device, iters=5):
    """One C2 field per call through the reference-facing functions exactly as pipe_core.py:217 invokes them:
    ``process_tree_masks(tree, masks, pixels, extract_tree)`` + ``format_extraction`` with PAGEABLE NumPy inputs;
    the Arrow table is built inside the timed region."""
    import torch

    import bench
    from aliby_b200 import extract, synth

    px, lab = synth.make_field(synth.CONFIG_SEEDS["C2"] + 7, (2160, 2160), 5, 2000)
    tree = bench.c2_tree()

    def call():
        res = extract.process_tree_masks(tree, lab, px, extract.extract_tree)
        return extract.format_extraction(res)

    with torch.cuda.device(device):
        for _ in range(4):  # (the third call with the same shapes captures the CUDA graph that later calls replay)
            call()
        t0 = time.perf_counter()
        for _ in range(iters):
            table = call()
        dt = (time.perf_counter() - t0) / iters
    return {
        "ms_per_call": 1e3 * dt,
        "rows": table.num_rows,
        "columns": table.num_columns,
        "object_features_per_s": table.num_rows * (table.num_columns - 2) / dt,
        "api": "process_tree_masks(tree, masks, pixels, extract_tree) + format_extraction, one C2 field per call, pageable "
               "NumPy inputs, pyarrow table built inside the timed region (pipe_core.py:217, extract.py:520-599)",
    }


# ------------------------------------------------------------------------------------------ reference arm
def _port_job(args):
    from oracle import port

    lab, px, k, inst = args
    ch, red, metric = inst
    plane = lab == k
    img = None
    if ch != "None":
        img = port.project_z(px[ch], port.Z_REDUCERS[red])
    if metric in ("imBackground", "background_max5"):
        fn = port.t_background_median if metric == "imBackground" else port.t_background_max5
        return float(fn(lab, img))
    return float(port.CELL_METRICS[metric](plane, img))


def reference_configs(cores):
    """oracle.port (the reference's algorithm, restated) on C1 and on one C3 time point, IN FULL, joblib over all cores."""
    from joblib import Parallel, delayed

    from aliby_b200 import synth
    from oracle import port

    out = {}
    with Parallel(n_jobs=cores, backend="loky") as par:
        px, lab = synth.make_field(synth.CONFIG_SEEDS["C1"], (1080, 1080), 2, 300)
        tree = {k: {r: [m for m in ms if m != "max"] for r, ms in v.items()} for k, v in C1_TREE.items()}  # registry metrics only
        insts = port.tree_instructions(tree)
        jobs = [(lab, px[0], k, inst) for k in range(1, int(lab.max()) + 1) for inst in insts]
        par(delayed(_port_job)(j) for j in jobs[: 2 * cores])  # start the workers
        t0 = time.perf_counter()
        par(delayed(_port_job)(j) for j in jobs)
        dt = time.perf_counter() - t0
        out["C1"] = {"items": len(jobs), "seconds": dt, "object_features_per_s": len(jobs) / dt, "cores": cores,
                     "sample": "every object x every registry instruction of the C1 field (in full)"}
        frames, centres, labels = synth.make_trap_position(synth.CONFIG_SEEDS["C3"], n_tp=1, n_channels=5, frame=(1200, 1200),
                                                          n_tiles=40, tile_size=96)
        insts = port.tree_instructions(c3_tree(5))
        crops = port.crop_tiles(frames[0], centres, (96, 96))  # (tiles, C, Z, 96, 96): the reference's materialised tiles
        jobs = [(labels[0, t], crops[t], k, inst) for t in range(40) for k in range(1, int(labels[0, t].max()) + 1) for inst in insts]
        t0 = time.perf_counter()
        par(delayed(_port_job)(j) for j in jobs)
        dt = time.perf_counter() - t0
        out["C3_per_timepoint"] = {"items": len(jobs), "seconds": dt, "object_features_per_s": len(jobs) / dt, "cores": cores,
                                   "sample": "one time point: 40 tiles of 96^2, every cell x every instruction (in full)"}
    return out


# ------------------------------------------------------------------------------------------ C5: plate sweep
def c5_sweep(args):
    """BASELINE.json configs[4] at a bounded size: ``--fields`` (default 64) DISTINCT C2 fields sharded over the ranks by
    ``aliby_b200.sharding.extract_sharded`` (contiguous blocks, no collective on the data path), host arrays in, the
    per-field tables gathered on rank 0.  Timed end to end (uploads, kernels, table download, gather) between barriers;
    strong scaling: the sweep is the same whatever the number of ranks.  Rank 0 recomputes a few units on its own and
    checks that the gathered tables are identical."""
    import json

    import torch

    import bench
    from aliby_b200 import extract, sharding

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    n_units = max(world, args.fields if args.fields != 32 else 64)
    units = [7000 + i for i in range(n_units)]
    mine = sharding.shard_units(n_units, rank, world)
    px, lab = bench.make_fields(len(mine), 7000 + int(mine[0]), workers=max(1, (os.cpu_count() or 1) // world))
    px, lab = torch.from_numpy(px).pin_memory().numpy(), torch.from_numpy(lab).pin_memory().numpy()  # a reader's staging buffers
    store = {units[i]: (lab[k], px[k][None]) for k, i in enumerate(mine)}
    tree = bench.c2_tree()

    def load(u):
        return store[u]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sharding.extract_sharded(tree, units[: world], lambda u: store.get(u, next(iter(store.values()))))  # warm-up (with the gather)
    times = []
    out = None
    for _ in range(max(1, args.steps // 5)):
        barrier()
        t0 = time.perf_counter()
        out = sharding.extract_sharded(tree, units, load)
        barrier()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert [u for u, *_ in out] == units
        n_feat = sum(v.size for _, _, _, v in out)
        n_obj = sum(len(o) for _, o, _, _ in out)
        for u, objs, names, vals in out[:: max(1, len(out) // 4)][:4]:  # spot check against this rank alone
            f_px, f_lab = bench._make(u)
            ref = extract.extract_table(tree, f_lab, f_px[None], device=device)
            assert np.array_equal(objs, ref.objects) and names == ref.names
            assert np.array_equal(np.nan_to_num(vals, nan=-1.0), np.nan_to_num(ref.values, nan=-1.0)), u
        line = {
            "metric": "object_features_per_s", "value": n_feat / float(t[0]), "unit": "object-features/s", "n_gpus": world,
            "steps": len(times), "warmup": 1, "ms_per_step": 1e3 * float(t[0]), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": f"C5 plate sweep at a bounded size: {n_units} distinct C2 fields (5ch x 2160^2 uint16, ~2k cells, "
                                   "full cell-function set) through sharding.extract_sharded, host arrays in, tables gathered on rank 0",
                       "fields": n_units, "objects": n_obj, "gathered_tables_checked_against_single_rank": True},
            "e2e": {"value": n_feat / float(t[0]), "unit": "object-features/s",
                    "h2d_bytes_per_step": n_units * (5 * 2160 * 2160 * 2 + 2160 * 2160 * 2), "d2h_bytes_per_step": n_feat * 8},
            "gpu_launches": n_units * bench.LAUNCHES_PER_STEP,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0
