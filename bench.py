#!/usr/bin/env python
"""Benchmark of the per-object feature-extraction hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over one batch of synthetic Cell-Painting fields
(BASELINE.json configs[1]: 5 channels x 2160 x 2160 uint16, ~2k labelled cells per field,
full cell-function set).  Rank r of N owns its own batch (fields shard across GPUs, no
collective on the data path: weak scaling).  The batch (``--fields`` x 65 MB of inputs) is
larger than the 126 MB L2, so every step re-reads its inputs from HBM.

Printed JSON line (rank 0):
  value      object-features/s, inputs resident in HBM, CUDA-event timed, max over ranks; every step finds its own
             per-field label maxima on the device (abx_label_max, what masks.max() is to the reference,
             extract.py:279) and reads them back inside the timed region
  e2e        same metric through the public API ``aliby_b200.extract.extract_table`` with
             pinned HOST inputs: H2D of labels+pixels and D2H of the table inside the timed region;
             ``h2d_only_ms`` = the same bytes copied with no kernel at all (the PCIe / host-memory floor of the box,
             all ranks copying at once)
  e2e_dropin one C2 field per call through process_tree_masks + extract_tree + format_extraction with pageable
             NumPy inputs — the calls pipe_core.py:217 makes
  configs    device-timed and end-to-end numbers of C1, C2 (one field per call), C3 (per time point and batched,
             fused crop, background metrics) and C4 (tools/bench_configs.py)
  roofline   dominant kernel (per-stage CUDA events recorded inside the timed steps) against
             the measured HBM copy peak of MEASURED_PEAKS.json
  cpu_baseline  the reference's CPU algorithm (oracle.port, faithful restatement) on a bounded
             sample of the same workload, 1 core
``--impl reference`` times the reference's CPU algorithm with all host cores (joblib, the
reference's own fan-out, extract.py:360-374) on bounded samples of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stages of abx_extract bracketed by its stage events, and the kernels each one launches
# object_stats: plan kernel + object_sweep (TMA-staged windows; object_stats_warp when the layout does not qualify);
# object_edt: object_edt_grid, with the few (object, request) pairs the sweep kernel left over on a helper stream next to it
STAGES = ["label_scan", "object_stats", "object_edt", "large_objects", "finalize"]
N_STAGES = len(STAGES)
LAUNCHES_PER_STEP = 11  # label_max, init_records (+ bitmaps, sqrt table), label_scan, plan, object_sweep, object_edt_grid,
# object_stats_warp (left-over pairs), object_stats, shape_edt x2, finalize
FIELD = (2160, 2160)
N_CHANNELS = 5
N_OBJECTS = 2000
SHAPE_FEATURES = ["area", "centroid_x", "centroid_y", "conical_volume", "eccentricity", "spherical_volume", "volume"]
INTENSITY_FEATURES = ["max2p5pc", "max5px_median", "mean", "median", "moment_of_inertia", "ratio", "std", "total",
                      "total_squared"]


def c2_tree():
    """Every scalar function of the reference's CELL_FUNS registry on every channel."""
    tree = {"None": {"None": list(SHAPE_FEATURES)}}
    for ch in range(N_CHANNELS):
        tree[ch] = {"max": list(INTENSITY_FEATURES)}
    return tree


def _make(seed):
    from aliby_b200 import synth

    px, lab = synth.make_field(seed, FIELD, N_CHANNELS, N_OBJECTS)
    if os.environ.get("ABX_RELABEL") == "raster":  # experiment: ids in raster order of the objects' first pixel
        n = int(lab.max())
        first = np.full(n + 1, lab.size, dtype=np.int64)
        flat = lab.ravel()
        idx = np.flatnonzero(flat)
        np.minimum.at(first, flat[idx], idx)
        order = np.argsort(first[1:], kind="stable")
        lut = np.zeros(n + 1, dtype=lab.dtype)
        lut[order + 1] = np.arange(1, n + 1, dtype=lab.dtype)
        lab = lut[lab]
    return px[0], lab


def make_fields(n, seed0, workers=None):
    """n distinct C2 fields (pixels (n,5,1,H,W) uint16, labels (n,H,W) uint16)."""
    seeds = [seed0 + i for i in range(n)]
    workers = workers or min(n, os.cpu_count() or 1)
    if workers > 1:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(workers) as pool:
            out = pool.map(_make, seeds)
    else:
        out = [_make(s) for s in seeds]
    return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 os.environ.get("ABX_CLOCK_PERIOD_MS", "20")],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:  # noqa: BLE001
            return None
    return None


def cpu_port_sample(n_objects_sample=6, seed=4242):
    """oracle.port (the reference's algorithm) on a bounded sample of one C2 field, 1 core."""
    from oracle import port

    px, lab = _make(seed)
    rng = np.random.default_rng(0)
    present = np.unique(lab)
    present = present[present > 0]
    pick = rng.choice(present, size=min(n_objects_sample, len(present)), replace=False)
    objs = [(0, int(k)) for k in pick]
    tree = c2_tree()
    t0 = time.perf_counter()
    items, res = port.run_tree_sample(tree, lab, px[None], objs)
    dt = time.perf_counter() - t0
    return len(items) / dt, len(items), dt, int(lab.max())


def cpu_fast_sample(seed=4242):
    """oracle.fast — the O(YX log) sort-by-label CPU implementation of the same numbers (the "fair CPU" of
    SURVEY.md 8d) — on one whole C2 field, 1 core."""
    from oracle import fast

    px, lab = _make(seed)
    t0 = time.perf_counter()
    items, _ = fast.run_tree(c2_tree(), lab, px[None])
    dt = time.perf_counter() - t0
    return len(items) / dt, len(items), dt


# ------------------------------------------------------------------------------------------ reference arm
def _ref_job(args):
    from oracle import port

    lab, px, obj, inst = args
    (tile_i, k), (ch, red, metric) = obj, inst
    plane = lab == k
    img = None
    if ch != "None":
        img = port.project_z(px[ch], port.Z_REDUCERS[red])
    return float(port.CELL_METRICS[metric](plane, img))


def reference_arm(args):
    """The reference's own CPU algorithm on all host cores (joblib, like extract.py:360-374)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from joblib import Parallel, delayed

    from oracle import port

    cores = os.cpu_count() or 1
    px, lab = _make(4242)
    tree = c2_tree()
    instructions = port.tree_instructions(tree)
    present = np.unique(lab)
    present = present[present > 0]
    rng = np.random.default_rng(1)
    per_step = max(2, min(8, cores // 8 or 1))  # objects per step: a bounded sample of one field
    times = []
    n_feat = 0
    with Parallel(n_jobs=cores, backend="loky") as par:
        for step in range(args.warmup + args.steps):
            pick = rng.choice(present, size=per_step, replace=False)
            jobs = [(lab, px, (0, int(k)), inst) for k in pick for inst in instructions]
            t0 = time.perf_counter()
            res = par(delayed(_ref_job)(j) for j in jobs)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
                n_feat += len(res)
    total = sum(times)
    value = n_feat / total
    from tools import bench_configs

    ref_configs = bench_configs.reference_configs(cores)
    line = {
        "impl": "reference",
        "metric": "object_features_per_s",
        "value": value,
        "unit": "object-features/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u16",
        "data": "synthetic",
        "config": workload_config(per_step_note=f"{per_step} objects x {len(instructions)} instructions per step"),
        "cpu_baseline": {
            "value": value, "unit": "object-features/s", "cores": cores, "kind": "port",
            "sample": f"{per_step} random objects x {len(instructions)} instructions of one C2 field per step, "
                      f"joblib loky over {cores} cores (oracle.port = faithful restatement; the Python reference "
                      "itself cannot travel to the GPU box)",
        },
        "e2e": {"value": value, "unit": "object-features/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "configs": ref_configs,  # C1 and one C3 time point in full, same algorithm, all cores
    }
    print(json.dumps(line))
    return 0


def workload_config(fields=None, per_step_note=None):
    cfg = {
        "workload": "C2 Cell Painting: 5ch x 2160x2160 uint16 field, ~2k labelled cells, full cell-function set "
                    f"({len(SHAPE_FEATURES)} shape + {N_CHANNELS}x{len(INTENSITY_FEATURES)} intensity features)",
        "features_per_object": len(SHAPE_FEATURES) + N_CHANNELS * len(INTENSITY_FEATURES),
        "l2_policy": "inputs larger than L2 (batch of fields per step, 65 MB each)",
    }
    if fields is not None:
        cfg["fields_per_step_per_gpu"] = fields
    if per_step_note:
        cfg["reference_sample"] = per_step_note
    return cfg


# ------------------------------------------------------------------------------------------ our arm
def ours(args):
    import ctypes as C

    import torch

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
    if world > 1 and hasattr(os, "sched_setaffinity"):
        # one slice of the host's cores per rank, chosen BEFORE the pinned buffers are allocated (first touch): the ranks'
        # copy threads and staging memory then do not share cores (all eight GPUs of the box hang off one NUMA node)
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // world)
        mine = cpus[local_rank * per:(local_rank + 1) * per] or cpus
        os.sched_setaffinity(0, mine)

    from aliby_b200 import _native as nat
    from aliby_b200 import engine, extract

    lib = nat.lib()
    F = args.fields
    tree = c2_tree()
    plan = engine.compile_tree(tree)
    n_feat_cols = len(plan.instructions)

    # ---- synthetic batch (distinct per rank) ----
    workers = max(1, (os.cpu_count() or 1) // max(1, world))
    px_np, lab_np = make_fields(F, int(os.environ.get("ABX_SEED_BASE", "5000")) + 100 * rank, workers=min(F, workers))
    n_labels = lab_np.reshape(F, -1).max(axis=1).astype(np.int64)
    n_objects = int(n_labels.sum())
    H, W = FIELD
    in_bytes = px_np.nbytes + lab_np.nbytes
    table_bytes = n_objects * plan.n_columns * 8
    algo_bytes = in_bytes + table_bytes  # C*Z*Y*X*2 + Y*X*2 + rows*cols*8 per field

    # pinned host copies for the e2e leg, device copies for the resident leg
    px_pin = torch.from_numpy(px_np).pin_memory()
    lab_pin = torch.from_numpy(lab_np).pin_memory()
    px_dev = px_pin.to(device)
    lab_dev = lab_pin.to(device)
    offs = np.arange(F, dtype=np.int64) * (N_CHANNELS * H * W)
    plane_tile = np.arange(F, dtype=np.int32)
    out_buf, out, status = engine.alloc_table(n_objects, plan.n_columns, device)

    # The per-field label maxima (masks.max() of extract.py:279) are part of every step: abx_label_max on a side stream,
    # read back through pinned memory.  The maxima of step k + 2 are found while the kernels of step k run (two buffers
    # in rotation), so the launch stream never waits for the host and the host may fall up to two steps behind the GPU
    # (a hiccup of the launching thread — the clock sampler's NVML queries take driver locks — does not drain the queue).
    side = torch.cuda.Stream(device=device)
    nmax_host = [torch.zeros(F, dtype=torch.int32).pin_memory() for _ in range(2)]
    nmax_ready = [torch.cuda.Event() for _ in range(2)]
    step_no = [0]

    def launch_label_max(slot):
        with torch.cuda.stream(side):
            nmax_host[slot].copy_(extract._label_max(lab_dev, device), non_blocking=True)
            nmax_ready[slot].record(side)

    launch_label_max(0)
    launch_label_max(1)

    def step(events=None):
        slot = step_no[0] & 1
        step_no[0] += 1
        nmax_ready[slot].synchronize()
        n_lab = nmax_host[slot].numpy().astype(np.int64)
        launch_label_max(slot)  # for the step after the next one
        engine.run_planes(plan, lab_dev, plane_tile, n_lab, px_dev, offs, H * W, H * W, W, N_CHANNELS, 1,
                          out=out, stage_events=events, status=status)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()  # samples clocks over the warm-up, the timed steps and the e2e leg (the timed steps alone are ~10 ms)
    for _ in range(max(3, args.warmup)):
        step()
    barrier()

    # ---- resident leg: K steps between two CUDA events on the launch stream ----
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    h0 = time.perf_counter()
    for k in range(args.steps):
        step()
    host_ms = 1e3 * (time.perf_counter() - h0) / args.steps  # host time to enqueue one step (incl. the wait for the maxima)
    t_end.record()
    barrier()
    ms_total = t_start.elapsed_time(t_end)
    # ---- stage profile: the same step with stage events, which keeps the statistics and the shape chain in line on one
    # stream (abx_extract runs them on two streams otherwise) so that every stage can be timed on its own ----
    stage_ev = []
    for _ in range(args.steps):
        evs = []
        for _ in range(N_STAGES + 1):
            h = C.c_void_p()
            nat.check(lib.abx_event_create(C.byref(h)), "abx_event_create")
            evs.append(h)
        stage_ev.append(evs)
    for k in range(args.steps):
        step(stage_ev[k])
    barrier()
    stage_ms = np.zeros(N_STAGES)
    for evs in stage_ev:
        for i in range(N_STAGES):
            ms = C.c_float()
            nat.check(lib.abx_event_elapsed_ms(evs[i], evs[i + 1], C.byref(ms)), "abx_event_elapsed_ms")
            stage_ms[i] += ms.value
        for h in evs:
            lib.abx_event_destroy(h)
    stage_ms /= args.steps
    engine.raise_on_status(int(status.cpu()[0]))
    torch.cuda.synchronize()
    assert all(np.array_equal(h.numpy().astype(np.int64), n_labels) for h in nmax_host)

    # ---- e2e leg: public API, pinned host inputs, H2D + D2H inside the timed region ----
    masks_host = [lab_pin[i].numpy() for i in range(F)]
    px_host = px_pin.numpy()

    def e2e_step():
        return extract.extract_table(tree, masks_host, px_host, device=device, plan=plan)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_s = float("nan")
    if not args.no_e2e:
        for _ in range(2):
            tab = e2e_step()
        assert tab.values.shape == (n_objects, n_feat_cols)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            tab = e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0

    # ---- the copy floor: the same host-to-device bytes with no kernel, every rank at once ----
    h2d_ms = float("nan")
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            lab_dev.copy_(lab_pin, non_blocking=True)
            px_dev.copy_(px_pin, non_blocking=True)
        torch.cuda.synchronize()
        h2d_ms = 1e3 * (time.perf_counter() - t0) / 3

    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed steps + e2e leg"

    # ---- reduce over ranks: max time, sum of units ----
    per_rank = [ms_total / args.steps]
    if dist is not None:  # every rank's own step time and stage sum, for the record (the reported time is their maximum)
        mine = torch.tensor([ms_total / args.steps, float(stage_ms.sum())], dtype=torch.float64, device=device)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(x[0]), 4), round(float(x[1]), 4)] for x in allr]
    t = torch.tensor([ms_total, e2e_s, h2d_ms], dtype=torch.float64, device=device)
    units = torch.tensor([float(n_objects * n_feat_cols), float(algo_bytes)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    ms_total, e2e_s, h2d_ms = t.tolist()
    feat_per_step, bytes_per_step = units.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        names = STAGES
        dom = int(np.argmax(stage_ms))
        achieved = (algo_bytes / 1e9) / (stage_ms[dom] / 1e3)
        prof = profiled_traffic() or {}
        step_traffic = sum(v for k, v in prof.items() if k in STAGES and isinstance(v, (int, float)))
        line = {
            "metric": "object_features_per_s",
            "value": feat_per_step * args.steps / (ms_total / 1e3),
            "unit": "object-features/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u16",
            "data": "synthetic",
            "config": workload_config(fields=F),
            "image_gbs": (bytes_per_step * args.steps / 1e9) / (ms_total / 1e3),
            "image_gbs_frac_of_hbm_peak": (bytes_per_step / world * args.steps / 1e9) / (ms_total / 1e3) / peak,
            "objects_per_step": n_objects if world == 1 else None,
            "stage_ms": {n: float(v) for n, v in zip(names, stage_ms)},  # separate steps, the two chains in line
            "stage_ms_sum": float(stage_ms.sum()),
            "roofline": {
                "bound": "hbm",
                "kernel": names[dom],
                "achieved": achieved,
                "peak": peak,
                "peak_source": peak_src,
                "unit": "GB/s",
                "frac": achieved / peak,
                "algorithmic_bytes_per_launch": algo_bytes,
                "traffic": prof.get(names[dom]),
                # all hot kernels of one step (ncu dram bytes, profiles/traffic.json) over the step's algorithmic bytes
                "step_traffic_over_algorithmic": (step_traffic / algo_bytes) if step_traffic else None,
            },
            "e2e": {
                "value": feat_per_step * e2e_steps / e2e_s,
                "unit": "object-features/s",
                "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": table_bytes + 4 * F,
                "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "h2d_only_ms": h2d_ms,  # max over ranks, all ranks copying at once: what the box's PCIe / host memory allows
                "frac_of_copy_floor": h2d_ms / (1e3 * e2e_s / e2e_steps) if e2e_s == e2e_s else None,
                "api": "aliby_b200.extract.extract_table(tree, masks, pixels) with pinned host arrays",
            },
            "gpu_launches": args.steps * LAUNCHES_PER_STEP,
            "host_enqueue_ms_per_step": host_ms,
            "ms_per_step_by_rank": per_rank,  # N > 1: [step time, sum of the stage times] of every rank
            "host_cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count(),
            "clocks": clocks,
        }
        if world == 1 and not args.no_configs:
            from tools import bench_configs

            del px_dev, lab_dev, out_buf
            torch.cuda.empty_cache()
            line["e2e_dropin"] = bench_configs.dropin_call_ms(device)
            line["configs"] = bench_configs.measure_configs(device, peak)
        if world == 1 and not args.no_cpu:
            v, n_items, dt, _ = cpu_port_sample()
            line["cpu_baseline"] = {
                "value": v, "unit": "object-features/s", "cores": 1, "kind": "port",
                "sample": f"{n_items} (object, instruction) items = 6 random objects x {n_feat_cols} instructions of one "
                          f"C2 field in {dt:.1f} s; oracle.port, the faithful restatement of the reference loop",
            }
            v, n_items, dt = cpu_fast_sample()
            line["cpu_fast"] = {
                "value": v, "unit": "object-features/s", "cores": 1, "kind": "port-fast",
                "sample": f"{n_items} items = every object x {n_feat_cols} instructions of one whole C2 field "
                          f"in {dt:.1f} s; oracle.fast (sort-by-label NumPy/SciPy, same numbers as the reference algorithm)",
            }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fields", type=int, default=32, help="C2 fields per step per GPU (1.8 GB of inputs; 8 -> 1.59e9, 32 -> 1.92e9 object-features/s: launch latencies and kernel tails amortise)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration section (C1, C2 single field, C3, C4)")
    ap.add_argument("--config", default="c2", choices=["c2", "c5"],
                    help="c2 (default): the headline workload; c5: the plate sweep through sharding.extract_sharded "
                         "(distinct fields sharded over the ranks, gathered tables checked against a single rank)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.config == "c5":
        from tools import bench_configs

        return bench_configs.c5_sweep(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
