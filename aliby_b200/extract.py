"""Drop-in for ``src/extraction/extract.py`` of the reference, backed by the CUDA hot path.

Same names, signatures, return structures and error behaviour as the reference's
functional extraction API, so that ``pipeline["steps"]["extract_*"]`` can be served by
this module unchanged (see ``aliby_b200.pipe.init_step``):

* :func:`process_tree_masks`, :func:`process_tree_masks_overlap`  (extract.py:240-301, 456-517)
* :func:`extract_tree`                                            (extract.py:304-375)
* :func:`extract_tree_multi`                                      (extract.py:378-453)
* :func:`format_extraction`                                       (extract.py:520-599)
* :func:`flatten`, :func:`kv`                                     (extract.py:33-74)

Differences, all deliberate: results are Python ``float`` (the reference returns
``np.int64``/``np.uint64`` for ``area``/``total`` which its own ``format_extraction``
rejects); ``ncores``/``progress_bar`` are accepted and ignored (one launch does
the whole |objects| x |instructions| product); cp_measure features have no kernel and
raise ``KeyError`` naming the metric.
"""

from __future__ import annotations

from collections.abc import Sequence
from itertools import product, repeat

import numpy as np

from . import engine
from .engine import flatten, kv  # noqa: F401  (re-exported, same names as the reference)
from .functions.loaders import load_funs, load_redfuns

CELL_FUNS, TRAP_FUNS, ALL_FUNS = load_funs()
REDUCTION_FUNS = load_redfuns()


class ExtractionResults(list):
    """``list`` of per-(object, instruction) results that also carries the dense table."""

    dense: np.ndarray | None = None  # (n_objects, n_dense_columns) float64, host
    plan: engine.Plan | None = None
    objects: np.ndarray | None = None  # (n_objects, 2|3) int64 object ids
    items: tuple | None = None


class ItemTuple(Sequence):
    """The ``tileid_instructions`` of :func:`process_tree_masks`: ``product(objects, instructions)``, object-major, as an
    immutable sequence that is NOT materialised — a C2 field is 2 000 objects x 52 instructions = 104 000 nested tuples,
    10 ms of pure allocation per call that the only consumer the reference has (``format_extraction``: one ``zip`` over
    it, extract.py:536) does not need.  Indexing, slicing, iteration, ``len``, ``==`` with a tuple and ``tuple(items)``
    behave like the reference's tuple.  It also carries what it was built from (the compiled plan, the object list, the
    id count of every plane) so that :func:`extract_tree` need not re-derive them; nothing is kept in module state, so
    concurrent callers do not interfere."""

    plan: "engine.Plan | None" = None
    counts: "dict | None" = None  # plane key ((tile,) or (tile, stack)) -> number of ids enumerated for it
    label_stack: "np.ndarray | None" = None  # (tiles, Y, X) uint16 stack of the masks, when process_tree_masks made one

    def __init__(self, objects: list, instructions: list):
        self.objects = objects
        self.instructions = instructions

    def __len__(self):
        return len(self.objects) * len(self.instructions)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return tuple(self[j] for j in range(*i.indices(len(self))))
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("tuple index out of range")
        o, k = divmod(i, len(self.instructions))
        return (self.objects[o], self.instructions[k])

    def __iter__(self):
        return product(self.objects, self.instructions)

    def __eq__(self, other):
        if isinstance(other, (tuple, list, ItemTuple)):
            return len(self) == len(other) and all(a == b for a, b in zip(self, other))
        return NotImplemented

    def __hash__(self):
        return hash((tuple(self.objects), tuple(self.instructions)))

    def __repr__(self):
        return f"ItemTuple({len(self.objects)} objects x {len(self.instructions)} instructions)"


def _as_mask_list(masks):
    if not isinstance(masks, list):  # "Hacky fix when tile level is not provided" (extract.py:271-272)
        masks = [masks]
    return masks


def _host_label_stack(planes: list[np.ndarray]) -> np.ndarray:
    """Equally-shaped 2-D label planes as one contiguous uint16 array (a view when they already are one)."""
    if all(isinstance(p, np.ndarray) and p.dtype == np.uint16 and p.flags.c_contiguous for p in planes):
        if len(planes) == 1:
            return planes[0][None]
        if planes[0].nbytes >= (1 << 19):  # (small tiles: stacking them costs less than checking their addresses)
            run = _contiguous_run(planes)
            if run is not None:
                return run
    stack = np.stack(planes)
    if stack.dtype != np.uint16:
        if stack.size and (stack.min() < 0 or stack.max() > 65535):
            raise OverflowError("label ids must fit uint16 (segment/dispatch.py:14-19 enforces the same)")
        stack = stack.astype(np.uint16)
    return np.ascontiguousarray(stack)


def _to_device_labels(planes: list[np.ndarray], device):
    """Stack equally-shaped 2-D label planes into one uint16 device tensor."""
    import torch

    return torch.from_numpy(_host_label_stack(planes)).to(device, non_blocking=True)


# ---- CUDA graphs for the small-call regime (one call per time point with a few hundred objects) ----
_GRAPH_MAX_OBJECTS = 8192   # above this the launches are no longer what the call costs
_GRAPH_AFTER_CALLS = 2      # eager calls with the same shapes before a graph is captured
_graph_seen: dict = {}
_graph_cache: dict = {}
_graph_lock = __import__("threading").Lock()


def _graphs_enabled() -> bool:
    import os

    return os.environ.get("ALIBY_B200_GRAPHS", "1") != "0"


def _graphed(key, build):
    """Per (thread, key) instance of engine.GraphedExtract, built after the key has been seen a few times."""
    import threading

    key = (threading.get_ident(), *key)
    g = _graph_cache.get(key)
    if g is not None:
        return g
    seen = _graph_seen.get(key, 0) + 1
    _graph_seen[key] = seen
    if seen <= _GRAPH_AFTER_CALLS:
        return None
    g = build()
    with _graph_lock:  # (entries are per thread, the dictionaries are shared)
        while len(_graph_cache) >= 8:
            _graph_cache.pop(next(iter(_graph_cache)))
        if len(_graph_seen) > 256:
            _graph_seen.clear()
        _graph_cache[key] = g
    return g


def _run_dense(plan, planes, plane_tile, n_labels, pixels, device=None, strict_labels: bool = True):
    """planes: list of (Y, X) arrays, or the (P, Y, X) uint16 stack of them; pixels: (tiles, C, Z, Y, X) ndarray / tensor /
    TileView.  ``strict_labels=False``: ids above a plane's ``n_labels`` are ignored instead of raising (the live overlap
    path of the reference enumerates 1..k for k distinct ids and never looks at larger ones)."""
    import torch

    from .tile import TileView

    n_objects = int(np.sum(n_labels))
    if n_objects == 0:
        return np.zeros((0, plan.n_columns))
    if plan.error is not None:  # (before anything touches CUDA: an unknown metric is a host-side error)
        raise plan.error
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    host_labels = planes if isinstance(planes, np.ndarray) else _host_label_stack(planes)
    # ---- small calls of repeating shapes: replay a captured graph (engine.GraphedExtract) ----
    if (_graphs_enabled() and n_objects <= _GRAPH_MAX_OBJECTS and plan.n_columns and plan.requests
            and isinstance(pixels, (np.ndarray, TileView)) and str(pixels.dtype) in ("uint8", "uint16")
            and not (isinstance(pixels, TileView) and pixels.out_of_frame.any())):
        n_labels = np.asarray(n_labels, dtype=np.int64)
        cap = max(16, int(-(-2 * int(n_labels.max()) // 8) * 8))
        P, H, W = host_labels.shape
        if isinstance(pixels, TileView):
            frame = np.ascontiguousarray(pixels.frame)
            C_, Z_, FH, FW = frame.shape
            offs = (pixels.origins[:, 0] * FW + pixels.origins[:, 1]).astype(np.int64)
            layout = (frame.shape, str(frame.dtype), offs.tobytes(), Z_ * FH * FW, FH * FW, FW, C_, Z_)
            src = frame
        else:
            T, C_, Z_, Y, X = pixels.shape
            offs = np.arange(T, dtype=np.int64) * (C_ * Z_ * Y * X)
            layout = (pixels.shape, str(pixels.dtype), offs.tobytes(), Z_ * Y * X, Y * X, X, C_, Z_)
            src = np.ascontiguousarray(pixels)
        key = (str(device), tuple(plan.instructions), (P, H, W), tuple(np.asarray(plane_tile).tolist()), layout)
        cur = _graph_cache.get((__import__("threading").get_ident(), *key))
        if cur is not None and cur.cap < int(n_labels.max()):  # more labels than captured: a larger instance
            _graph_cache.pop((__import__("threading").get_ident(), *key), None)
            cur = None
        g = cur or _graphed(key, lambda: engine.GraphedExtract(
            plan, P, H, W, plane_tile, layout[0], getattr(torch, layout[1]), offs, layout[3], layout[4], layout[5],
            layout[6], layout[7], cap, device))
        if g is not None:
            g.run(host_labels, src)
            host = g.buf.cpu()
            n_cells = g.P * g.cap * plan.n_columns
            engine.raise_on_status(int(host[n_cells:].view(torch.int32)[0]) & (~0 if strict_labels else ~1))
            return host[:n_cells].view(g.P * g.cap, plan.n_columns).numpy()[g.rows(n_labels)]
    labels_dev = torch.empty(host_labels.shape, dtype=torch.uint16, device=device)
    engine.host_to_device(labels_dev, host_labels, slot="labels")
    if isinstance(pixels, TileView):
        px_dev, offs, cs, zs, rs, C_, Z_ = pixels.addressing(device)
    else:
        if isinstance(pixels, np.ndarray):
            if plan.requests:
                if str(pixels.dtype) not in engine.PIXEL_DTYPES:
                    raise NotImplementedError(
                        f"pixel dtype {pixels.dtype} has no CUDA kernel in aliby_b200 "
                        "(uint8/uint16/float32/float64) and there is no CPU fallback"
                    )
                px_host = np.ascontiguousarray(pixels)
                px_dev = torch.empty(px_host.shape, dtype=getattr(torch, str(px_host.dtype)), device=device)
                engine.host_to_device(px_dev, px_host)
            else:
                px_dev = torch.empty(0, dtype=torch.uint16, device=device)
        else:
            px_dev = pixels.to(device).contiguous()
        T, C_, Z_, Y, X = pixels.shape
        offs = np.arange(T, dtype=np.int64) * (C_ * Z_ * Y * X)
        cs, zs, rs = Z_ * Y * X, Y * X, X
    buf, table, status = engine.alloc_table(n_objects, plan.n_columns, device)
    engine.run_planes(plan, labels_dev, plane_tile, n_labels, px_dev, offs, cs, zs, rs, C_, Z_, out=table, status=status)
    host = buf.cpu()  # one copy: the table and the call's error flags
    engine.raise_on_status(int(host[n_objects * plan.n_columns :].view(torch.int32)[0]) & (~0 if strict_labels else ~1))
    return host[: n_objects * plan.n_columns].view(n_objects, plan.n_columns).numpy()


class ExtractionTable:
    """Dense result of :func:`extract_table`: one row per object, one column per feature."""

    def __init__(self, objects: np.ndarray, names: list[str], values: np.ndarray):
        self.objects = objects  # (n, 2) int64: tile, label
        self.names = names  # reference column names "{ch}/{red}/{metric}/{metric}"
        self.values = values  # (n, len(names)) float64, host

    def to_arrow(self):
        """Same table as ``format_extraction`` builds (tile, label, sorted feature columns)."""
        import pyarrow as pa

        data = {"tile": pa.array(self.objects[:, 0], pa.int64()), "label": pa.array(self.objects[:, 1], pa.int64())}
        for j in np.argsort(np.asarray(self.names, dtype=object), kind="stable") if self.names else []:
            data[self.names[j]] = pa.array(np.ascontiguousarray(self.values[:, j]), pa.float64())
        return pa.table(data)


def _label_max(labels_dev, device):
    """Per-plane maximum label on the device (the reference's ``masks.max()``, extract.py:279)."""
    import ctypes as C

    import torch

    from . import _native as nat

    P, H, W = labels_dev.shape
    nmax = torch.empty(P, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        nat.check(
            nat.lib().abx_label_max(
                labels_dev.data_ptr(), nat.U16, P, H, W, labels_dev.stride(0), labels_dev.stride(1), nmax.data_ptr(),
                C.c_void_p(torch.cuda.current_stream(device).cuda_stream),
            ),
            "abx_label_max",
        )
    return nmax


_copy_streams: dict = {}


def _copy_stream(device):
    import torch

    key = str(device)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=device)
    return _copy_streams[key]


def _host_u16(plane) -> "np.ndarray":
    plane = np.asarray(plane)
    if plane.dtype != np.uint16:
        if plane.size and (plane.min() < 0 or plane.max() > 65535):
            raise OverflowError("label ids must fit uint16 (segment/dispatch.py:14-19 enforces the same)")
        plane = plane.astype(np.uint16)
    return np.ascontiguousarray(plane)


def _contiguous_run(planes):
    """One (k, Y, X) view over `planes` when they sit back to back in memory (slices of one array), else None."""
    if len(planes) < 2:
        return None
    first = planes[0]
    step = first.nbytes
    addr0 = first.__array_interface__["data"][0]
    for j, pl in enumerate(planes):
        if pl.shape != first.shape or not pl.flags.c_contiguous or pl.__array_interface__["data"][0] != addr0 + j * step:
            return None
    # the planes themselves keep the memory alive while the view is in use
    return np.lib.stride_tricks.as_strided(first, shape=(len(planes), *first.shape), strides=(step, *first.strides))


def _pipelined_host_table(plan, masks, keep, pixels, device, chunk_bytes):
    """Host arrays in, host table out, with the uploads overlapping the kernels.

    Copy stream: all label planes first (1/6 of the bytes), then the pixels in chunks of tiles (asynchronous
    when the arrays are pinned).  Compute stream: the per-plane label maxima as soon as the labels have landed
    (their D2H is the only host synchronisation before the end: the row count sizes the table), then one
    ``abx_extract`` per chunk as its pixels land, all writing into one device table that goes back with a
    single D2H copy.  Returns ``(values (n, dense columns), n_labels per kept tile)``."""
    import torch

    T, C_, Z_, Y, X = pixels.shape
    tile_bytes = C_ * Z_ * Y * X * pixels.dtype.itemsize
    per_chunk = max(1, int(chunk_bytes // max(1, tile_bytes)))
    chunks = [keep[i : i + per_chunk] for i in range(0, len(keep), per_chunk)]
    cur = torch.cuda.current_stream(device)
    cps = _copy_stream(device)
    cps.wait_stream(cur)  # buffers handed out by the caching allocator may still be in use on `cur`
    staged = []
    with torch.cuda.stream(cps):
        lab = torch.empty((len(keep), Y, X), dtype=torch.uint16, device=device)
        planes = [_host_u16(masks[t]) for t in keep]
        run = _contiguous_run(planes)
        if run is not None:  # the planes are slices of one array: one copy instead of one per plane
            engine.host_to_device(lab, run, slot="labels")
        else:
            for j, plane in enumerate(planes):
                engine.host_to_device(lab[j], plane, slot="labels")
        ev_lab = torch.cuda.Event()
        ev_lab.record(cps)
        for tiles in chunks:
            if plan.requests:
                consecutive = tiles == list(range(tiles[0], tiles[0] + len(tiles)))
                src = pixels[tiles[0] : tiles[0] + len(tiles)] if consecutive else pixels[tiles]
                px = torch.empty(src.shape, dtype=getattr(torch, str(src.dtype)), device=device)
                engine.host_to_device(px, np.ascontiguousarray(src))
            else:
                px = torch.empty(0, dtype=torch.uint16, device=device)
            ev = torch.cuda.Event()
            ev.record(cps)
            staged.append((tiles, px, ev))
    cur.wait_event(ev_lab)
    nmax_host = torch.empty(len(keep), dtype=torch.int32, pin_memory=True)
    nmax_host.copy_(_label_max(lab, device), non_blocking=True)
    done = torch.cuda.Event()
    done.record(cur)
    done.synchronize()  # the row count has to reach the host; the pixel uploads keep running meanwhile
    n_labels = nmax_host.numpy().astype(np.int64)
    rows = np.concatenate([[0], np.cumsum(n_labels)])
    n_rows, n_cols = int(rows[-1]), plan.n_columns
    buf, table, status = engine.alloc_table(n_rows, n_cols, device, n_status=len(staged))
    p0 = 0
    for k, (tiles, px, ev) in enumerate(staged):
        cur.wait_event(ev)
        p1 = p0 + len(tiles)
        offs = np.arange(len(tiles), dtype=np.int64) * (C_ * Z_ * Y * X)
        engine.run_planes(plan, lab[p0:p1], np.arange(len(tiles), dtype=np.int32), n_labels[p0:p1], px, offs,
                          Z_ * Y * X, Y * X, X, C_, Z_, out=table[int(rows[p0]) : int(rows[p1])], status=status[2 * k :])
        p0 = p1
    host = torch.empty(buf.shape, dtype=torch.float64, pin_memory=True)
    host.copy_(buf, non_blocking=True)  # one copy: the table and every chunk's error flags
    cur.synchronize()
    for word in host[n_rows * n_cols :].view(torch.int32)[::2]:
        engine.raise_on_status(word)
    values = host[: n_rows * n_cols].view(n_rows, n_cols).numpy()
    if values.nbytes <= (8 << 20):
        # a caller that keeps its tables (a plate sweep, one field per call) would keep the pinned staging block of every
        # one of them, and every later call would pay a fresh cudaHostAlloc (3 ms per C2 field): small tables are handed
        # over as pageable copies (0.1 ms), the block goes back to the host allocator's cache
        values = values.copy()
    return values, n_labels


def extract_table(tree: dict, masks, pixels, device=None, plan: engine.Plan | None = None,
                  chunk_bytes: int = 160 << 20, cp_measure_kwargs=None) -> ExtractionTable:
    """Fast public entry point: tree + host (or device) arrays in, dense per-object table out.

    Does what ``process_tree_masks`` + ``extract_tree`` + the pivot of ``format_extraction`` do
    (extract.py:240-375, 574-598) without materialising the |objects| x |instructions| Python
    lists.  Host inputs are uploaded in chunks of tiles on a copy stream while earlier chunks are
    being extracted (pinned arrays make the copies asynchronous); the per-plane label maxima are
    found on the device (the reference's ``masks.max()``, extract.py:279), one ``abx_extract`` call
    per chunk fills the table and an asynchronous device-to-host copy returns it."""
    import torch

    from .tile import TileView

    masks = _as_mask_list(masks)
    plan = plan or engine.compile_tree(tree, cp_measure_kwargs)
    if any(len(c) != 1 and k is None for c, k in zip(plan.inst_cols, plan.inst_keys) if c):
        raise Exception("tuple-valued metrics (centroid, min_maj_approximation) cannot be table columns")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    keep = [i for i, m in enumerate(masks) if len(m)]
    names, name_cols = [], []  # reference column names; a dict-valued metric contributes one column per key
    for inst, cols_, keys_ in zip(plan.instructions, plan.inst_cols, plan.inst_keys):
        branch = "/".join(str(x) for x in inst)
        for k, j in zip(keys_ if keys_ is not None else [inst[-1]], cols_):
            names.append(f"{branch}/{k}")
            name_cols.append(j)
    if not keep:
        return ExtractionTable(np.zeros((0, 2), np.int64), names, np.zeros((0, len(names))))
    host_inputs = isinstance(pixels, np.ndarray) and not isinstance(masks[keep[0]], torch.Tensor)
    if host_inputs:
        if plan.requests and str(pixels.dtype) not in engine.PIXEL_DTYPES:
            raise NotImplementedError(
                f"pixel dtype {pixels.dtype} has no CUDA kernel in aliby_b200 (uint8/uint16/float32/float64) and there is no CPU fallback"
            )
        values, n_labels = _pipelined_host_table(plan, masks, keep, pixels, device, chunk_bytes)
    else:
        if isinstance(masks[keep[0]], torch.Tensor):
            labels_dev = torch.stack([masks[i] for i in keep]).to(device)
        else:
            labels_dev = _to_device_labels([np.asarray(masks[i]) for i in keep], device)
        nmax = _label_max(labels_dev, device)
        if isinstance(pixels, TileView):
            px_dev, offs, cs, zs, rs, C_, Z_ = pixels.addressing(device)
        else:
            px_dev = pixels.to(device).contiguous()
            T, C_, Z_, Y, X = px_dev.shape
            offs = np.arange(T, dtype=np.int64) * (C_ * Z_ * Y * X)
            cs, zs, rs = Z_ * Y * X, Y * X, X
        n_labels = nmax.cpu().numpy().astype(np.int64)  # small D2H: the row count has to reach the host
        n_rows = int(n_labels.sum())
        buf, table, status = engine.alloc_table(n_rows, plan.n_columns, device)
        engine.run_planes(plan, labels_dev, np.asarray(keep, dtype=np.int32), n_labels, px_dev, offs, cs, zs, rs, C_, Z_,
                          out=table, status=status)
        host = buf.cpu()
        engine.raise_on_status(host[n_rows * plan.n_columns :].view(torch.int32)[0])
        values = host[: n_rows * plan.n_columns].view(n_rows, plan.n_columns).numpy()
    cols = np.asarray(name_cols, dtype=np.int64)
    objects = np.stack(
        [np.repeat(np.asarray(keep, dtype=np.int64), n_labels), np.concatenate([np.arange(1, k + 1) for k in n_labels])],
        axis=1,
    ) if n_labels.sum() else np.zeros((0, 2), np.int64)
    if not len(cols):
        values = values[:, :0]
    elif not np.array_equal(cols, np.arange(values.shape[1])):  # duplicate instructions share a dense column
        values = values[:, cols]
    return ExtractionTable(objects, names, values)


def _results_from_dense(plan, dense, row_of_item, inst_of_item):
    """Flat python list in item order; tuple-valued metrics become tuples, cp_measure features dicts of
    ``{key: ndarray of length 1}`` (what ``wrap_cp_measure_features`` returns, loaders.py:135-150)."""
    single = all(len(c) == 1 and k is None for c, k in zip(plan.inst_cols, plan.inst_keys))
    if single and len(row_of_item):
        cols = np.fromiter((c[0] for c in plan.inst_cols), dtype=np.int64)
        if inst_of_item is None:  # the object-major product over the rows `row_of_item` (one entry per OBJECT)
            return dense[row_of_item][:, cols].ravel().tolist()
        return dense[row_of_item, cols[inst_of_item]].tolist()
    if inst_of_item is None:
        # object-major product with dict- or tuple-valued metrics: one column block per instruction, cut into per-object
        # values by C-level iteration (a (k, 1) view iterates into k arrays of length 1), then interleaved object-major
        n_inst, n_obj = len(plan.inst_cols), len(row_of_item)
        per_inst = []
        for c, keys in zip(plan.inst_cols, plan.inst_keys):
            if not len(c):
                per_inst.append([None] * n_obj)  # (an instruction that failed to compile: plan.error was raised before)
                continue
            block = dense[row_of_item][:, list(c)]
            if keys is not None:
                per_inst.append([dict(zip(keys, sub)) for sub in block.reshape(n_obj, len(c), 1)])
            elif len(c) == 1:
                per_inst.append(block[:, 0].tolist())
            else:
                per_inst.append([tuple(v) for v in block.tolist()])
        out = [None] * (n_obj * n_inst)
        for i, vals in enumerate(per_inst):
            out[i::n_inst] = vals
        return out
    out = []
    for r, i in zip(row_of_item, inst_of_item):
        c, keys = plan.inst_cols[i], plan.inst_keys[i]
        if keys is not None:
            out.append({k: dense[r, j : j + 1].copy() for k, j in zip(keys, c)})
        else:
            out.append(float(dense[r, c[0]]) if len(c) == 1 else tuple(float(dense[r, k]) for k in c))
    return out


def process_tree_masks(
    tree: dict,
    masks,
    pixels: np.ndarray,
    measure_fn,
    ncores=None,
    progress_bar: bool = False,
    cp_measure_kwargs=None,
):
    """Same contract as the reference (extract.py:240-301): ``(tileid_instructions, results)``
    with ``tileid_instructions = product(objects, instructions)``, object-major; objects are
    every id ``1..max`` of every non-empty tile, absent ids included."""
    masks = _as_mask_list(masks)
    instructions = kv(flatten(tree))
    ind_masks = []
    counts = {}
    stack = None
    if len(masks) > 1 and all(isinstance(m, np.ndarray) and m.ndim == 2 and m.size and m.shape == masks[0].shape for m in masks):
        # many small tiles (a time point of a trap position): one stack, one reduction for all the maxima — and the
        # stack is what goes to the device afterwards
        try:
            stack = _host_label_stack(masks)
            maxima = stack.reshape(len(masks), -1).max(axis=1).tolist()
        except OverflowError:
            stack = None
    for tile_i, masks_in_tile in enumerate(masks):
        if len(masks_in_tile):
            counts[(tile_i,)] = k = int(maxima[tile_i] if stack is not None else masks_in_tile.max())
            ind_masks.extend(zip(repeat(tile_i, k), range(1, k + 1)))
    tileid_instructions = ItemTuple(ind_masks, instructions)
    tileid_instructions.label_stack = stack
    tileid_instructions.plan = engine.compile_cached(instructions, cp_measure_kwargs)
    tileid_instructions.counts = counts
    extra = {}
    if cp_measure_kwargs is not None:
        extra["cp_measure_kwargs"] = cp_measure_kwargs
    result = measure_fn(tileid_instructions, masks, pixels, ncores=ncores, progress_bar=progress_bar, **extra)
    return tileid_instructions, result


def process_tree_masks_overlap(
    tree: dict,
    masks,
    pixels: np.ndarray,
    measure_fn,
    ncores=None,
    progress_bar: bool = False,
    overlap: bool = True,
    cp_measure_kwargs=None,
    original_ids: bool = False,
):
    """BABY-style ``(tile, stack, label)`` enumeration (extract.py:456-517).

    Default (``original_ids=False``) = the live reference path: the ids enumerated for a stack are ``1..k``
    with ``k`` the number of distinct non-zero labels in that stack (``relabel_sequential`` ids), while the
    measurement reads the plane of the id itself in the ORIGINAL labelling; for
    sequential labels — the supported case of the reference — both coincide.

    ``original_ids=True`` = what the reference's (uncalled) ``format_extraction_overlap`` was written for
    (extract.py:602-682): sequential id ``j`` measures the object whose original id is the ``j``-th smallest
    of its stack, and a third element ``inverse_mappings[(tile, stack)][j] -> original id`` is returned for
    :func:`format_extraction_overlap`."""
    masks = _as_mask_list(masks)
    instructions = kv(flatten(tree))
    tile_stack_mask = []
    inverse_mappings = {}
    counts = {}
    for tile_i, masks_in_tile in enumerate(masks):
        for stack_i, stack_pixels in enumerate(masks_in_tile):
            ids = np.unique(stack_pixels)
            ids = ids[ids > 0]
            counts[(tile_i, stack_i)] = len(ids)
            inverse_mappings[(tile_i, stack_i)] = np.concatenate([[0], ids]).astype(np.int64)
            tile_stack_mask.extend((tile_i, stack_i, mask_i) for mask_i in range(1, len(ids) + 1))
    tileid_instructions = ItemTuple(tile_stack_mask, instructions)
    tileid_instructions.plan = engine.compile_cached(instructions, cp_measure_kwargs)
    tileid_instructions.counts = counts
    extra = {}
    if cp_measure_kwargs is not None:
        extra["cp_measure_kwargs"] = cp_measure_kwargs
    if original_ids:
        extra["inverse_mappings"] = inverse_mappings
    result = measure_fn(tileid_instructions, masks, pixels, ncores=ncores, progress_bar=progress_bar, **extra)
    if original_ids:
        return tileid_instructions, result, inverse_mappings
    return tileid_instructions, result


def extract_tree(
    tileid_instructions,
    masks,
    pixels,
    ncores=False,
    progress_bar: bool = False,
    overlap: bool = False,
    cp_measure_kwargs=None,
    inverse_mappings=None,
):
    """All measurements of ``tileid_instructions`` in one pass on the GPU (extract.py:304-375).

    ``tileid_instructions`` may be any subset/order of ``((tile, label), (ch, red, metric))``
    items (``(tile, stack, label)`` with ``overlap=True``)."""
    results = ExtractionResults()
    if not len(tileid_instructions):
        return results
    masks = _as_mask_list(masks)
    counts = None
    if isinstance(tileid_instructions, ItemTuple) and tileid_instructions.plan is not None:
        plan, objects = tileid_instructions.plan, tileid_instructions.objects
        n_inst = len(plan.instructions)
        row_of_item = inst_of_item = None  # object-major product: implicit
        if inverse_mappings is None:
            counts = tileid_instructions.counts  # ids 1..k of every plane, planes in this function's own order
    else:
        inst_index: dict = {}
        obj_index: dict = {}
        row_of_item = np.empty(len(tileid_instructions), dtype=np.int64)
        inst_of_item = np.empty(len(tileid_instructions), dtype=np.int64)
        for k, (obj, inst) in enumerate(tileid_instructions):
            row_of_item[k] = obj_index.setdefault(tuple(obj), len(obj_index))
            inst_of_item[k] = inst_index.setdefault(tuple(inst), len(inst_index))
        plan = engine.compile_cached(list(inst_index), cp_measure_kwargs)
        objects = list(obj_index)
        n_inst = len(inst_index)

    # label planes: one per tile, or one per (tile, stack) for overlapping masks
    planes, plane_tile, n_labels, plane_of = [], [], [], {}
    stack = getattr(tileid_instructions, "label_stack", None) if counts is not None and not overlap else None
    if stack is not None and len(stack) == len(counts) == len(masks):
        # process_tree_masks stacked every tile and counted its ids: nothing to look at again
        plane_of = dict(zip(counts, range(len(counts))))
        plane_tile = list(range(len(masks)))
        n_labels = list(counts.values())
        planes = stack
    else:
        stack = None
    for tile_i, m in enumerate(masks if stack is None else ()):
        if not len(m):
            continue
        stacks = list(m) if overlap else [m]
        for stack_i, plane in enumerate(stacks):
            plane = np.asarray(plane)
            key = (tile_i, stack_i) if overlap else (tile_i,)
            plane_of[key] = len(planes)
            planes.append(plane)
            plane_tile.append(tile_i)
            if counts is not None:  # process_tree_masks* has looked at this plane already
                n_labels.append(counts[key])
            elif overlap and inverse_mappings is None:  # ids 1..k, k = number of distinct labels of the stack
                n_labels.append(int(np.count_nonzero(np.unique(plane))))
            else:
                n_labels.append(int(plane.max()) if plane.size else 0)
    n_labels = np.asarray(n_labels, dtype=np.int64)
    base = np.concatenate([[0], np.cumsum(n_labels)])
    # (overlap without original ids: the reference enumerates 1..k of every stack and ignores larger ids, extract.py:478-500)
    dense = _run_dense(plan, planes, np.asarray(plane_tile, dtype=np.int32), n_labels, pixels,
                       strict_labels=not (overlap and inverse_mappings is None))

    if counts is not None:
        # the objects are the ids 1..k of every plane in plane order: row = position, no per-object Python work
        obj_rows = np.arange(len(objects), dtype=np.int64)
        keys = np.asarray(list(plane_of), dtype=np.int64).reshape(len(plane_of), -1)
        starts = base[:-1]
        objects_arr = np.concatenate(
            [np.repeat(keys, n_labels, axis=0), (obj_rows - np.repeat(starts, n_labels) + 1)[:, None]], axis=1)
    else:
        objects_arr = None
        obj_rows = np.empty(len(objects), dtype=np.int64)
        for k, obj in enumerate(objects):
            p = plane_of[tuple(obj[:-1])]
            label = obj[-1]
            if inverse_mappings is not None:  # sequential id -> original id of its (tile, stack)
                label = int(inverse_mappings[tuple(obj[:-1])][label])
            if not (1 <= label <= n_labels[p]):
                raise IndexError(f"index {label - 1} is out of bounds for axis 0 with size {n_labels[p]}")
            obj_rows[k] = base[p] + label - 1
    if row_of_item is None:
        row_of_item = obj_rows  # (object-major product: _results_from_dense expands it)
    else:
        row_of_item = obj_rows[row_of_item]
    results.extend(_results_from_dense(plan, dense, row_of_item, inst_of_item))
    results.dense, results.plan, results.items = dense, plan, tileid_instructions
    results.objects = objects_arr if objects_arr is not None else np.asarray(objects, dtype=np.int64).reshape(len(objects), -1)
    results.obj_rows = obj_rows
    return results


def extract_tree_multi(
    tileid_instructions,
    masks,
    pixels,
    ncores=None,
    progress_bar: bool = False,
    cp_measure_kwargs=None,
):
    """Two-channel measurements of an ``extractmulti_*`` step in one pass on the GPU (extract.py:378-453).

    Items are ``((tile, label), ((ch0, ch1), red_ch, red_z, metric))`` (extract.py:222); ``red_ch`` must be ``"None"``
    (the two-image branch of ``measure_multi``, extract.py:223-226) and ``metric`` one of ``pearson``, ``manders_fold``,
    ``rwc``, ``overlap`` — CellProfiler MeasureColocalization for one object at a time, what ``wrap_cp_corr_features``
    (loaders.py:153-168) obtains from cp_measure; results are dicts ``{key: ndarray(1)}`` like theirs.  ``costes`` has
    no kernel: :func:`aliby_b200.pipe.init_step` splits such a tree between this function and the reference."""
    assert isinstance(masks, list) or masks.ndim >= 3, "Masks dimensions < 2. It should include batch/tile dimension."
    return extract_tree(tileid_instructions, masks, pixels, ncores=ncores, progress_bar=progress_bar,
                        cp_measure_kwargs=cp_measure_kwargs)


def format_extraction(instructions_result):
    """Long -> wide ``pyarrow.Table`` with the reference's naming (extract.py:520-599):
    scalar results land in ``"{ch}/{red}/{metric}/{metric}"``, dict results in
    ``"{ch}/{red}/{metric}/{key}"``, ndarrays (embedders) in ``X_{c}``; columns are
    ``tile, label, <sorted metric names>``; missing cells are null."""
    import pyarrow as pa

    instructions, results = instructions_result
    if isinstance(results, ExtractionResults) and results.items is instructions and results.dense is not None:
        table = _format_dense(results, pa)
        if table is not None:
            return table
    names = ("tile", "label", "metric", "value")
    formatted = {k: [] for k in names}
    for inst, metrics in zip(instructions, results, strict=True):
        tileid = inst[0][0]
        label = inst[0][-1]
        branch = "/".join(str(x) for x in inst[1])
        if isinstance(metrics, (int, float)):
            formatted["tile"].append(tileid)
            formatted["label"].append(label)
            formatted["metric"].append(f"{branch}/{inst[1][-1]}")
            formatted["value"].append(metrics)
        elif isinstance(metrics, dict):
            for k, values in metrics.items():
                for value in values:
                    formatted["value"].append(value)
                    formatted["tile"].append(tileid)
                    formatted["label"].append(label)
                    formatted["metric"].append(f"{branch}/{k}")
        elif isinstance(metrics, np.ndarray):
            for (r, c), value in np.ndenumerate(metrics):
                formatted["tile"].append(r)
                formatted["label"].append(0)
                formatted["metric"].append(f"X_{c}")
                formatted["value"].append(value)
        else:
            raise Exception(
                f"the metrics are in an invalid value: {type(metrics)}. Valid values are int/float, dict or numpy array."
            )
    pivoted: dict = {}
    for t, lbl, m, v in zip(formatted["tile"], formatted["label"], formatted["metric"], formatted["value"], strict=True):
        pivoted.setdefault((t, lbl), {"tile": t, "label": lbl})[m] = v
    metrics_list = sorted(set(formatted["metric"]))
    wide = {"tile": [], "label": []}
    wide.update({m: [] for m in metrics_list})
    for row in pivoted.values():
        wide["tile"].append(row["tile"])
        wide["label"].append(row["label"])
        for m in metrics_list:
            wide[m].append(row.get(m, None))
    return pa.Table.from_pydict(wide)


def format_extraction_overlap(instructions_result):
    """``(instructions, results, inverse_mappings)`` -> wide table keyed by the ORIGINAL label ids
    (extract.py:602-682): like :func:`format_extraction`, with ``label = inverse_mappings[(tile, stack)][label]``
    and the columns already renamed to ``metadata_tile`` / ``metadata_label``."""
    import pyarrow as pa

    instructions, results, inverse_mappings = instructions_result
    names = ("tile", "label", "metric", "value")
    formatted = {k: [] for k in names}
    for inst, metrics in zip(instructions, results, strict=True):
        tileid, stack_id, label = inst[0]
        branch = "/".join(str(x) for x in inst[1])
        original = int(inverse_mappings[tileid, stack_id][label])
        if isinstance(metrics, (int, float)):
            rows = [(f"{branch}/{inst[1][-1]}", metrics)]
        elif isinstance(metrics, dict):
            rows = [(f"{branch}/{k}", value) for k, values in metrics.items() for value in values]
        elif isinstance(metrics, list):
            rows = [(f"{branch}/{inst[1][-1]}", value) for value in metrics]
        else:
            rows = []
        for name, value in rows:
            formatted["tile"].append(tileid)
            formatted["label"].append(original)
            formatted["metric"].append(name)
            formatted["value"].append(value)
    pivoted: dict = {}
    for t, lbl, m, v in zip(formatted["tile"], formatted["label"], formatted["metric"], formatted["value"], strict=True):
        pivoted.setdefault((t, lbl), {"tile": t, "label": lbl})[m] = v
    metrics_list = sorted(set(formatted["metric"]))
    wide = {"tile": [], "label": []}
    wide.update({m: [] for m in metrics_list})
    for row in pivoted.values():
        wide["tile"].append(row["tile"])
        wide["label"].append(row["label"])
        for m in metrics_list:
            wide[m].append(row.get(m, None))
    table = pa.Table.from_pydict(wide)
    return table.rename_columns([{"tile": "metadata_tile", "label": "metadata_label"}.get(c, c) for c in table.column_names])


def _format_dense(results: ExtractionResults, pa):
    """Arrow table straight from the dense block (no per-value Python work)."""
    plan = results.plan
    if any(len(c) != 1 and k is None for c, k in zip(plan.inst_cols, plan.inst_keys)):
        return None  # tuple-valued metric: let the generic path raise like the reference
    objs = results.objects
    keys = objs[:, [0, -1]]
    # the reference keys rows by (tile, label): a later object with the same key overwrites
    _, first = np.unique(keys, axis=0, return_index=True)
    if len(first) != len(keys):
        return None
    names = {}
    for inst, cols, dict_keys in zip(plan.instructions, plan.inst_cols, plan.inst_keys):
        branch = "/".join(str(x) for x in inst)
        if dict_keys is None:
            names[f"{branch}/{inst[-1]}"] = cols[0]
        else:  # dict-valued metric: one column per key (extract.py:553-562)
            for k, j in zip(dict_keys, cols):
                names[f"{branch}/{k}"] = j
    data = {"tile": pa.array(keys[:, 0], pa.int64()), "label": pa.array(keys[:, 1], pa.int64())}
    block = results.dense[results.obj_rows]
    for name in sorted(names):
        data[name] = pa.array(np.ascontiguousarray(block[:, names[name]]), pa.float64())
    return pa.table(data)
