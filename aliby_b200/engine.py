"""Host side of the extraction hot path: tree -> plan, plan + device buffers -> dense table.

The plan compiler mirrors how the reference resolves a tree
(``src/extraction/extract.py:33-74`` flatten/kv, ``extract.py:147-153`` registry lookups,
``distributors.py:19-24`` reducer check) and reproduces its error behaviour; the engine
is a thin wrapper over the C-ABI (``abx_extract``).  PyTorch is used for device memory
and streams only.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from functools import reduce
from itertools import product

import numpy as np

from . import _native as nat

# REDUCTION_FUNS of loaders.py:110-127: which names exist, which are ufuncs
REDUCERS = {"max": "ufunc", "add": "ufunc", "div": "ufunc", "mean": "<function mean>", "median": "<function median>", "None": "None"}

# cell.py functions that take the mask only (wrapped by ignore_pixels, loaders.py:56-66,170-171)
SHAPE_METRICS = {
    "area": ("area",),
    "centroid": ("centroid_x", "centroid_y"),
    "centroid_x": ("centroid_x",),
    "centroid_y": ("centroid_y",),
    "conical_volume": ("conical_volume",),
    "eccentricity": ("eccentricity",),
    "min_maj_approximation": ("minor_axis", "major_axis"),
    "spherical_volume": ("spherical_volume",),
    "volume": ("volume",),
    # extensions (not in cell.py; cp_measure territory in the reference)
    "bbox_rmin": ("bbox_rmin",),
    "bbox_rmax": ("bbox_rmax",),
    "bbox_cmin": ("bbox_cmin",),
    "bbox_cmax": ("bbox_cmax",),
}
# cell.py functions of (mask, pixels) + extensions max/min + the per-tile background pair of trap.py
INTENSITY_METRICS = {
    "mean": 0,
    "total": nat.F_WRAPSQ * 0,
    "total_squared": nat.F_WRAPSQ,
    "std": 0,
    "median": nat.F_MEDIAN,
    "max2p5pc": nat.F_TOP2P5,
    "max5px_median": nat.F_TOP5 | nat.F_MEDIAN,
    "moment_of_inertia": nat.F_MOI,
    "ratio": 0,
    "max": 0,
    "min": 0,
}
BACKGROUND_METRICS = {"imBackground": nat.F_MEDIAN, "background_max5": nat.F_TOP5}

# cp_measure features with a kernel (loaders.py:71-77,135-150): dict-valued metrics, one table column per key.
# `intensity` = CellProfiler MeasureObjectIntensity WITHOUT the edge features (the caller has to switch those off with
# cp_measure_kwargs={"intensity": {"edge_measurements": False}}, pipe_builder.py:84-91); `sizeshape` = the subset of
# MeasureObjectSizeShape that the label scan, the first EDT and the coordinate moments give.  Parity is self-defined
# against oracle/cpm.py (cp_measure's source is not available: SURVEY.md 8c).
CP_INTENSITY = (  # (key, building block, request features)
    ("Intensity_IntegratedIntensity", "total", 0),
    ("Intensity_MeanIntensity", "mean", 0),
    ("Intensity_StdIntensity", "std", 0),
    ("Intensity_MinIntensity", "min", 0),
    ("Intensity_MaxIntensity", "max", 0),
    ("Intensity_MassDisplacement", "cp_mass_displacement", nat.F_MOI),
    ("Intensity_LowerQuartileIntensity", "cp_lower_quartile", nat.F_CPQ),
    ("Intensity_MedianIntensity", "cp_median", nat.F_CPQ),
    ("Intensity_MADIntensity", "cp_mad", nat.F_CPQ | nat.F_CPMAD),
    ("Intensity_UpperQuartileIntensity", "cp_upper_quartile", nat.F_CPQ),
    ("Location_CenterMassIntensity_X", "cp_center_mass_x", nat.F_MOI),
    ("Location_CenterMassIntensity_Y", "cp_center_mass_y", nat.F_MOI),
    ("Location_CenterMassIntensity_Z", "cp_center_mass_z", 0),
    ("Location_MaxIntensity_X", "cp_max_pos_x", nat.F_CPQ | nat.F_CPMAD),
    ("Location_MaxIntensity_Y", "cp_max_pos_y", nat.F_CPQ | nat.F_CPMAD),
    ("Location_MaxIntensity_Z", "cp_zero", 0),
)
CP_SIZESHAPE = (  # (key, building block)
    ("AreaShape_Area", "area"),
    ("AreaShape_BoundingBoxArea", "cp_bbox_area"),
    ("AreaShape_BoundingBoxMaximum_X", "cp_bbox_max_x"),
    ("AreaShape_BoundingBoxMaximum_Y", "cp_bbox_max_y"),
    ("AreaShape_BoundingBoxMinimum_X", "bbox_cmin"),
    ("AreaShape_BoundingBoxMinimum_Y", "bbox_rmin"),
    ("AreaShape_Center_X", "cp_center_x"),
    ("AreaShape_Center_Y", "cp_center_y"),
    ("AreaShape_Eccentricity", "cp_eccentricity"),
    ("AreaShape_EquivalentDiameter", "cp_equivalent_diameter"),
    ("AreaShape_Extent", "cp_extent"),
    ("AreaShape_MajorAxisLength", "cp_major_axis_length"),
    ("AreaShape_MaximumRadius", "cp_maximum_radius"),
    ("AreaShape_MeanRadius", "cp_mean_radius"),
    ("AreaShape_MinorAxisLength", "cp_minor_axis_length"),
)
CP_FEATURE_NAMES = ("intensity", "sizeshape")
# cp_measure two-image features of `extractmulti_*` trees (loaders.py:75-77,153-168; extract.py:200-237) with a kernel:
# (key, building block) per feature and the pair features they need.  CellProfiler MeasureColocalization for objects,
# self-defined against oracle/cpm.py like the features above; the keys are CellProfiler's feature stems.
CP_CORRELATION = {
    "pearson": ((("Correlation_Pearson", "co_pearson"),), 0),
    "manders_fold": ((("Correlation_Manders_1", "co_manders_1"), ("Correlation_Manders_2", "co_manders_2")), nat.PF_THRESHOLDED),
    "rwc": ((("Correlation_RWC_1", "co_rwc_1"), ("Correlation_RWC_2", "co_rwc_2")), nat.PF_THRESHOLDED | nat.PF_RWC),
    "overlap": ((("Correlation_Overlap", "co_overlap"), ("Correlation_K_1", "co_k_1"), ("Correlation_K_2", "co_k_2")),
                nat.PF_THRESHOLDED),
}
CP_CORRELATION_WITHOUT_KERNEL = ("costes",)  # known to the reference's registry, no CUDA kernel here

CELL_FUN_NAMES = tuple(
    sorted(
        [
            "area", "centroid", "centroid_x", "centroid_y", "conical_volume", "eccentricity", "max2p5pc",
            "max5px_median", "mean", "median", "min_maj_approximation", "moment_of_inertia", "ratio",
            "spherical_volume", "std", "total", "total_squared", "volume",
        ]
    )
)
EXTENSION_NAMES = ("bbox_cmax", "bbox_cmin", "bbox_rmax", "bbox_rmin", "max", "min")
TRAP_FUN_NAMES = ("background_max5", "imBackground")


def flatten(d: dict, pref=()) -> dict:
    """Nested dict -> {path tuple: leaf}, insertion order kept (extract.py:33-57)."""
    return reduce(
        lambda acc, kv_: ({**acc, **flatten(kv_[1], (*pref, kv_[0]))} if isinstance(kv_[1], dict) else {**acc, (*pref, kv_[0]): kv_[1]}),
        d.items(),
        {},
    )


def kv(flat: dict) -> list:
    """{(ch, red): [metrics]} -> [(ch, red, metric)] (extract.py:60-74)."""
    return [(*k1, v1) for k, v in flat.items() for k1, v1 in product((k,), v)]


@dataclass
class Plan:
    """A compiled extraction tree."""

    instructions: list
    requests: list = field(default_factory=list)  # [(channel, red_enum, features, bg_features)]
    columns: list = field(default_factory=list)  # [(request_idx, metric_enum)] (pair index for the two-image metrics)
    pairs: list = field(default_factory=list)  # [(request_a, request_b, pair features, threshold fraction)]
    inst_cols: list = field(default_factory=list)  # per instruction: tuple of dense column indices
    inst_keys: list = field(default_factory=list)  # per instruction: None, or the dict keys of a dict-valued metric
    need_edt: int = 0  # bit 0: axes (eccentricity/volume/min/maj), bit 1: conical_volume
    with_background: bool = False
    error: Exception | None = None  # raised only when there is at least one object, like the reference
    _dev: dict = field(default_factory=dict)

    @property
    def n_columns(self) -> int:
        return len(self.columns)

    @property
    def max_channel(self) -> int:
        return max((r[0] for r in self.requests), default=-1)

    def device_arrays(self, device):
        """(requests, columns) as device tensors, uploaded once per device."""
        import torch

        key = str(device)
        if key not in self._dev:
            req = np.zeros((max(1, len(self.requests)), 4), dtype=np.int32)
            for i, r in enumerate(self.requests):
                req[i] = r
            col = np.zeros((max(1, len(self.columns)), 2), dtype=np.int32)
            for i, c in enumerate(self.columns):
                col[i] = c
            self._dev[key] = (torch.from_numpy(req).to(device), torch.from_numpy(col).to(device))
        return self._dev[key]

    def device_pairs(self, device):
        """``abx_pair[]`` as a device tensor (None without two-image metrics)."""
        import torch

        if not self.pairs:
            return None
        key = ("pairs", str(device))
        if key not in self._dev:
            arr = (nat.Pair * len(self.pairs))()
            for i, (ra, rb, feats, frac) in enumerate(self.pairs):
                arr[i].request_a, arr[i].request_b, arr[i].features, arr[i].threshold_fraction = ra, rb, feats, frac
            self._dev[key] = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        return self._dev[key]


def compile_tree(tree: dict, cp_measure_kwargs=None) -> Plan:
    return compile_instructions(kv(flatten(tree)), cp_measure_kwargs)


_plan_cache: dict = {}
_plan_cache_lock = __import__("threading").Lock()


def compile_cached(instructions: list, cp_measure_kwargs=None) -> Plan:
    """:func:`compile_instructions` memoised on the instruction list: a pipeline calls its extract step with the same tree at
    every time point, and a cached plan also keeps its device copies of the request / column / pair tables."""
    try:
        key = (tuple(instructions), repr(sorted((k, sorted(v.items())) for k, v in (cp_measure_kwargs or {}).items())))
        plan = _plan_cache.get(key)
    except TypeError:  # an unhashable instruction: compile it for the error message it deserves
        return compile_instructions(instructions, cp_measure_kwargs)
    if plan is None:
        plan = compile_instructions(instructions, cp_measure_kwargs)
        with _plan_cache_lock:  # (two threads may both compile a new tree; the cache keeps one of the plans)
            if len(_plan_cache) >= 64:
                _plan_cache.pop(next(iter(_plan_cache)))
            plan = _plan_cache.setdefault(key, plan)
    return plan


def compile_instructions(instructions: list, cp_measure_kwargs=None) -> Plan:
    plan = Plan(instructions=list(instructions))
    cp_kw = dict(cp_measure_kwargs or {})
    req_index: dict = {}
    col_index: dict = {}

    def request(ch, red):
        key = (int(ch), red)
        if key not in req_index:
            req_index[key] = len(plan.requests)
            plan.requests.append([int(ch), {"max": nat.RED_MAX, "add": nat.RED_ADD, "div": nat.RED_DIV}[red], 0, 0])
        return req_index[key]

    def column(req, metric_name):
        key = (req, nat.METRIC[metric_name])
        if key not in col_index:
            col_index[key] = len(plan.columns)
            plan.columns.append(key)
            if key[1] in nat.EDT_METRICS:
                plan.need_edt |= 1
            if key[1] in nat.CONICAL_METRICS:
                plan.need_edt |= 2
            if key[1] in nat.MOMENT_METRICS:
                plan.need_edt |= 4
        return col_index[key]

    pair_index: dict = {}

    def multi(inst):
        """((ch0, ch1), red_ch, red_z, metric) of an extractmulti tree (extract.py:200-237)."""
        chs, red_ch, red, metric = inst
        ch0, ch1 = chs
        if red_ch not in REDUCERS or red not in REDUCERS:
            raise KeyError(red if red_ch in REDUCERS else red_ch)
        if metric in CP_CORRELATION_WITHOUT_KERNEL:
            raise NotImplementedError(f"two-image feature '{metric}' has no CUDA kernel in aliby_b200 and there is no CPU fallback")
        if red_ch != "None":
            # the reference's other branch combines the two channels and calls measure_mono without its registries
            # (extract.py:227-235): a TypeError there
            raise NotImplementedError("extractmulti branches with a channel reduction (red_ch != 'None') have no CUDA kernel")
        if metric not in CP_CORRELATION:
            raise KeyError(metric)
        if REDUCERS[red] != "ufunc":
            raise Exception(f"{REDUCERS[red]} is an invalid reducer.")  # distributors.py:24
        if red == "div":
            raise NotImplementedError("two-image features of `div`-reduced (floating point) stacks have no CUDA kernel")
        blocks, feats = CP_CORRELATION[metric]
        thr = cp_kw.get(metric, {}).get("thr", 15)
        key = (request(ch0, red), request(ch1, red), thr)
        if key not in pair_index:
            pair_index[key] = len(plan.pairs)
            plan.pairs.append([key[0], key[1], 0, thr / 100])
        p = pair_index[key]
        plan.pairs[p][2] |= feats
        return tuple(column(p, block) for _, block in blocks), [k for k, _ in blocks]

    for inst in instructions:
        try:
            if len(inst) == 4:
                cols, keys = multi(inst)
                plan.inst_cols.append(cols)
                plan.inst_keys.append(keys)
                continue
            ch, red, metric = inst
            if red not in REDUCERS:
                raise KeyError(red)  # REDUCTION_FUNS[red_z], extract.py:151
            known = (metric in SHAPE_METRICS or metric in INTENSITY_METRICS or metric in BACKGROUND_METRICS
                     or metric in CP_FEATURE_NAMES)
            if not known:
                raise KeyError(metric)  # CELL_FUNS[metric], extract.py:152
            has_pixels = not (isinstance(ch, str) and ch == "None")
            if has_pixels:
                if REDUCERS[red] != "ufunc":
                    raise Exception(f"{REDUCERS[red]} is an invalid reducer.")  # distributors.py:24
            keys = None
            if metric == "sizeshape":
                cols = tuple(column(-1, block) for _, block in CP_SIZESHAPE)
                keys = [k for k, _ in CP_SIZESHAPE]
            elif metric == "intensity":
                if not has_pixels:
                    raise TypeError("metric 'intensity' needs pixels but the channel is 'None'")
                if cp_kw.get("intensity", {}).get("edge_measurements", True):
                    raise NotImplementedError(
                        "cp_measure 'intensity' with edge_measurements=True (the *Edge features need the object's outline) has no "
                        "CUDA kernel: pass cp_measure_kwargs={'intensity': {'edge_measurements': False}}"
                    )
                r = request(ch, red)
                for _, _, feats in CP_INTENSITY:
                    plan.requests[r][2] |= feats
                cols = tuple(column(r, block) for _, block, _ in CP_INTENSITY)
                keys = [k for k, _, _ in CP_INTENSITY]
            elif metric in SHAPE_METRICS:
                cols = tuple(column(-1, m) for m in SHAPE_METRICS[metric])
            else:
                if not has_pixels:
                    raise TypeError(f"metric '{metric}' needs pixels but the channel is 'None'")
                r = request(ch, red)
                if metric in BACKGROUND_METRICS:
                    plan.requests[r][3] |= BACKGROUND_METRICS[metric]
                    plan.with_background = True
                else:
                    plan.requests[r][2] |= INTENSITY_METRICS[metric]
                cols = (column(r, metric),)
            plan.inst_cols.append(cols)
            plan.inst_keys.append(keys)
        except Exception as e:  # noqa: BLE001 - deferred, see Plan.error
            if plan.error is None:
                plan.error = e
            plan.inst_cols.append(())
            plan.inst_keys.append(None)
    return plan


_DTYPES = {"uint8": nat.U8, "uint16": nat.U16, "float32": nat.F32, "float64": nat.F64}
PIXEL_DTYPES = tuple(_DTYPES)  # numpy / torch dtype names with a kernel


def alloc_table(n_objects: int, n_cols: int, device, n_status: int = 1):
    """Device table plus status words in ONE buffer, so that a single device-to-host copy returns both.

    Returns ``(buf, table, status)``: ``buf`` float64 ``[n_objects * n_cols + n_status]``, ``table`` its first part as
    ``(n_objects, n_cols)``, ``status`` an int32 view of the last ``n_status`` slots (two int32 per slot; the library
    writes the first)."""
    import torch

    buf = torch.empty(n_objects * n_cols + n_status, dtype=torch.float64, device=device)
    status = buf[n_objects * n_cols :].view(torch.int32)
    status.zero_()
    return buf, buf[: n_objects * n_cols].view(n_objects, n_cols), status


def raise_on_status(word: int) -> None:
    """Error flags of a finished call (``abx_extract_args.status``)."""
    if int(word) & 1:
        raise IndexError(
            "a label id exceeds the number of labels given for its plane (n_labels / plane_base): the masks changed "
            "after their maximum was taken, or n_labels is stale — the table would be missing those pixels"
        )
    if int(word) & 2:
        raise NotImplementedError(
            "cp_measure 'intensity' was requested under the `div` reducer or the `add` reducer of a Z stack: the kernels "
            "that serve those have no quartiles / MAD / maximum position and there is no CPU fallback"
        )
    if int(word) & 4:
        raise NotImplementedError(
            "a two-image feature met values of 65536 or more (the `add` reduction of a Z stack): no CUDA kernel covers "
            "them and there is no CPU fallback"
        )


_meta_cache: dict = {}


def _device_meta(device, plane_tile: np.ndarray, plane_base: np.ndarray, tile_offset: np.ndarray):
    """tile_offset (i64) | plane_tile (i32) | plane_base (i32) as ONE device buffer, uploaded with one
    asynchronous copy from pinned staging and cached by content: a pipeline calls the extract step with the
    same tile table at every time point (and the bench with identical arguments at every step), so the usual
    call uploads nothing and the host never blocks."""
    import torch

    blob = tile_offset.tobytes() + plane_tile.tobytes() + plane_base.tobytes()
    key = (str(device), blob)
    hit = _meta_cache.get(key)
    if hit is not None:
        # (inside a graph capture neither a query nor a wait on an outside event is legal; GraphedExtract has
        # synchronised with the upload before it captures)
        if not torch.cuda.is_current_stream_capturing() and not hit._abx_ready.query():
            torch.cuda.current_stream(device).wait_event(hit._abx_ready)  # no-op on the uploading stream
        return hit
    if len(_meta_cache) >= 64:
        _meta_cache.pop(next(iter(_meta_cache)))
    host = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
    dev = torch.empty(len(blob), dtype=torch.uint8, device=device)
    dev.copy_(host, non_blocking=True)
    _meta_cache[key] = dev
    dev._abx_host = host  # keep the pinned staging alive until the copy has certainly run
    dev._abx_ready = torch.cuda.Event()
    dev._abx_ready.record(torch.cuda.current_stream(device))
    return dev


def pixel_dtype_enum(torch_dtype) -> int:
    name = str(torch_dtype).replace("torch.", "")
    if name not in _DTYPES:
        raise NotImplementedError(
            f"pixel dtype {name} has no CUDA kernel in aliby_b200 (uint8/uint16/float32/float64) and there is no CPU fallback"
        )
    return _DTYPES[name]


def run_planes(
    plan: Plan,
    labels,  # torch cuda uint16 (P, H, W), last dim contiguous
    plane_tile: np.ndarray,
    n_labels: np.ndarray,
    pixels,  # torch cuda tensor holding the pixel data (any shape; addressed through offsets/strides)
    tile_offset: np.ndarray,
    chan_stride: int,
    z_stride: int,
    row_stride: int,
    n_channels: int,
    n_z: int,
    out=None,
    stage_events=None,
    status=None,
):
    """Launch the hot path on the current stream; returns the dense fp64 table (device).

    ``status``: optional int32 device tensor whose first element receives the call's error flags (stream-ordered;
    check it with :func:`raise_on_status` once it is on the host).  Re-entrant: the scratch memory of a call comes from
    the caching allocator on the current stream, nothing is shared between calls in flight."""
    import torch

    if plan.error is not None and int(np.sum(n_labels)) > 0:
        raise plan.error
    lib = nat.lib()
    device = labels.device
    P, H, W = labels.shape
    assert labels.dtype == torch.uint16 and labels.stride(2) == 1
    n_labels = np.asarray(n_labels, dtype=np.int64)
    n_objects = int(n_labels.sum())
    n_cols = plan.n_columns
    if out is None:
        out = torch.empty((n_objects, n_cols), dtype=torch.float64, device=device)
    if n_objects == 0 or n_cols == 0:
        return out
    if plan.requests and plan.max_channel >= n_channels:
        raise IndexError(f"index {plan.max_channel} is out of bounds for axis 1 with size {n_channels}")
    plane_tile = np.asarray(plane_tile, dtype=np.int32)
    tile_offset = np.ascontiguousarray(tile_offset, dtype=np.int64)
    if plan.requests:
        # every label plane must lie inside the pixel window of its tile (the reference indexes pixels[tile][mask] and
        # raises IndexError when the shapes differ; the kernels would read neighbouring tiles)
        if len(plane_tile) and (plane_tile.min() < 0 or plane_tile.max() >= len(tile_offset)):
            raise IndexError(f"tile index {int(plane_tile.max())} is out of bounds for {len(tile_offset)} pixel tiles")
        elems = (pixels.untyped_storage().nbytes() // pixels.element_size()) - pixels.storage_offset()
        last = int(tile_offset.max()) + (max(1, n_channels) - 1) * int(chan_stride) + (max(1, n_z) - 1) * int(z_stride) \
            + (H - 1) * int(row_stride) + W
        if int(tile_offset.min()) < 0 or last > elems or W > int(row_stride):
            raise IndexError(
                f"label planes of {H} x {W} do not fit the pixel tiles (row stride {int(row_stride)}, last element "
                f"{last} of {elems}): masks and pixels must have the same (Y, X)"
            )
    base = np.zeros(P + 1, dtype=np.int32)
    np.cumsum(n_labels, out=base[1:])
    meta = _device_meta(device, plane_tile, base, tile_offset)
    req_t, col_t = plan.device_arrays(device)

    a = nat.ExtractArgs()
    a.labels = labels.data_ptr()
    a.label_dtype = nat.U16
    a.n_planes, a.H, a.W = P, H, W
    a.label_plane_stride, a.label_row_stride = labels.stride(0), labels.stride(1)
    n_tiles = len(tile_offset)
    a.tile_offset = meta.data_ptr()  # int64 [n_tiles] first (8-byte aligned), then the two int32 arrays
    a.plane_tile = meta.data_ptr() + 8 * n_tiles
    a.plane_base = meta.data_ptr() + 8 * n_tiles + 4 * P
    a.n_objects = n_objects
    a.with_background = int(plan.with_background)
    a.pixels = pixels.data_ptr() if plan.requests else None
    a.pixel_dtype = pixel_dtype_enum(pixels.dtype) if plan.requests else nat.U16
    a.n_tiles = n_tiles
    a.C, a.Z = max(1, n_channels), max(1, n_z)
    a.chan_stride, a.z_stride, a.row_stride = int(chan_stride), int(z_stride), int(row_stride)
    a.requests = req_t.data_ptr()
    a.n_requests = len(plan.requests)
    a.columns = col_t.data_ptr()
    a.n_columns = n_cols
    a.need_edt = int(plan.need_edt)
    a.request_feature_union = nat.F_HAS_DIV if any(r[1] == nat.RED_DIV for r in plan.requests) else 0
    for r in plan.requests:
        a.request_feature_union |= int(r[2]) | int(r[3])
    if a.request_feature_union & (nat.F_CPQ | nat.F_CPMAD):
        if a.pixel_dtype not in (nat.U8, nat.U16):
            raise NotImplementedError("cp_measure 'intensity' has a CUDA kernel for uint8/uint16 pixels only; there is no CPU fallback")
        for r in plan.requests:
            if int(r[2]) & (nat.F_CPQ | nat.F_CPMAD) and (r[1] == nat.RED_DIV or (r[1] == nat.RED_ADD and a.Z > 1)):
                raise NotImplementedError(
                    "cp_measure 'intensity' of a Z stack has a CUDA kernel for the `max` reduction only (`add` sums and "
                    "`div` quotients take kernels without its rank statistics); there is no CPU fallback")
    if plan.requests:  # extent of the pixel buffer behind data_ptr(), for the TMA description of it
        a.pixel_elems = (pixels.untyped_storage().nbytes() // pixels.element_size()) - pixels.storage_offset()
    pairs_t = plan.device_pairs(device)
    if pairs_t is not None:
        if a.pixel_dtype not in (nat.U8, nat.U16):
            raise NotImplementedError("two-image features have a CUDA kernel for uint8/uint16 pixels only; there is no CPU fallback")
        a.pairs, a.n_pairs = pairs_t.data_ptr(), len(plan.pairs)
    a.table = out.data_ptr()
    a.stream = torch.cuda.current_stream(device).cuda_stream
    if stage_events is not None:  # 6 handles from abx_event_create (bench.py: live per-stage timing)
        ev = (C.c_void_p * 6)(*stage_events)
        a.stage_events = C.cast(ev, C.POINTER(C.c_void_p))
    need = C.c_size_t(0)
    nat.check(lib.abx_extract_workspace_bytes(C.byref(a), C.byref(need)), "abx_extract_workspace_bytes")
    # scratch of this call: from the caching allocator, on the current stream (a block freed here is handed out again
    # only to later work of the same stream, so nothing in flight can be overwritten)
    with torch.cuda.device(device):
        ws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=device)
        a.workspace = ws.data_ptr()
        a.workspace_bytes = ws.numel()
        if status is not None:
            assert status.dtype == torch.int32 and status.device == device
            a.status = status.data_ptr()
        nat.check(lib.abx_extract(C.byref(a)), "abx_extract")
    # meta must outlive the launch: it is kept alive by the per-device cache
    return out


import threading as _threading

# ---- pageable host arrays -> device: through pinned staging filled by a few threads ----
# A pipeline hands NumPy arrays that are pageable; cudaMemcpy of pageable memory goes through the driver's own bounce
# buffer on the calling thread at about 10 GB/s (a 56 MB C2 field: 5 ms of a 7 ms call).  Here the array is cut into
# chunks, a small thread pool copies the chunks into a pinned staging buffer (NumPy releases the GIL for the copy) and
# every chunk goes to the device asynchronously as soon as it is staged.
_UPLOAD_MIN_BYTES = 4 << 20     # below this the plain copy is as fast
_UPLOAD_MAX_BYTES = 256 << 20   # above this the caller's batches should be pinned by the caller
_UPLOAD_CHUNK_BYTES = 4 << 20
_upload_state = _threading.local()


def _upload_pool():
    from concurrent.futures import ThreadPoolExecutor

    pool = getattr(_upload_state, "pool", None)
    if pool is None:
        pool = _upload_state.pool = ThreadPoolExecutor(max_workers=4, thread_name_prefix="abx-upload")
    return pool


def host_to_device(dst, src, slot: str = "pixels"):
    """``dst.copy_(src)`` for a host array ``src`` (ndarray or CPU tensor) and a CUDA tensor ``dst`` of the same shape
    and dtype, asynchronous on the current stream when ``src`` is pinned, staged through pinned memory by a few threads
    when it is pageable.  Staging buffers are per (host thread, slot) and reused; a buffer is not refilled before the
    copies that read it have finished."""
    import torch

    t = src if isinstance(src, torch.Tensor) else torch.from_numpy(src)
    nbytes = t.numel() * t.element_size()
    # (the driver knows whether the memory is page-locked; a tensor made from a NumPy view of a pinned buffer does not)
    if (nbytes < _UPLOAD_MIN_BYTES or nbytes > _UPLOAD_MAX_BYTES or not t.is_contiguous() or not dst.is_contiguous()
            or nat.lib().abx_host_is_pinned(t.data_ptr()) != 0):
        dst.copy_(t, non_blocking=True)
        return
    key = (slot, t.dtype, str(dst.device))
    bufs = getattr(_upload_state, "bufs", None)
    if bufs is None:
        bufs = _upload_state.bufs = {}
    entry = bufs.get(key)
    if entry is None or entry[0].numel() < t.numel():
        entry = bufs[key] = [torch.empty(t.numel(), dtype=t.dtype).pin_memory(), None]
    staging, busy = entry
    if busy is not None:
        busy.synchronize()  # the previous upload out of this buffer
    flat_src, flat_dst, flat_stage = t.reshape(-1).numpy(), dst.reshape(-1), staging[: t.numel()]
    stage_np = flat_stage.numpy()
    per = max(1, _UPLOAD_CHUNK_BYTES // t.element_size())
    bounds = [(i, min(i + per, t.numel())) for i in range(0, t.numel(), per)]
    pool = _upload_pool()
    futures = [pool.submit(np.copyto, stage_np[a:b], flat_src[a:b]) for a, b in bounds]
    for (a, b), fut in zip(bounds, futures):
        fut.result()
        flat_dst[a:b].copy_(flat_stage[a:b], non_blocking=True)
    done = torch.cuda.Event()
    done.record(torch.cuda.current_stream(dst.device))
    entry[1] = done


_capture_lock = _threading.Lock()  # torch's capture context synchronises the device: one capture at a time per process


class GraphedExtract:
    """One extract call of FIXED shapes captured in a CUDA graph and replayed: the regime of a time-lapse pipeline,
    which calls the extract step once per time point with a few hundred objects — a dozen launches whose latencies,
    not their work, set the time of the call (0.11 ms per C3 time point eagerly).

    Static device buffers hold the inputs (labels ``(P, H, W)`` uint16; pixels in the caller's layout) and the table;
    every plane owns ``cap`` table rows (``plane_base[p] = p * cap``), so that nothing the graph bakes in depends on the
    label counts of a time point.  :meth:`run` copies the new inputs in, replays the graph and returns the table (still
    on the device, rows of absent labels NaN); a label above ``cap`` sets the status word — the caller then builds a
    larger instance.  One instance belongs to one stream of one device."""

    def __init__(self, plan: Plan, n_planes: int, H: int, W: int, plane_tile, pixel_shape, pixel_dtype, tile_offset,
                 chan_stride: int, z_stride: int, row_stride: int, n_channels: int, n_z: int, cap: int, device):
        import torch

        self.plan, self.cap, self.P = plan, int(cap), int(n_planes)
        self.device = device
        self.labels = torch.zeros((n_planes, H, W), dtype=torch.uint16, device=device)
        self.pixels = torch.zeros(pixel_shape, dtype=pixel_dtype, device=device)
        self.n_labels = np.full(n_planes, self.cap, dtype=np.int64)
        self.buf, self.table, self.status = alloc_table(n_planes * self.cap, plan.n_columns, device)
        self._args = (np.asarray(plane_tile, dtype=np.int32), self.n_labels, self.pixels,
                      np.ascontiguousarray(tile_offset, dtype=np.int64), int(chan_stride), int(z_stride), int(row_stride),
                      int(n_channels), int(n_z))
        self.stream = torch.cuda.Stream(device=device)
        self.stream.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(self.stream):
            for _ in range(2):  # first calls set kernel attributes, encode tensor maps, upload the plan: not capturable
                self._launch()
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread-local capture mode: other host threads keep making (synchronising) CUDA calls while this one captures
        with _capture_lock:
            with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
                self._launch()
        torch.cuda.current_stream(device).wait_stream(self.stream)

    def _launch(self):
        pt, nl, px, off, cs, zs, rs, C_, Z_ = self._args
        run_planes(self.plan, self.labels, pt, nl, px, off, cs, zs, rs, C_, Z_, out=self.table, status=self.status)

    def run(self, labels=None, pixels=None):
        """Copy new inputs into the static buffers (tensors or arrays of the captured shapes; ``None`` = unchanged), replay."""
        import torch

        if labels is not None:
            if isinstance(labels, torch.Tensor) and labels.is_cuda:
                self.labels.copy_(labels, non_blocking=True)
            else:
                host_to_device(self.labels, labels, slot="labels")
        if pixels is not None:
            if isinstance(pixels, torch.Tensor) and pixels.is_cuda:
                self.pixels.copy_(pixels, non_blocking=True)
            else:
                host_to_device(self.pixels, pixels)
        self.graph.replay()
        return self.table

    def rows(self, n_labels) -> np.ndarray:
        """Row indices of the objects ``1..n_labels[p]`` of every plane inside the capacity-padded table."""
        n_labels = np.asarray(n_labels, dtype=np.int64)
        if n_labels.max(initial=0) > self.cap:
            raise IndexError(f"{int(n_labels.max())} labels in a plane exceed the captured capacity {self.cap}")
        if not len(n_labels):
            return np.zeros(0, np.int64)
        starts = np.cumsum(n_labels) - n_labels  # first output index of every plane
        return np.arange(int(n_labels.sum())) + np.repeat(np.arange(len(n_labels)) * self.cap - starts, n_labels)
