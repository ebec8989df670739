"""Generate ``tests/golden/*.npz`` by running the REAL reference (build container only).

    python oracle/make_golden.py

Imports ``/root/reference/src`` (read-only) behind two throw-away import shims for
packages that are absent here and irrelevant to the pinned functions
(``skimage.segmentation.relabel_sequential`` and ``cp_measure.bulk`` — see
SURVEY.md §8c).  ``/root/reference`` does not exist on the GPU box, so nothing at
test/bench time runs this script; only its committed outputs are read.
TEST INFRASTRUCTURE.
"""

from __future__ import annotations

import ast
import json
import os
import sys
import tempfile
import textwrap
import warnings

import numpy as np

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))


def install_shims():
    d = tempfile.mkdtemp(prefix="aliby_shim_")
    os.makedirs(f"{d}/skimage")
    os.makedirs(f"{d}/cp_measure")
    open(f"{d}/skimage/__init__.py", "w").close()
    open(f"{d}/cp_measure/__init__.py", "w").close()
    with open(f"{d}/skimage/segmentation.py", "w") as f:
        f.write(
            textwrap.dedent(
                """
                import numpy as np
                class _Map:
                    def __init__(self, i, o):
                        self.in_values = np.asarray(i); self.out_values = np.asarray(o)
                    def __getitem__(self, k):
                        return self.out_values[np.searchsorted(self.in_values, k)]
                def relabel_sequential(field, offset=1):
                    vals = np.unique(field); vals = vals[vals > 0]
                    new = np.arange(offset, offset + len(vals))
                    out = np.zeros_like(field)
                    for v, n in zip(vals, new):
                        out[field == v] = n
                    z = np.zeros(1, dtype=vals.dtype)
                    return (out, _Map(np.concatenate([z, vals]), np.concatenate([z, new])),
                            _Map(np.concatenate([z, new]), np.concatenate([z, vals])))
                """
            )
        )
    with open(f"{d}/cp_measure/bulk.py", "w") as f:
        f.write("def get_core_measurements():\n    return {}\ndef get_correlation_measurements():\n    return {}\n")
    sys.path.insert(0, d)
    sys.path.insert(0, REF)


def to_float_rows(results):
    """Scalars -> float64; tuples are spread over two slots (second array)."""
    a = np.full(len(results), np.nan)
    b = np.full(len(results), np.nan)
    is_int = np.zeros(len(results), dtype=bool)
    for i, r in enumerate(results):
        if isinstance(r, tuple):
            a[i], b[i] = float(r[0]), float(r[1])
        else:
            a[i] = float(r)
            is_int[i] = isinstance(r, (int, np.integer))
    return a, b, is_int


def numpy_disk(r):
    yy, xx = np.mgrid[-r : r + 1, -r : r + 1]
    return (yy * yy + xx * xx <= r * r).astype(np.uint8)


def numpy_ellipse(x, y, rot_deg):
    """Filled ellipse in a (4x, 4y) frame, semi-axes (x, y) along (rows, cols), rotated."""
    img = np.zeros((4 * x, 4 * y), dtype=np.uint8)
    rr, cc = np.mgrid[0 : 4 * x, 0 : 4 * y]
    dr, dc = rr - 2 * x, cc - 2 * y
    th = np.deg2rad(rot_deg)
    u = dr * np.cos(th) + dc * np.sin(th)
    v = -dr * np.sin(th) + dc * np.cos(th)
    img[(u / x) ** 2 + (v / y) ** 2 <= 1.0] = 1
    return img


def main():
    warnings.simplefilter("ignore")
    install_shims()
    from aliby_b200 import synth
    from extraction import extract as ref_extract
    from extraction.core.functions import cell as ref_cell
    from extraction.core.functions import trap as ref_trap

    os.makedirs(OUT, exist_ok=True)
    meta = {"numpy": np.__version__}
    import scipy

    meta["scipy"] = scipy.__version__

    shape_metrics = [
        "area",
        "centroid",
        "centroid_x",
        "centroid_y",
        "conical_volume",
        "eccentricity",
        "min_maj_approximation",
        "spherical_volume",
        "volume",
    ]
    intensity_metrics = [
        "mean",
        "std",
        "median",
        "total",
        "total_squared",
        "max2p5pc",
        "max5px_median",
        "moment_of_inertia",
        "ratio",
    ]

    # ---- 1. small single field, Z = 3, reductions max + add -------------------
    pixels, labels = synth.make_field(seed=7, shape=(96, 128), n_channels=2, n_objects=14, n_z=3, semi_axes=(4, 12))
    tree = {"None": {"None": shape_metrics}, 0: {"max": intensity_metrics}, 1: {"max": ["mean", "median"], "add": intensity_metrics}}
    items, res = ref_extract.process_tree_masks(tree, labels, pixels, ref_extract.extract_tree)
    a, b, is_int = to_float_rows(res)
    np.savez_compressed(
        f"{OUT}/field_small.npz",
        pixels=pixels,
        labels=labels,
        tree=json.dumps({str(k): v for k, v in tree.items()}),
        values=a,
        values2=b,
        is_int=is_int,
        n_items=len(items),
    )
    print("field_small", len(items), "items", int(labels.max()), "labels")

    # ---- 2. multi-tile list of masks (one empty tile, one without objects) ----
    rng = np.random.default_rng(11)
    tiles_px, tiles_lab = [], []
    for t in range(4):
        p, l = synth.make_field(seed=20 + t, shape=(64, 64), n_channels=3, n_objects=5, n_z=1, semi_axes=(3, 9))
        if t == 2:
            l = np.zeros_like(l)
        tiles_px.append(p[0])
        tiles_lab.append(l)
    px = np.stack(tiles_px)
    tree2 = {"None": {"None": ["area", "centroid_x", "centroid_y", "eccentricity", "volume"]},
             1: {"max": ["mean", "median", "std", "max5px_median", "max2p5pc"]},
             2: {"add": ["total", "median"]}}
    items, res = ref_extract.process_tree_masks(tree2, tiles_lab, px, ref_extract.extract_tree)
    a, b, is_int = to_float_rows(res)
    tab_res = [float(r) for r in res]
    table = ref_extract.format_extraction((items, tab_res))
    np.savez_compressed(
        f"{OUT}/tiles_list.npz",
        pixels=px,
        labels=np.stack(tiles_lab),
        tree=json.dumps({str(k): v for k, v in tree2.items()}),
        values=a,
        is_int=is_int,
        item_tile=np.array([it[0][0] for it in items]),
        item_label=np.array([it[0][1] for it in items]),
        table_columns=json.dumps(table.column_names),
        table_types=json.dumps([str(t) for t in table.schema.types]),
        table_tile=np.asarray(table.column("tile").to_pylist()),
        table_label=np.asarray(table.column("label").to_pylist()),
        table_values=np.stack(
            [np.asarray(table.column(c).to_pylist(), dtype=float) for c in table.column_names[2:]], axis=1
        ),
    )
    print("tiles_list", len(items), "items; table", table.num_rows, "x", table.num_columns)

    # ---- 3. analytic shapes of tests/extraction/test_volume.py:32-74 (numpy-drawn) ----
    radii = list(range(10, 100, 10))
    recs = []
    for r in radii:
        m = numpy_disk(r)
        mn, mj = ref_cell.min_maj_approximation(m)
        recs.append(("disk", r, 0.0, 0, mn, mj, ref_cell.volume(m), ref_cell.eccentricity(m), ref_cell.conical_volume(m)))
    for x in (10, 20, 30, 50):
        for ecc in (0.0, 0.3, 0.6, 0.8):
            for rot in (10, 40, 90):
                y = int(np.round(np.sqrt(x**2 / (1 - ecc**2))))
                m = numpy_ellipse(x, y, rot)
                mn, mj = ref_cell.min_maj_approximation(m)
                recs.append(("ellipse", x, ecc, rot, mn, mj, ref_cell.volume(m), ref_cell.eccentricity(m), ref_cell.conical_volume(m)))
    np.savez_compressed(
        f"{OUT}/volume_shapes.npz",
        kind=np.array([r[0] for r in recs]),
        x=np.array([r[1] for r in recs]),
        ecc=np.array([r[2] for r in recs]),
        rot=np.array([r[3] for r in recs]),
        out=np.array([r[4:] for r in recs], dtype=float),
    )
    print("volume_shapes", len(recs))

    # ---- 4. degenerate shapes: 1..6 px objects anywhere in a plane, border objects ----
    plane = np.zeros((40, 50), dtype=np.uint16)
    plane[0, 0] = 1
    plane[5, 7:9] = 2
    plane[10:12, 20:23] = 3
    plane[39, 45:50] = 4
    plane[20:23, 30] = 5
    plane[30:36, 10:14] = 6
    plane[0:2, 47:50] = 7
    plane[15, 15:21] = 9  # id 8 absent
    out = []
    for lab in range(1, 10):
        m = plane == lab
        mn, mj = ref_cell.min_maj_approximation(m)
        out.append((mn, mj, ref_cell.volume(m), ref_cell.eccentricity(m), ref_cell.conical_volume(m), ref_cell.area(m)))
    np.savez_compressed(f"{OUT}/degenerate_shapes.npz", labels=plane, out=np.array(out, dtype=float))
    print("degenerate_shapes", len(out))

    # ---- 5. trap/background functions on (Y, X, N) masks (trap.py:6-43) ----
    p, l = synth.make_field(seed=31, shape=(48, 48), n_channels=1, n_objects=4, n_z=1, semi_axes=(3, 8))
    img = p[0, 0, 0]
    stack = np.stack([l == k for k in range(1, int(l.max()) + 1)], axis=2)
    np.savez_compressed(
        f"{OUT}/background.npz",
        image=img,
        labels=l,
        imBackground=float(ref_trap.imBackground(stack, img)),
        background_max5=float(ref_trap.background_max5(stack, img)),
    )

    # ---- 6. tile crop: exec the two functions out of tiler.py (module needs dask/skimage) ----
    src = open("/root/reference/src/aliby/tile/tiler.py").read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "if_out_of_bounds_pad")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "tiler.py", "exec"), ns)
    from aliby.tile.tiles import TileLocations

    frame = np.random.default_rng(5).integers(0, 4000, size=(2, 40, 60)).astype(np.uint16)
    locs = TileLocations([(20, 30), (8, 10), (36, 55), (2, 2)], tile_size=16, max_size=(40, 60), drifts=[[0.0, 0.0], [1.6, -2.4]])
    crops = {}
    for tp in (0, 1):
        for i, tile in enumerate(locs):
            crops[f"tp{tp}_tile{i}"] = ns["if_out_of_bounds_pad"](frame, tile.as_range(tp))
    np.savez_compressed(
        f"{OUT}/tile_crop.npz",
        frame=frame,
        centres=np.array([(20, 30), (8, 10), (36, 55), (2, 2)]),
        drifts=np.array([[0.0, 0.0], [1.6, -2.4]]),
        **crops,
    )
    print("tile_crop", len(crops))

    # ---- 7. overlap (BABY) path on sequential labels, one stack per tile ----
    ov_masks = [tiles_lab[0][None], tiles_lab[1][None]]
    # make ids sequential so that the live reference path reads the right planes
    from skimage.segmentation import relabel_sequential

    ov_masks = [relabel_sequential(m[0])[0][None] for m in ov_masks]
    tree3 = {"None": {"None": ["centroid_x"]}, 0: {"max": ["mean", "median"]}}
    from functools import partial

    items, res = ref_extract.process_tree_masks_overlap(
        tree3, ov_masks, px[:2], partial(ref_extract.extract_tree, overlap=True)
    )
    np.savez_compressed(
        f"{OUT}/overlap.npz",
        pixels=px[:2],
        masks=np.stack(ov_masks),
        tree=json.dumps({str(k): v for k, v in tree3.items()}),
        values=np.array([float(r) for r in res]),
        item_ids=np.array([list(map(int, it[0])) for it in items]),
    )
    print("overlap", len(items))

    # ---- 8. profile assembly: get_profiles_from_state exec'd out of pipe_core.py (the module needs numcodecs/dask/nahual) ----
    # state of a 3-time-point run with two extract steps of one prefix (concatenated) and one extractmulti step
    # (joined on the metadata keys), one time point without objects; results are what the reference's own
    # process_tree_masks returns, cast to float like a pipeline does before format_extraction accepts them.
    import pyarrow
    import pyarrow.parquet

    src = open("/root/reference/src/aliby/pipe_core.py").read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_profiles_from_state")
    ns = {"numpy": np, "np": np, "pyarrow": pyarrow, "pa": pyarrow, "format_extraction": ref_extract.format_extraction}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "pipe_core.py", "exec"), ns)
    # (steps of one prefix must share their columns: pyarrow.concat_tables, pipe_core.py:501, rejects differing schemas —
    # the stock builder gives every object the same tree, pipe_builder.py:115-134)
    tree_a = {"None": {"None": ["area", "centroid_x"]}, 0: {"max": ["mean", "median"]}, 1: {"max": ["max5px_median", "total"]}}
    state = {"data": {"extract_nuclei": [], "extract_cell": [], "extractmulti_nuclei": []}}
    inputs = {}
    for tp in range(3):
        p, l = synth.make_field(seed=60 + tp, shape=(48, 64), n_channels=2, n_objects=4, n_z=1, semi_axes=(3, 7))
        if tp == 1:
            l = np.zeros_like(l)  # a time point without objects: its tables are skipped (pipe_core.py:486)
        _, l2 = synth.make_field(seed=80 + tp, shape=(48, 64), n_channels=2, n_objects=3, n_z=1, semi_axes=(4, 9))
        inputs[f"pixels{tp}"], inputs[f"labels{tp}"], inputs[f"labels_cell{tp}"] = p, l, l2
        for step, lab_img in (("extract_nuclei", l), ("extract_cell", l2)):
            items, res = ref_extract.process_tree_masks(tree_a, lab_img, p, ref_extract.extract_tree)
            state["data"][step].append((items, [float(r) for r in res]))
        # colocalisation-shaped results: dict-valued metrics keyed by ((ch0, ch1), red_ch, red_z, metric) instructions
        n_obj = int(l.max())
        multi_items = tuple(((0, lab), ((0, 1), "None", "max", "pearson")) for lab in range(1, n_obj + 1))
        multi_res = [{"Correlation_Pearson": np.array([0.25 * lab + tp])} for lab in range(1, n_obj + 1)]
        state["data"]["extractmulti_nuclei"].append((multi_items, multi_res))
    pipeline = {"steps": {"tile": {}, "segment_nuclei": {}, "extract_nuclei": {}, "extract_cell": {}, "extractmulti_nuclei": {}}}
    profiles = ns["get_profiles_from_state"](state, pipeline)
    pq_path = os.path.join(tempfile.mkdtemp(prefix="aliby_pq_"), "profiles.parquet")
    pyarrow.parquet.write_table(profiles, pq_path, compression="zstd")  # pipe_core.py:411-413
    back = pyarrow.parquet.read_table(pq_path)
    cols = profiles.column_names
    np.savez_compressed(
        f"{OUT}/profiles.npz",
        tree_a=json.dumps({str(k): v for k, v in tree_a.items()}),
        columns=json.dumps(cols),
        types=json.dumps([str(t) for t in profiles.schema.types]),
        parquet_types=json.dumps([str(t) for t in back.schema.types]),
        n_rows=profiles.num_rows,
        **{f"col{j}": (np.array([x or "" for x in profiles.column(c).to_pylist()], dtype=str)
                       if str(profiles.schema.types[j]) == "string" else np.asarray(profiles.column(c).to_pylist(), dtype=float))
           for j, c in enumerate(cols)},
        **{f"null{j}": np.asarray(profiles.column(c).is_null().to_pylist()) for j, c in enumerate(cols)},
        **{f"state_{step}_{tp}_values": np.array([r if isinstance(r, float) else float(r["Correlation_Pearson"][0]) for r in out[1]])
           for step, outs in state["data"].items() for tp, out in enumerate(outs)},
        **inputs,
    )
    print("profiles", profiles.num_rows, "rows x", profiles.num_columns, "columns")

    with open(f"{OUT}/META.json", "w") as f:
        json.dump(meta, f)


if __name__ == "__main__":
    main()
