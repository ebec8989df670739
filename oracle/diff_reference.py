"""Randomised differential check of the oracle against the REAL reference (build container only) — TEST INFRASTRUCTURE.

    python oracle/diff_reference.py [n_cases]

``tests/golden/*.npz`` pin the oracle on fixed cases; this script throws random small planes at both — odd sizes, 1-pixel
cells, id gaps, touching cells, cells on the border, constant / zero / saturated cells, uint8 and uint16, Z stacks under
max and add, several tiles — runs ``extraction.extract.process_tree_masks`` of ``/root/reference/src`` and
``oracle.port.run_tree`` / ``oracle.fast.run_tree`` on the same inputs and compares every value (bit-exact for what
NumPy derives from integers, 1e-12 for fp64 sums).  ``/root/reference`` does not exist on the GPU box: nothing in
``tests/`` runs this; its last result is recorded in DESIGN.md §5."""

from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import make_golden as mg  # noqa: E402

SHAPE = ["area", "centroid_x", "centroid_y", "eccentricity", "volume", "conical_volume", "spherical_volume", "min_maj_approximation"]
INTENSITY = ["mean", "std", "median", "total", "total_squared", "max2p5pc", "max5px_median", "moment_of_inertia", "ratio"]


def random_case(rng):
    H, W = int(rng.integers(9, 60)), int(rng.integers(9, 70))
    n_tiles = int(rng.integers(1, 4))
    Z = int(rng.integers(1, 4))
    dtype = np.uint8 if rng.random() < 0.3 else np.uint16
    top = np.iinfo(dtype).max
    masks = []
    for _ in range(n_tiles):
        lab = np.zeros((H, W), np.uint16)
        nid = 1
        for _ in range(int(rng.integers(0, 7))):
            h, w = int(rng.integers(1, max(2, H // 2))), int(rng.integers(1, max(2, W // 2)))
            r, c = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
            yy, xx = np.mgrid[0:h, 0:w]
            blob = ((yy - (h - 1) / 2) / max(1, h / 2)) ** 2 + ((xx - (w - 1) / 2) / max(1, w / 2)) ** 2 <= 1.0 if rng.random() < 0.6 else np.ones((h, w), bool)
            lab[r : r + h, c : c + w][blob] = nid
            nid += int(rng.integers(1, 3))
        masks.append(lab)
    kind = rng.integers(0, 3)
    if kind == 0:
        px = rng.integers(0, top + 1, size=(n_tiles, 2, Z, H, W)).astype(dtype)
    elif kind == 1:
        px = np.clip(rng.integers(0, top // 2 + 1) + rng.poisson(20, size=(n_tiles, 2, Z, H, W)), 0, top).astype(dtype)
    else:
        px = np.full((n_tiles, 2, Z, H, W), rng.choice([0, 7, top]), dtype=dtype)
    tree = {"None": {"None": list(SHAPE)}, 0: {"max": list(INTENSITY)}, 1: {"add": ["mean", "median", "total", "std", "max2p5pc"]}}
    return tree, masks, px


def same(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    if a.shape != b.shape:
        return False
    nan = np.isnan(a)
    return bool(np.array_equal(nan, np.isnan(b)) and np.all(np.abs(a[~nan] - b[~nan]) <= tol * np.maximum(1.0, np.abs(b[~nan]))))


def main(n_cases: int = 40) -> int:
    mg.install_shims()
    sys.path.insert(0, mg.REF)
    warnings.filterwarnings("ignore")
    from extraction import extract as ref

    from oracle import fast, port

    rng = np.random.default_rng(20260)
    exact = {"area", "centroid_x", "centroid_y", "eccentricity", "volume", "min_maj_approximation", "mean", "median", "total",
             "total_squared", "max2p5pc", "max5px_median", "ratio"}
    n_values = 0
    for case in range(n_cases):
        tree, masks, px = random_case(rng)
        m = masks if len(masks) > 1 else masks[0]
        r_items, r_res = ref.process_tree_masks(tree, m, px, ref.extract_tree)
        for name, impl in (("port", port), ("fast", fast)):
            o_items, o_res = impl.run_tree(tree, m, px)
            assert len(o_items) == len(r_items), (case, name, len(o_items), len(r_items))
            for (ri, rr), (oi, orr) in zip(zip(r_items, r_res), zip(o_items, o_res)):
                assert tuple(ri[0]) == tuple(oi[0]) and tuple(ri[1]) == tuple(oi[1]), (case, name, ri, oi)
                tol = 0.0 if ri[1][2] in exact else 1e-12
                assert same(rr, orr, tol), (case, name, ri, rr, orr)
                n_values += 1
    print(f"{n_cases} random cases, {n_values} values: oracle.port and oracle.fast agree with the reference")
    # ---- tile windows and the out-of-frame padding (tiles.py:109-166, tiler.py:601-650) ----
    import ast

    src = open(os.path.join(mg.REF, "aliby/tile/tiler.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "if_out_of_bounds_pad")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "tiler.py", "exec"), ns)
    from aliby.tile.tiles import TileLocations

    n_crops = 0
    for case in range(n_cases):
        H, W = int(rng.integers(20, 90)), int(rng.integers(20, 90))
        size = int(rng.choice([8, 11, 16, 17]))
        frame = rng.integers(0, 4000, size=(int(rng.integers(1, 4)), H, W)).astype(np.uint16 if rng.random() < 0.7 else np.uint8)
        centres = [(int(rng.integers(-2, H + 3)), int(rng.integers(-2, W + 3))) for _ in range(int(rng.integers(1, 6)))]
        drifts = [[0.0, 0.0]] + [[float(rng.normal(0, 2)), float(rng.normal(0, 2))] for _ in range(int(rng.integers(0, 3)))]
        locs = TileLocations(centres, tile_size=size, max_size=(H, W), drifts=drifts)
        for tp in range(len(drifts)):
            for i, tile in enumerate(locs):
                want = ns["if_out_of_bounds_pad"](frame, tile.as_range(tp))
                got = port.crop_with_padding(frame, port.tile_window(centres[i], (size, size), drifts, tp))
                assert want.shape == got.shape and want.dtype == got.dtype, (case, tp, i, want.shape, got.shape, want.dtype, got.dtype)
                assert np.array_equal(want, got, equal_nan=True), (case, tp, i)
                n_crops += 1
    print(f"{n_crops} tile crops (drifts, windows that leave the frame, NaN tiles): oracle.port agrees with the reference")
    # ---- the long -> wide pivot of the HOST side (aliby_b200.extract.format_extraction, generic path) ----
    from aliby_b200 import extract as ours

    n_tables = 0
    for case in range(n_cases):
        objs = [(int(t), int(k)) for t in range(int(rng.integers(1, 4))) for k in range(1, int(rng.integers(1, 6)))]
        insts = [("None", "None", "area"), (0, "max", "mean"), (1, "add", "median"), (0, "max", "intensity"), ((0, 1), "None", "max", "pearson")]
        insts = [insts[j] for j in rng.permutation(len(insts))[: int(rng.integers(1, len(insts) + 1))]]
        items, results = [], []
        for o in objs:
            for inst in insts:
                if rng.random() < 0.1:
                    continue  # a ragged product: missing cells become nulls
                items.append((o, inst))
                if inst[-1] in ("intensity", "pearson"):
                    results.append({f"K_{j}": np.array([float(rng.normal())]) for j in range(int(rng.integers(1, 4)))})
                else:
                    results.append(float(rng.normal()) if rng.random() < 0.9 else float("nan"))
        if not items:
            continue
        a = ref.format_extraction((tuple(items), list(results)))
        b = ours.format_extraction((tuple(items), list(results)))
        assert a.column_names == b.column_names and a.schema == b.schema, (case, a.schema, b.schema)
        for c in a.column_names:
            x, y = a.column(c).to_pylist(), b.column(c).to_pylist()
            assert len(x) == len(y) and all((p == q) or (p is not None and q is not None and p != p and q != q) for p, q in zip(x, y)), (case, c)
        n_tables += 1
    print(f"{n_tables} random (instructions, results) lists: aliby_b200.extract.format_extraction builds the reference's table")
    # ---- profile assembly over time points and steps (pipe_core.py:453-512) ----
    import pyarrow

    from aliby_b200 import pipe as our_pipe

    src = open(os.path.join(mg.REF, "aliby/pipe_core.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_profiles_from_state")
    ns = {"numpy": np, "np": np, "pyarrow": pyarrow, "pa": pyarrow, "format_extraction": ref.format_extraction}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "pipe_core.py", "exec"), ns)
    n_states = 0
    for case in range(n_cases):
        steps = ["extract_nuclei"] + (["extract_cell"] if rng.random() < 0.6 else []) + (["extractmulti_nuclei"] if rng.random() < 0.6 else [])
        insts = {"extract": [("None", "None", "area"), (0, "max", "mean"), (1, "max", "median")],
                 "extractmulti": [((0, 1), "None", "max", "pearson")]}
        n_tp = int(rng.integers(1, 5))
        # (the join of extract with extractmulti is keyed by tp / tile / object / label: the same objects on both sides)
        objs_of = {(obj, tp): ([] if rng.random() < 0.2 else [(int(t), int(k)) for t in range(2) for k in range(1, int(rng.integers(2, 5)))])
                   for obj in ("nuclei", "cell") for tp in range(n_tp)}
        state = {"data": {s_: [] for s_ in steps}}
        for s_ in steps:
            prefix, obj = s_.split("_")
            for tp in range(n_tp):
                items, results = [], []
                for o in objs_of[(obj, tp)]:
                    for inst in insts[prefix]:
                        items.append((o, inst))
                        results.append({"Correlation_Pearson": np.array([float(rng.normal())])} if prefix == "extractmulti" else float(rng.normal()))
                state["data"][s_].append((tuple(items), results))
        pipeline = {"steps": {"tile": {}, **{s_: {} for s_ in steps}}}
        a = ns["get_profiles_from_state"](state, pipeline)
        b = our_pipe.get_profiles_from_state(state, pipeline)
        assert a.column_names == b.column_names and a.schema == b.schema, (case, a.schema, b.schema)
        key_cols = [c for c in a.column_names if c.startswith("metadata_")]

        def keyed(t):
            rows = t.to_pylist()
            return {tuple(r[k] for k in key_cols): tuple((c, None if r[c] is None else round(r[c], 12)) for c in t.column_names if c not in key_cols)
                    for r in rows}

        assert a.num_rows == b.num_rows and keyed(a) == keyed(b), case
        n_states += 1
    print(f"{n_states} random pipeline states: aliby_b200.pipe.get_profiles_from_state builds the reference's profile table")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 40))
