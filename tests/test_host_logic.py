"""CPU-only checks: C-ABI exports, plan compiler, host-side table logic, sharding (gloo, world 2)."""

import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_same, golden_tree, load_golden


def test_library_exports_every_declared_symbol():
    """The shared library loads and exports exactly what include/aliby_b200.h declares."""
    from aliby_b200 import _native as nat
    from aliby_b200 import build

    build.build()
    header = open(os.path.join(ROOT, "include", "aliby_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(abx_\w+)\s*\(", header, flags=re.M))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    lib = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert nat.lib().abx_version() == nat.ABI_VERSION == 4
    # struct mirrors: sizes must agree with the C side (48-byte records, 16-byte requests)
    assert ctypes.sizeof(nat.ObjectRec) == 48 and ctypes.sizeof(nat.Request) == 16 and ctypes.sizeof(nat.Column) == 8 and ctypes.sizeof(nat.Pair) == 24


def test_graft_entry_build_runs_on_cpu():
    """The driver's "does it build" hook: compiles (or finds) the library, checks its ABI version, imports the oracles."""
    sys.path.insert(0, ROOT)
    import __graft_entry__

    __graft_entry__.build()


def test_abi_rejects_bad_arguments_without_a_gpu():
    from aliby_b200 import _native as nat

    lib = nat.lib()
    a = nat.ExtractArgs()
    a.label_dtype = nat.U32  # no kernel
    need = ctypes.c_size_t(0)
    assert lib.abx_extract_workspace_bytes(ctypes.byref(a), ctypes.byref(need)) == -2
    assert b"label dtype" in lib.abx_last_error()
    a.label_dtype = nat.U16
    a.n_planes, a.H, a.W, a.n_objects = 4, 2160, 2160, 8000
    a.n_requests, a.pixel_dtype, a.C, a.Z, a.row_stride = 5, nat.U32, 5, 1, 2160
    assert lib.abx_extract_workspace_bytes(ctypes.byref(a), ctypes.byref(need)) == -2
    a.pixel_dtype = nat.U16
    assert lib.abx_extract_workspace_bytes(ctypes.byref(a), ctypes.byref(need)) == 0
    assert need.value > 8004 * 48
    # Z stacks with 16-byte rows are reduced into planes of the workspace: one H x W x 4-byte slot per (tile, request)
    flat = need.value
    a.n_tiles, a.Z = 4, 16
    assert lib.abx_extract_workspace_bytes(ctypes.byref(a), ctypes.byref(need)) == 0
    assert 4 * 5 * 2160 * 2160 * 4 <= need.value - flat < 4 * 5 * 2160 * 2160 * 4 + 8192
    a.W, a.row_stride = 2161, 2161  # rows that are not 16-byte multiples keep the fused gathers: no planes
    assert lib.abx_extract_workspace_bytes(ctypes.byref(a), ctypes.byref(need)) == 0
    assert need.value < flat + (1 << 20)
    with pytest.raises(NotImplementedError):
        nat.check(-2, "x")


def test_plan_compiler_matches_reference_flattening():
    from aliby_b200 import engine
    from oracle import port

    tree = {"None": {"None": ["area", "centroid", "volume"]}, 0: {"max": ["mean", "median"], "add": ["total"]},
            2: {"max": ["max5px_median", "imBackground"]}}
    plan = engine.compile_tree(tree)
    assert plan.instructions == port.tree_instructions(tree)
    assert plan.error is None and plan.need_edt == 1 and plan.with_background
    assert [r[:2] for r in plan.requests] == [[0, 0], [0, 1], [2, 0]]
    assert len(plan.inst_cols[1]) == 2  # centroid -> (x, y)
    from aliby_b200 import _native as nat

    assert plan.requests[2][2] == nat.F_TOP5 | nat.F_MEDIAN and plan.requests[2][3] == nat.F_MEDIAN
    for bad, exc in [({0: {"max": ["nope"]}}, KeyError), ({0: {"nope": ["mean"]}}, KeyError),
                     ({0: {"median": ["mean"]}}, Exception), ({"None": {"None": ["mean"]}}, TypeError),
                     ]:
        assert isinstance(engine.compile_tree(bad).error, exc), bad
    div = engine.compile_tree({0: {"div": ["mean"]}})  # np.divide.reduce: served by the float kernel
    assert div.error is None and div.requests[0][1] == nat.RED_DIV


def test_registry_names_match_reference():
    """The 18 cell + 2 trap names of the reference registry."""
    from aliby_b200.functions.loaders import load_funs, load_redfuns
    from oracle import port

    cell, trap, both = load_funs()
    assert set(port.CELL_METRICS) <= set(cell) and set(trap) == {"imBackground", "background_max5"}
    assert set(both) == set(cell) | set(trap)
    red = load_redfuns()
    assert list(red) == ["max", "mean", "median", "div", "add", "None"] and red["max"] is np.maximum


def test_format_extraction_generic_paths():
    """Table contract of tests/test_nahual_embed_minimal.py:35-101 + the golden pivot."""
    import pyarrow as pa

    from aliby_b200.extract import format_extraction
    from aliby_b200.pipe import get_profiles_from_state

    emb = np.arange(12, dtype=np.float32).reshape(3, 4)
    table = format_extraction(((("__", "__"),), (emb,)))
    assert isinstance(table, pa.Table) and table.num_rows == 3
    assert len([c for c in table.column_names if c.startswith("X_")]) == 4
    from itertools import cycle

    with pytest.raises(ValueError, match="zip"):
        format_extraction((cycle((("__", "__"),)), (emb,)))
    with pytest.raises(Exception, match="invalid value"):
        format_extraction(((((0, 1), (0, "max", "centroid")),), ((1.0, 2.0),)))
    state = {"data": {"nahual_embed_cells": [emb[:2], emb[:2] + 100]}}
    prof = get_profiles_from_state(state, {"steps": {"nahual_embed_cells": {}}})
    assert set(prof.column("metadata_tp").to_pylist()) == {0, 1}
    assert set(prof.column("metadata_object").to_pylist()) == {"cells"}
    assert str(prof.schema.field("metadata_tp").type) == "uint16"
    # golden long -> wide pivot through the generic path (python lists, no GPU involved)
    g = load_golden("tiles_list.npz")
    from oracle import fast

    items, res = fast.run_tree(golden_tree(g), [m for m in g["labels"]], g["pixels"])
    table = format_extraction((items, [float(r) for r in res]))
    import json

    assert table.column_names == json.loads(str(g["table_columns"]))
    got = np.stack([np.asarray(table.column(c).to_pylist(), dtype=float) for c in table.column_names[2:]], axis=1)
    assert_same(got, g["table_values"], 1e-12, "pivot")
    # profiles of an extract step: metadata columns appended, tile/label renamed
    prof = get_profiles_from_state({"data": {"extract_nuclei": [(items, [float(r) for r in res])]}},
                                   {"steps": {"extract_nuclei": {}}})
    assert prof.column_names[:2] == ["metadata_tile", "metadata_label"]
    assert prof.column_names[-2:] == ["metadata_object", "metadata_tp"] and prof.num_rows == table.num_rows


def test_tile_origins_follow_reference_windows():
    from aliby_b200.tile import tile_origins
    from oracle import port

    centres = np.array([(20, 30), (8, 10), (36, 55)])
    drifts = np.array([[0.0, 0.0], [1.6, -2.4]])
    for tp in (0, 1):
        got = tile_origins(centres, (16, 12), drifts, tp)
        for c, o in zip(centres, got):
            win = port.tile_window(c, (16, 12), drifts, tp)
            assert (win[0].start, win[1].start) == tuple(o)


def test_shard_units_partition():
    from aliby_b200.sharding import shard_units

    for n in (0, 1, 7, 3456):
        for world in (1, 2, 4, 8):
            for mode in ("contiguous", "round_robin"):
                parts = [shard_units(n, r, world, mode) for r in range(world)]
                assert sorted(np.concatenate(parts).tolist()) == list(range(n))
                assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from aliby_b200 import synth
from aliby_b200.sharding import extract_sharded
from oracle import fast

dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
tree = {"None": {"None": ["area", "centroid_x"]}, 0: {"max": ["mean", "median", "max5px_median"]}}

def load(seed):
    px, lab = synth.make_field(seed, (64, 80), 1, 6, semi_axes=(3, 8))
    return lab, px

def compute(tree_, masks, pixels):  # CPU stand-in for the CUDA path: the oracle (tests only)
    items, res = fast.run_tree(tree_, masks, pixels)
    n_inst = 5
    vals = np.array([float(r) for r in res]).reshape(-1, n_inst)
    objs = np.array([it[0] for it in items[::n_inst]])
    return objs, [str(i) for i in range(n_inst)], vals

units = [100, 101, 102, 103, 104]
out = extract_sharded(tree, units, load, compute=compute)
if dist.get_rank() == 0:
    assert [u for u, *_ in out] == units
    for u, objs, names, vals in out:
        o2, _, v2 = compute(tree, *load(u))
        assert np.array_equal(objs, o2) and np.array_equal(np.nan_to_num(vals, nan=-1), np.nan_to_num(v2, nan=-1))
    print("SHARD_OK", len(out))
else:
    assert out is None
dist.destroy_process_group()
'''


def test_sharded_extraction_two_ranks_gloo(tmp_path):
    """World size 2 over gloo on CPU: units are split, tables gathered on rank 0, no data-path collective."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29500 + (os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK 5" in outs[0]


def test_profiles_from_dense_tables_match_item_path(tmp_path):
    """Dense tables -> profile parquet == the reference-shaped path (format_extraction + get_profiles_from_state)."""
    import pyarrow.parquet as pq

    from aliby_b200 import pipe
    from aliby_b200.extract import ExtractionTable

    rng = np.random.default_rng(5)
    insts = [("None", "None", "area"), (1, "max", "median"), (0, "max", "mean")]
    names = ["/".join(str(x) for x in i) + f"/{i[-1]}" for i in insts]
    state = {"data": {"extract_nuclei": []}}
    tables = []
    for tp in range(3):
        objs = np.array([(t, l) for t in range(2) for l in range(1, 3 + tp)], dtype=np.int64)
        vals = rng.normal(size=(len(objs), len(insts)))
        vals[0, 1] = np.nan  # absent label -> NaN row cell
        tables.append(ExtractionTable(objs, names, vals))
        items = tuple(((int(t), int(l)), i) for (t, l) in objs for i in insts)
        state["data"]["extract_nuclei"].append((items, vals.reshape(-1).tolist()))
    want = pipe.get_profiles_from_state(state, {"steps": {"extract_nuclei": {}}})
    got = pipe.profiles_from_tables(tables, "nuclei")
    assert got.column_names == want.column_names and got.schema.field("metadata_tp").type == "uint16"
    for c in want.column_names:
        a, b = got.column(c).to_pylist(), want.column(c).to_pylist()
        assert all((x == y) or (x != x and y != y) for x, y in zip(a, b, strict=True)), c
    path = pipe.write_profiles(got, tmp_path, "pos001")
    back = pq.read_table(path)
    assert back.num_rows == got.num_rows and back.column_names == got.column_names
    assert pq.ParquetFile(path).metadata.row_group(0).column(0).compression == "ZSTD"


def test_init_step_seam_matches_reference_partial_shape():
    """init_step_fn seam (pipe.py:47-72, pipe_core.py:68-81): extract_* -> partial(process, measure_fn, tree, **kwargs)."""
    from functools import partial

    from aliby_b200 import extract, pipe

    tree = {"None": {"None": ["area"]}, 0: {"max": ["mean"]}}
    step = pipe.init_step("extract_nuclei", {"tree": tree, "kwargs": {"ncores": None}})
    assert isinstance(step, partial) and step.func is extract.process_tree_masks
    assert step.keywords["measure_fn"] is extract.extract_tree and step.keywords["tree"] is tree
    assert step.keywords["ncores"] is None
    baby = pipe.init_step("extract_cells", {"tree": tree}, overlap=True)
    assert baby.func is extract.process_tree_masks_overlap and baby.keywords["measure_fn"].keywords == {"overlap": True}
    with pytest.raises(ValueError, match="missing required 'tree'"):
        pipe.init_step("extract_nuclei", {})
    # extractmulti_* (pipe_core.py:84-92): the same partial with extract_tree_multi; every non-extract step belongs to the
    # reference's own init_step (not installed next to us in this container)
    multi_tree = {(0, 1): {"None": {"max": ["pearson", "costes"]}}}
    multi = pipe.init_step("extractmulti_nuclei", {"tree": multi_tree, "kwargs": {"ncores": 2}})
    assert multi.func is extract.process_tree_masks and multi.keywords["measure_fn"] is extract.extract_tree_multi
    assert multi.keywords["tree"] is multi_tree and multi.keywords["ncores"] == 2
    with pytest.raises(ValueError, match="missing required 'tree'"):
        pipe.init_step("extractmulti_nuclei", {})
    with pytest.raises(ImportError):
        pipe.init_step("segment_nuclei", {})


def test_init_step_serves_tile_steps_with_the_fused_tiler():
    """tile* through the seam (pipe.py:56-57, tiler.py:393-448): pre-located tiles give a FusedTiler whose run_tp returns
    {"drift", "pixels": TileView}; a reference Tiler keeps everything but its crop (fuse_reference_tiler)."""
    from aliby_b200 import pipe
    from aliby_b200.tile import FusedTiler, TileView, fuse_reference_tiler, tile_origins

    pixels = np.zeros((3, 2, 1, 200, 240), np.uint16)
    centres = [(60, 70), (120, 150)]
    step = pipe.init_step("tile", {"pixels": pixels, "tile_centres": centres, "tile_size": 64})
    assert isinstance(step, FusedTiler)
    out = step.run_tp(1)
    assert set(out) == {"drift", "pixels"} and isinstance(out["pixels"], TileView)
    assert out["pixels"].shape == (2, 2, 1, 64, 64)
    assert np.array_equal(out["pixels"].origins, tile_origins(centres, 64))
    with pytest.raises(ValueError, match="tile_size"):
        pipe.init_step("tile", {"pixels": pixels, "tile_centres": centres})
    with pytest.raises(ImportError, match="image readers"):  # anything else needs the reference's Tiler (absent here)
        pipe.init_step("tile", {"image_kwargs": {"source": "x.tiff"}, "tile_size": 64})

    class Tile:  # stand-in with the reference's interface (tiles.py:109-166 as_range, tiler.py tile_locs / pixels)
        def __init__(self, c):
            self.c = c

        def as_range(self, tp):
            return slice(self.c[0] - 32, self.c[0] + 32), slice(self.c[1] - 32, self.c[1] + 32)

    class Locs:
        tiles = [Tile(c) for c in centres]
        tile_size = 64

    class RefTiler:
        def __init__(self):
            self.pixels, self.tile_locs, self.tile_size = pixels, Locs(), 64

        def get_fczyx(self, tp):
            raise AssertionError("the reference crop must not run")

        def run_tp(self, tp):
            return {"drift": [0.0, 0.0], "pixels": self.get_fczyx(tp)}

    fused = fuse_reference_tiler(RefTiler())
    view = fused.run_tp(2)["pixels"]
    assert isinstance(view, TileView) and np.array_equal(view.origins, tile_origins(centres, 64))


def test_stock_pipeline_steps_initialise():
    """The step dicts the reference's builder emits (pipe_builder.py:115-134, restated here because aliby.pipe_builder
    imports modules that are absent in this container): extract steps with cp_measure `sizeshape` + `intensity`
    compile to a plan with dict-valued columns; features without a kernel are reported at init time."""
    from aliby_b200 import engine, pipe

    channels = [1, 0]
    kw = {"ncores": None, "cp_measure_kwargs": {"intensity": {"edge_measurements": False}}}
    tree = {"None": {"None": ("sizeshape",)}, **{ch: {"max": ("intensity",)} for ch in channels}}
    step = pipe.init_step("extract_nuclei", {"tree": tree, "kwargs": kw})
    assert step.func.__name__ == "process_tree_masks" and step.keywords["cp_measure_kwargs"] == kw["cp_measure_kwargs"]
    plan = engine.compile_tree(tree, kw["cp_measure_kwargs"])
    assert plan.error is None and plan.need_edt == 7
    assert plan.inst_keys[0][0] == "AreaShape_Area" and "Intensity_MADIntensity" in plan.inst_keys[1]
    # default builder features (radial_zernikes, feret, texture, ...) and edge intensities have no kernel
    assert isinstance(engine.compile_tree({0: {"max": ("texture",)}}).error, KeyError)
    assert isinstance(engine.compile_tree({0: {"max": ("intensity",)}}).error, NotImplementedError)
    step = pipe.init_step("extract_cell", {"tree": {0: {"max": ("zernike",)}}, "kwargs": {"ncores": None}})
    with pytest.raises(KeyError, match="zernike"):  # (reference absent: the step raises at its first call with objects)
        step(masks=np.ones((8, 8), np.uint16), pixels=np.zeros((1, 1, 1, 8, 8), np.uint16))


def test_default_builder_tree_is_split_between_the_gpu_and_the_reference(monkeypatch):
    """The DEFAULT feature list of build_pipeline_steps (pipe_builder.py:49-56: radial_zernikes, intensity, feret, texture,
    radial_distribution, zernike; sizeshape on the masks): with the reference importable the seam keeps the branches that
    have a kernel and hands the others to the reference's own step built from the same parameters."""
    from aliby_b200 import pipe

    default = ("radial_zernikes", "intensity", "feret", "texture", "radial_distribution", "zernike")
    kw = {"ncores": None, "cp_measure_kwargs": {"intensity": {"edge_measurements": False}}}
    tree = {"None": {"None": ("sizeshape",)}, 1: {"max": default}, 0: {"max": default}}
    ours, theirs = pipe._split_tree(tree, kw["cp_measure_kwargs"])
    assert ours == {"None": {"None": ["sizeshape"]}, 1: {"max": ["intensity"]}, 0: {"max": ["intensity"]}}
    rest = ["radial_zernikes", "feret", "texture", "radial_distribution", "zernike"]
    assert theirs == {1: {"max": rest}, 0: {"max": rest}}
    # without the kwargs the edge features of `intensity` are wanted: no kernel, the whole feature goes to the reference
    assert pipe._split_tree(tree)[1][1]["max"] == list(default)
    seen = {}

    def fake_reference_init_step(step_name, parameters, other_steps, why):
        seen["step"], seen["tree"], seen["kwargs"] = step_name, parameters["tree"], parameters["kwargs"]
        return lambda masks, pixels, **kw_: ((("ref-item",),), ["ref-result"])

    monkeypatch.setattr(pipe, "_reference_init_step", fake_reference_init_step)
    step = pipe.init_step("extract_nuclei", {"tree": tree, "kwargs": kw})
    assert seen == {"step": "extract_nuclei", "tree": theirs, "kwargs": kw} and step.__name__ == "split_step"
    # a tree the kernels cover completely never asks for the reference
    monkeypatch.setattr(pipe, "_reference_init_step", lambda *a, **k: (_ for _ in ()).throw(AssertionError("not needed")))
    whole = pipe.init_step("extract_nuclei", {"tree": ours, "kwargs": kw})
    assert whole.func.__name__ == "process_tree_masks" and whole.keywords["tree"] == ours


def test_ctypes_mirrors_match_the_c_header(tmp_path):
    """Field offsets / sizes of the ctypes mirrors == what a C compiler lays out for include/aliby_b200.h."""
    import shutil

    from aliby_b200 import _native as nat

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    fields = [f[0] for f in nat.ExtractArgs._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "aliby_b200.h"', "int main(void) {"]
    prog += [f'  printf("{f} %zu\\n", offsetof(abx_extract_args, {f}));' for f in fields]
    prog += ['  printf("sizeof_args %zu\\n", sizeof(abx_extract_args));',
             '  printf("sizeof_request %zu\\n", sizeof(abx_request));',
             '  printf("sizeof_column %zu\\n", sizeof(abx_column));',
             '  printf("sizeof_rec %zu\\n", sizeof(abx_object_rec));',
             '  printf("pair %zu %zu %zu %d %d\\n", sizeof(abx_pair), offsetof(abx_pair, features), offsetof(abx_pair, threshold_fraction), ABX_M_CO_K_2, ABX_PF_RWC);',
             '  printf("enums %d %d %d %d %d %d\\n", ABX_F64, ABX_RED_DIV, ABX_M_BACKGROUND_MAX5, ABX_M_MEAN, ABX_F_MOI, ABX_ERR_UNSUPPORTED);',
             "  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split(" ", 1) for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for f in fields:
        assert int(out[f]) == getattr(nat.ExtractArgs, f).offset, f
    assert int(out["sizeof_args"]) == ctypes.sizeof(nat.ExtractArgs)
    assert int(out["sizeof_request"]) == ctypes.sizeof(nat.Request) and int(out["sizeof_column"]) == ctypes.sizeof(nat.Column)
    assert int(out["sizeof_rec"]) == ctypes.sizeof(nat.ObjectRec)
    assert out["pair"].split() == [str(v) for v in (ctypes.sizeof(nat.Pair), nat.Pair.features.offset,
                                                    nat.Pair.threshold_fraction.offset, nat.METRIC["co_k_2"], nat.PF_RWC)]
    assert out["enums"].split() == [str(v) for v in (nat.F64, nat.RED_DIV, nat.METRIC["background_max5"], nat.METRIC["mean"],
                                                     nat.F_MOI, -2)]


def test_extractmulti_tree_compiles_into_pairs():
    """The stock builder's colocalisation tree (pipe_builder.py:19-43): every channel pair becomes one abx_pair over two
    shared requests, each feature a dict-valued instruction; `costes` is the only part without a kernel and the seam
    splits the tree along that line (same nesting on both sides)."""
    from itertools import combinations

    from aliby_b200 import _native as nat
    from aliby_b200 import engine, pipe

    tree = {pair: {"None": {"max": ["pearson", "costes", "manders_fold", "rwc"]}} for pair in combinations((0, 2, 3), r=2)}
    plan = engine.compile_tree(tree)
    assert isinstance(plan.error, NotImplementedError) and "costes" in str(plan.error)
    ours, theirs = pipe._split_multi_tree(tree)
    assert theirs == {pair: {"None": {"max": ["costes"]}} for pair in tree}
    assert ours == {pair: {"None": {"max": ["pearson", "manders_fold", "rwc"]}} for pair in tree}
    plan = engine.compile_tree(ours, {"manders_fold": {"thr": 15}})
    assert plan.error is None
    assert [r[:2] for r in plan.requests] == [[0, nat.RED_MAX], [2, nat.RED_MAX], [3, nat.RED_MAX]]
    assert plan.pairs == [[0, 1, 3, 0.15], [0, 2, 3, 0.15], [1, 2, 3, 0.15]]
    assert plan.instructions[0] == ((0, 2), "None", "max", "pearson") and plan.inst_keys[0] == ["Correlation_Pearson"]
    assert plan.inst_keys[2] == ["Correlation_RWC_1", "Correlation_RWC_2"]
    assert [plan.columns[j] for j in plan.inst_cols[2]] == [(0, nat.METRIC["co_rwc_1"]), (0, nat.METRIC["co_rwc_2"])]
    # another threshold is another pair; the reference's error behaviour for unknown names
    plan = engine.compile_tree({(0, 1): {"None": {"max": ["manders_fold"]}}}, {"manders_fold": {"thr": 40}})
    assert plan.pairs == [[0, 1, 1, 0.4]]
    assert isinstance(engine.compile_tree({(0, 1): {"None": {"max": ["nope"]}}}).error, KeyError)
    assert isinstance(engine.compile_tree({(0, 1): {"None": {"nope": ["pearson"]}}}).error, KeyError)
    assert "invalid reducer" in str(engine.compile_tree({(0, 1): {"None": {"mean": ["pearson"]}}}).error)
    # a step whose tree the kernels cover is ours even without the reference installed
    step = pipe.init_step("extractmulti_cells", {"tree": ours, "kwargs": {"ncores": None}})
    assert step.func.__module__ == "aliby_b200.extract" and step.keywords["measure_fn"].__name__ == "extract_tree_multi"


def _golden_profile_state():
    """(state, pipeline, golden) of tests/golden/profiles.npz: the (instructions, results) pairs the reference produced."""
    from itertools import product

    from conftest import load_golden

    g = load_golden("profiles.npz")
    tree = {(int(k) if k.lstrip("-").isdigit() else k): v for k, v in json.loads(str(g["tree_a"])).items()}
    insts = [(ch, red, m) for ch, reds in tree.items() for red, ms in reds.items() for m in ms]
    state = {"data": {"extract_nuclei": [], "extract_cell": [], "extractmulti_nuclei": []}}
    for tp in range(3):
        for step, key in (("extract_nuclei", f"labels{tp}"), ("extract_cell", f"labels_cell{tp}")):
            n_obj = int(g[key].max())
            items = tuple(product([(0, lab) for lab in range(1, n_obj + 1)], insts))
            state["data"][step].append((items, [float(v) for v in g[f"state_{step}_{tp}_values"]]))
        n_obj = int(g[f"labels{tp}"].max())
        multi_items = tuple(((0, lab), ((0, 1), "None", "max", "pearson")) for lab in range(1, n_obj + 1))
        multi_res = [{"Correlation_Pearson": np.array([v])} for v in g[f"state_extractmulti_nuclei_{tp}_values"]]
        state["data"]["extractmulti_nuclei"].append((multi_items, multi_res))
    pipeline = {"steps": {"tile": {}, "segment_nuclei": {}, "extract_nuclei": {}, "extract_cell": {}, "extractmulti_nuclei": {}}}
    return state, pipeline, g, tree


def _assert_table_equals_golden(table, g):
    """Names, order and types of the columns exactly; rows as a set keyed by (tp, tile, object, label): the reference's
    pyarrow join (pipe_core.py:506-510) does not define a row order."""
    cols = json.loads(str(g["columns"]))
    types = json.loads(str(g["types"]))
    assert table.column_names == cols
    assert [str(t) for t in table.schema.types] == types
    keys = [cols.index(f"metadata_{k}") for k in ("tp", "tile", "object", "label")]

    def rows_of(columns, nulls):
        n = len(columns[0])
        rows = {}
        for i in range(n):
            vals = tuple(None if nulls[j][i] else (columns[j][i].item() if hasattr(columns[j][i], "item") else columns[j][i])
                         for j in range(len(cols)))
            rows[tuple(vals[k] for k in keys)] = vals
        assert len(rows) == n  # keys are unique
        return rows

    got_cols = [table.column(c).to_pylist() for c in cols]
    got = rows_of(got_cols, [[v is None for v in col] for col in got_cols])
    want = rows_of([g[f"col{j}"] for j in range(len(cols))], [g[f"null{j}"] for j in range(len(cols))])
    assert got.keys() == want.keys()
    for k, w in want.items():
        for c, a, b in zip(cols, got[k], w):
            assert (a == b) or (a is None and b is None) or (a != a and b != b), (k, c, a, b)


def test_profiles_match_the_reference_table(tmp_path):
    """pipe.get_profiles_from_state and write_profiles against the table the REFERENCE's get_profiles_from_state
    (pipe_core.py:453-512, exec'd from source by oracle/make_golden.py) built from the same state: names, order,
    types, values, nulls of the extract/extractmulti join, and the zstd Parquet schema (pipe_core.py:411-413)."""
    import pyarrow.parquet as pq

    from aliby_b200 import pipe

    state, pipeline, g, _ = _golden_profile_state()
    table = pipe.get_profiles_from_state(state, pipeline)
    assert table.num_rows == int(g["n_rows"])
    _assert_table_equals_golden(table, g)
    path = pipe.write_profiles(table, tmp_path, "pos001")
    back = pq.read_table(path)
    assert back.column_names == json.loads(str(g["columns"]))
    assert [str(t) for t in back.schema.types] == json.loads(str(g["parquet_types"]))
    assert pq.ParquetFile(path).metadata.row_group(0).column(0).compression == "ZSTD"


def test_dense_profile_path_matches_the_reference_table():
    """profiles_from_tables (dense ExtractionTable per time point, no per-item lists) reproduces the reference's rows of
    one extract step: same columns, types and values as the golden profile table restricted to that object."""
    from aliby_b200 import pipe
    from aliby_b200.extract import ExtractionTable

    state, pipeline, g, tree = _golden_profile_state()
    insts = [(ch, red, m) for ch, reds in tree.items() for red, ms in reds.items() for m in ms]
    names = ["/".join(str(x) for x in i) + f"/{i[-1]}" for i in insts]
    tables = []
    for tp in range(3):
        items, res = state["data"]["extract_nuclei"][tp]
        n_obj = len(items) // len(insts)
        objs = np.array([(0, lab) for lab in range(1, n_obj + 1)], dtype=np.int64).reshape(n_obj, 2)
        tables.append(ExtractionTable(objs, names, np.asarray(res, dtype=float).reshape(n_obj, len(insts))) if n_obj else None)
    got = pipe.profiles_from_tables(tables, "nuclei")
    # the golden rows of this object, in the golden's order
    cols = json.loads(str(g["columns"]))
    obj_col = g[f"col{cols.index('metadata_object')}"]
    sel = np.flatnonzero(obj_col == "nuclei")
    assert got.num_rows == len(sel)
    for c in got.column_names:
        j = cols.index(c)
        want = g[f"col{j}"][sel]
        for a, b in zip(got.column(c).to_pylist(), want):
            assert (a == b) or (a != a and b != b), (c, a, b)
        assert str(got.schema.field(c).type) == json.loads(str(g["types"]))[j], c


def test_reciprocal_row_index():
    """object_sweep.cu recovers the window row of a list entry as umulhi(off, 2^32 / pitch + 1): exact for every offset of
    a 64-row window at every pitch the kernel uses (16 ... 144 bytes)."""
    for pitch in range(16, 145, 16):
        inv = 0xFFFFFFFF // pitch + 1
        off = np.arange(64 * pitch, dtype=np.uint64)
        assert np.array_equal((off * np.uint64(inv)) >> np.uint64(32), off // np.uint64(pitch)), pitch
