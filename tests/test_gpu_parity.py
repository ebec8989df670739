"""Parity of the CUDA hot path (through the C-ABI) with the reference-pinned oracle.

Integer-derived outputs (area, total, total_squared, median, max2p5pc, max5px_median, mean,
centroid, bbox, max/min, minor/major axis, volume, eccentricity) must be bit-exact; the
fp64 outputs whose summation order differs from NumPy's (std, moment_of_inertia,
conical_volume, spherical_volume) must agree within the north star's 1e-6 relative — the
tests ask for 1e-9.
"""

import json

import numpy as np
import pytest

from conftest import as_float_pairs, assert_same, golden_tree, load_golden

pytestmark = pytest.mark.gpu

LOOSE = {"std", "moment_of_inertia", "conical_volume", "spherical_volume"}
RTOL_LOOSE = 1e-9

SHAPE = ["area", "centroid", "centroid_x", "centroid_y", "conical_volume", "eccentricity",
         "min_maj_approximation", "spherical_volume", "volume"]
BBOX = ["bbox_rmin", "bbox_rmax", "bbox_cmin", "bbox_cmax"]  # extensions: inclusive bounding box, NaN for absent ids
INTENSITY = ["mean", "std", "median", "total", "total_squared", "max2p5pc", "max5px_median",
             "moment_of_inertia", "ratio"]


@pytest.fixture(scope="module")
def ab():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from aliby_b200 import extract

    return extract


def check_items(items, got, want_a, want_b=None):
    ga, gb = as_float_pairs(got)
    metrics = np.array([it[1][2] for it in items])
    loose = np.isin(metrics, list(LOOSE))
    assert_same(ga[~loose], np.asarray(want_a)[~loose], 0.0, "exact metrics")
    assert_same(ga[loose], np.asarray(want_a)[loose], RTOL_LOOSE, "fp64 metrics")
    if want_b is not None:
        assert_same(gb, want_b, 0.0, "tuple second slot")


def against_oracle(ab, tree, masks, pixels, oracle=None):
    from oracle import fast

    oracle = oracle or fast
    items, got = ab.process_tree_masks(tree, masks, pixels, ab.extract_tree)
    o_items, want = oracle.run_tree(tree, masks, pixels)
    assert len(items) == len(o_items)
    assert all(a[0] == tuple(b[0]) and tuple(a[1]) == tuple(b[1]) for a, b in zip(items[:200], o_items[:200]))
    wa, wb = as_float_pairs(want)
    check_items(items, got, wa, wb)
    return items, got


# ---------------------------------------------------------------- golden vectors (real reference)
def test_golden_field_small(ab):
    g = load_golden("field_small.npz")
    items, got = ab.process_tree_masks(golden_tree(g), g["labels"], g["pixels"], ab.extract_tree)
    assert len(items) == int(g["n_items"])
    check_items(items, got, g["values"], g["values2"])


def test_golden_tiles_list_and_table(ab):
    g = load_golden("tiles_list.npz")
    masks = [m for m in g["labels"]]
    items, got = ab.process_tree_masks(golden_tree(g), masks, g["pixels"], ab.extract_tree)
    assert [it[0][0] for it in items] == g["item_tile"].tolist()
    assert [it[0][1] for it in items] == g["item_label"].tolist()
    check_items(items, got, g["values"])
    table = ab.format_extraction((items, got))
    assert table.column_names == json.loads(str(g["table_columns"]))
    assert [str(t) for t in table.schema.types] == json.loads(str(g["table_types"]))
    assert table.column("tile").to_pylist() == g["table_tile"].tolist()
    assert table.column("label").to_pylist() == g["table_label"].tolist()
    vals = np.stack([np.asarray(table.column(c).to_pylist(), dtype=float) for c in table.column_names[2:]], axis=1)
    loose_cols = np.array([c.split("/")[-1] in LOOSE for c in table.column_names[2:]])
    assert_same(vals[:, ~loose_cols], g["table_values"][:, ~loose_cols], 0.0, "table exact")
    assert_same(vals[:, loose_cols], g["table_values"][:, loose_cols], RTOL_LOOSE, "table fp64")
    # the generic (non-dense) formatter must build the same table from plain python lists
    table2 = ab.format_extraction((items, list(got)))
    assert table2.schema.equals(table.schema) and table2.num_rows == table.num_rows
    for c in table.column_names:
        assert_same(table2.column(c).to_pylist(), table.column(c).to_pylist(), 0.0, c)


def test_golden_volume_shapes(ab):
    """Reference outputs for the analytic shapes of tests/extraction/test_volume.py."""
    from oracle.make_golden import numpy_disk, numpy_ellipse

    g = load_golden("volume_shapes.npz")
    tree = {"None": {"None": ["min_maj_approximation", "volume", "eccentricity", "conical_volume"]}}
    for kind, x, ecc, rot, want in zip(g["kind"], g["x"], g["ecc"], g["rot"], g["out"]):
        if kind == "disk":
            m = numpy_disk(int(x))
        else:
            y = int(np.round(np.sqrt(x**2 / (1 - ecc**2))))
            m = numpy_ellipse(int(x), y, int(rot))
        _, got = ab.process_tree_masks(tree, m.astype(np.uint16), np.zeros((1, 1, 1, *m.shape), np.uint16), ab.extract_tree)
        assert got[0] == (want[0], want[1]), (kind, x, ecc, rot, got[0], want[:2])
        assert got[1] == want[2] and got[2] == want[3]
        assert abs(got[3] - want[4]) <= RTOL_LOOSE * abs(want[4])
        if kind == "disk":  # the reference's own analytic 1 % bound
            real_v = 4 * np.pi * float(x) ** 3 / 3
            assert abs(got[1] - real_v) / real_v < 0.01


def test_golden_degenerate_shapes(ab):
    g = load_golden("degenerate_shapes.npz")
    tree = {"None": {"None": ["min_maj_approximation", "volume", "eccentricity", "conical_volume", "area"]}}
    lab = g["labels"]
    items, got = ab.process_tree_masks(tree, lab, np.zeros((1, 1, 1, *lab.shape), np.uint16), ab.extract_tree)
    got = np.array([[*got[5 * k][:2], got[5 * k + 1], got[5 * k + 2], got[5 * k + 3], got[5 * k + 4]] for k in range(9)])
    assert_same(got, g["out"], 0.0, "degenerate shapes")


def test_golden_background(ab):
    g = load_golden("background.npz")
    tree = {0: {"max": ["imBackground", "background_max5"]}}
    items, got = ab.process_tree_masks(tree, g["labels"], g["image"][None, None, None], ab.extract_tree)
    assert got[0] == float(g["imBackground"]) and got[1] == float(g["background_max5"])
    # registry view with the reference's (Y, X, N) calling convention (trap.py:6-43)
    lab = g["labels"]
    stack = np.stack([lab == k for k in range(1, int(lab.max()) + 1)], axis=2)
    assert ab.TRAP_FUNS["imBackground"](stack, g["image"]) == float(g["imBackground"])
    assert ab.TRAP_FUNS["background_max5"](stack, g["image"]) == float(g["background_max5"])


def test_golden_overlap(ab):
    from functools import partial

    g = load_golden("overlap.npz")
    masks = [m for m in g["masks"]]
    items, got = ab.process_tree_masks_overlap(golden_tree(g), masks, g["pixels"], partial(ab.extract_tree, overlap=True))
    assert np.array_equal(np.array([list(it[0]) for it in items]), g["item_ids"])
    assert_same(got, g["values"], 0.0, "overlap")


def test_overlap_non_sequential_ids_like_the_reference(ab):
    """The live overlap path with ids that are NOT sequential (values from the real reference, run in the build
    container): for k distinct ids the reference enumerates 1..k and reads the plane of the id itself — id 1 absent (area
    0, mean NaN), id 2 measured, id 5 never looked at; no error."""
    from functools import partial

    m = np.zeros((1, 12, 12), np.uint16)
    m[0, 1:4, 1:4] = 2
    m[0, 6:10, 5:9] = 5
    px = np.arange(2 * 12 * 12, dtype=np.uint16).reshape(1, 2, 1, 12, 12)
    tree = {"None": {"None": ["area"]}, 0: {"max": ["mean"]}}
    items, got = ab.process_tree_masks_overlap(tree, [m], px, partial(ab.extract_tree, overlap=True))
    assert [tuple(i[0]) for i in items] == [(0, 0, 1), (0, 0, 1), (0, 0, 2), (0, 0, 2)]
    assert got[0] == 0.0 and got[1] != got[1] and got[2] == 9.0 and got[3] == 26.0


# ---------------------------------------------------------------- oracle comparisons on synthetic fields
def test_config_c1(ab):
    """BASELINE.json configs[0]: 2 channels x 1080^2, ~300 objects, intensity + sizeshape."""
    from aliby_b200 import synth

    pixels, labels = synth.make_field(synth.CONFIG_SEEDS["C1"], (1080, 1080), 2, 300)
    tree = {"None": {"None": ["area", "centroid_x", "centroid_y", "eccentricity", "volume", "conical_volume"] + BBOX},
            0: {"max": ["mean", "std", "median", "total", "max2p5pc", "max5px_median", "max", "min"]},
            1: {"max": ["mean", "std", "median", "total", "total_squared", "max2p5pc", "max5px_median", "moment_of_inertia"]}}
    items, got = against_oracle(ab, tree, labels, pixels)
    assert len(items) == int(labels.max()) * 26


def test_config_c2_full_size(ab):
    """BASELINE.json configs[1]: 5 channels x 2160^2, ~2k cells, full cell-function set."""
    from aliby_b200 import synth

    pixels, labels = synth.make_field(synth.CONFIG_SEEDS["C2"], (2160, 2160), 5, 2000)
    tree = {"None": {"None": SHAPE + BBOX}}
    for ch in range(5):
        tree[ch] = {"max": INTENSITY + ["max", "min"]}
    items, got = against_oracle(ab, tree, labels, pixels)
    # size-independent properties
    a, _ = as_float_pairs(got)
    metrics = np.array([it[1][2] for it in items])
    chans = np.array([str(it[1][0]) for it in items])
    assert a[metrics == "area"].sum() == np.count_nonzero(labels)
    for ch in range(5):
        tot = a[(metrics == "total") & (chans == str(ch))].sum()
        assert tot == pixels[0, ch, 0][labels > 0].astype(np.int64).sum()


def test_z_stack_max_and_add(ab):
    """configs[3] shape at a reduced size: Z = 16, reductions max and add fused into the load."""
    from aliby_b200 import synth

    pixels, labels = synth.make_field(1004, (256, 320), 3, 40, n_z=16)
    tree = {0: {"max": INTENSITY}, 1: {"add": INTENSITY}, 2: {"add": ["mean", "median"], "max": ["median", "max2p5pc"]}}
    against_oracle(ab, tree, labels, pixels)


def test_uint8_pixels_and_odd_width(ab):
    rng = np.random.default_rng(3)
    from aliby_b200 import synth

    labels = synth.ellipse_labels(rng, (101, 203), 25, semi_axes=(4, 14))
    pixels = rng.integers(0, 256, size=(1, 2, 2, 101, 203)).astype(np.uint8)
    tree = {"None": {"None": ["area", "centroid_x", "centroid_y", "eccentricity"]},
            0: {"max": INTENSITY}, 1: {"add": INTENSITY}}
    against_oracle(ab, tree, labels, pixels)


def test_wide_value_range_needs_refinement(ab):
    """Objects whose values span the whole uint16 range force the multi-level radix select."""
    rng = np.random.default_rng(9)
    labels = np.zeros((200, 300), np.uint16)
    labels[10:90, 10:140] = 1     # 10 400 px  (> shared-memory capacity: window re-scan path)
    labels[100:150, 20:60] = 2    # 2 000 px
    labels[160:163, 5:9] = 3
    labels[100:190, 200:290] = 5  # id 4 absent
    pixels = rng.integers(0, 65536, size=(1, 2, 3, 200, 300)).astype(np.uint16)
    pixels[0, 0, :, 160:163, 5:9] = 77  # constant object
    tree = {0: {"max": INTENSITY}, 1: {"add": INTENSITY, "max": ["median", "max2p5pc", "max5px_median"]}}
    against_oracle(ab, tree, labels, pixels)


def test_whole_plane_object_and_large_edt(ab):
    """One label covering almost the whole plane: non-compacted statistics + global-scratch EDT."""
    rng = np.random.default_rng(4)
    labels = np.ones((150, 180), np.uint16)
    labels[0, :] = 0
    labels[40:60, 50:90] = 2
    pixels = rng.integers(0, 5000, size=(1, 1, 1, 150, 180)).astype(np.uint16)
    tree = {"None": {"None": SHAPE}, 0: {"max": INTENSITY + ["imBackground", "background_max5"]}}
    against_oracle(ab, tree, labels, pixels)


def test_empty_inputs(ab):
    tree = {"None": {"None": ["area"]}, 0: {"max": ["mean"]}}
    px = np.zeros((1, 1, 1, 32, 32), np.uint16)
    items, got = ab.process_tree_masks(tree, np.zeros((32, 32), np.uint16), px, ab.extract_tree)
    assert items == () and list(got) == []
    items, got = ab.process_tree_masks(tree, [np.zeros((0,), np.uint16)], px, ab.extract_tree)
    assert items == () and list(got) == []
    table = ab.format_extraction((items, got))
    assert table.num_rows == 0 and table.column_names == ["tile", "label"]


def test_error_behaviour(ab):
    px = np.zeros((1, 1, 2, 16, 16), np.uint16)
    lab = np.zeros((16, 16), np.uint16)
    lab[2:5, 2:5] = 1
    with pytest.raises(KeyError):
        ab.process_tree_masks({0: {"max": ["not_a_metric"]}}, lab, px, ab.extract_tree)
    with pytest.raises(KeyError):
        ab.process_tree_masks({0: {"nope": ["mean"]}}, lab, px, ab.extract_tree)
    with pytest.raises(Exception, match="invalid reducer"):
        ab.process_tree_masks({0: {"mean": ["mean"]}}, lab, px, ab.extract_tree)
    with pytest.raises(Exception, match="invalid reducer"):
        ab.process_tree_masks({0: {"None": ["mean"]}}, lab, px, ab.extract_tree)
    with pytest.raises(NotImplementedError):
        ab.process_tree_masks({0: {"max": ["mean"]}}, lab, px.astype(np.int32), ab.extract_tree)
    # no objects -> the reference never reaches the lookups, neither do we
    items, got = ab.process_tree_masks({0: {"max": ["not_a_metric"]}}, np.zeros_like(lab), px, ab.extract_tree)
    assert items == ()


def test_extract_tree_arbitrary_subset(ab):
    """extract_tree called directly with a shuffled subset of items (extract.py:304-375 contract)."""
    from oracle import fast

    g = load_golden("tiles_list.npz")
    masks = [m for m in g["labels"]]
    tree = golden_tree(g)
    o_items, o_res = fast.run_tree(tree, masks, g["pixels"])
    rng = np.random.default_rng(0)
    pick = rng.permutation(len(o_items))[:97]
    sub = tuple(o_items[i] for i in pick)
    got = ab.extract_tree(sub, masks, g["pixels"], ncores=None)
    check_items(sub, got, as_float_pairs([o_res[i] for i in pick])[0])


def test_registry_single_mask_calls(ab):
    """CELL_FUNS[name](mask, pixels) — the registry contract of loaders.py:28-79."""
    from oracle import port

    rng = np.random.default_rng(2)
    yy, xx = np.mgrid[0:40, 0:50]
    mask = ((yy - 18) ** 2 / 90.0 + (xx - 22) ** 2 / 200.0) <= 1.0
    img = rng.integers(100, 3000, size=(40, 50)).astype(np.uint16)
    for name in ["area", "mean", "median", "max2p5pc", "max5px_median", "total", "centroid_x", "centroid_y",
                 "volume", "eccentricity", "total_squared"]:
        want = port.CELL_METRICS[name](mask, img.copy())
        assert ab.CELL_FUNS[name](mask, img) == float(want), name
    assert ab.CELL_FUNS["centroid"](mask, None) == tuple(float(v) for v in port.m_centroid(mask))
    assert set(port.CELL_METRICS) <= set(ab.CELL_FUNS)


def test_fused_tile_crop_matches_reference_crop(ab):
    """configs[2] (yeast traps): extraction straight out of the frame == crop (tiler.py) then extract."""
    from aliby_b200 import synth
    from aliby_b200.tile import TileView, tile_origins
    from oracle import fast, port

    frames, centres, labels = synth.make_trap_position(1003, n_tp=2, n_channels=3, frame=(600, 640), n_tiles=12, tile_size=96)
    tree = {"None": {"None": ["area", "volume", "eccentricity", "centroid_x", "centroid_y"]},
            0: {"max": ["mean", "median", "std", "max5px_median", "imBackground"]},
            2: {"max": ["median", "background_max5"]}}
    for tp in range(2):
        masks = [m for m in labels[tp]]
        crop = port.crop_tiles(frames[tp], centres, (96, 96), (), tp)  # the reference's materialised tiles
        view = TileView(frames[tp], tile_origins(centres, 96), 96)
        assert np.array_equal(np.asarray(view), crop)  # abx_crop_tiles == tiler.py crop
        items, got = ab.process_tree_masks(tree, masks, view, ab.extract_tree)
        o_items, want = fast.run_tree(tree, masks, crop)
        check_items(items, got, as_float_pairs(want)[0])


def test_extract_table_pipelined_matches_item_api(ab):
    """The chunked, copy/compute-overlapped public entry returns the same numbers as the
    reference-shaped item API, for pinned and pageable host arrays and any chunking."""
    import torch

    from aliby_b200 import synth

    fields = [synth.make_field(70 + i, (160, 224), 3, 18 + 3 * i, semi_axes=(4, 12)) for i in range(5)]
    pixels = np.concatenate([f[0] for f in fields])
    masks = [f[1] for f in fields]
    masks[3] = np.zeros_like(masks[3])  # a tile without any cell
    tree = {"None": {"None": ["area", "centroid_x", "eccentricity", "volume"]},
            0: {"max": ["mean", "median", "max2p5pc"]}, 2: {"max": ["std", "max5px_median", "total"]}}
    items, want = ab.process_tree_masks(tree, masks, pixels, ab.extract_tree)
    want = np.asarray(want, dtype=float).reshape(-1, 10)
    pinned = torch.from_numpy(pixels).pin_memory().numpy()
    for px, chunk in [(pixels, 1), (pinned, 1), (pinned, 3 * pixels[0].nbytes), (pixels, 1 << 30)]:
        tab = ab.extract_table(tree, masks, px, chunk_bytes=chunk)
        assert tab.values.shape == want.shape
        assert_same(tab.values, want, 0.0, f"chunk_bytes={chunk}")
        assert [tuple(o) for o in tab.objects.tolist()] == [it[0] for it in items[::10]]
    assert tab.to_arrow().num_rows == want.shape[0]


def _float_case(ab, pixels, labels, tree, rtol, atol=1e-12):
    from oracle import fast

    items, got = ab.process_tree_masks(tree, labels, pixels, ab.extract_tree)
    o_items, want = fast.run_tree(tree, labels, pixels)
    assert [tuple(i[1]) for i in items] == [tuple(i[1]) for i in o_items]
    ga, _ = as_float_pairs(got)
    wa, _ = as_float_pairs(want)
    metrics = np.array([it[1][2] for it in items])
    exact = np.isin(metrics, ["median", "max", "min", "imBackground"])  # order statistics: bit-exact
    assert_same(ga[exact], wa[exact], 0.0, "float order statistics")
    # fp64 sums: relative tolerance, plus an absolute floor for quantities that are rounding noise in NumPy
    # itself (the std of a constant object is ~1e-17 instead of 0)
    g, w_ = ga[~exact], wa[~exact]
    noise = np.abs(w_) < atol
    assert_same(np.where(noise, 0.0, g), np.where(noise, 0.0, w_), rtol, "float sums")
    assert (np.abs(g[noise & ~np.isnan(w_)]) < atol).all()
    return items, ga


FLOAT_METRICS = ["mean", "std", "median", "total", "total_squared", "max2p5pc", "max5px_median", "moment_of_inertia",
                 "max", "min"]


def test_float64_pixels_cropTiler_standard_scale(ab):
    """C1 variant with float64 pixels: what CropTiler's standard_scale hands to extraction (tiler.py:95-102)."""
    from aliby_b200 import synth

    px16, labels = synth.make_field(1011, (540, 600), 2, 90, n_z=2)
    pix = px16.astype(np.float64)
    mean = pix.mean(axis=(-3, -2, -1))
    std = pix.std(axis=(-3, -2, -1))
    pixels = ((pix.T - mean.T) / std.T).T  # per (tile, channel) standardisation, negative and positive values
    tree = {"None": {"None": ["area", "centroid_x", "eccentricity"]},
            0: {"max": FLOAT_METRICS + ["imBackground", "background_max5"]}, 1: {"add": FLOAT_METRICS}}
    _float_case(ab, pixels, labels, tree, 1e-9)


def test_float32_pixels_and_nan_tile(ab):
    """float32 pixels (NumPy computes in float32: agreement within 1e-5) and a NaN-poisoned object
    (NaN tiles of tiler.py:644-646): every statistic of an object that holds a NaN is NaN."""
    from aliby_b200 import synth

    px16, labels = synth.make_field(1012, (200, 260), 1, 25, semi_axes=(4, 12))
    pixels = (px16.astype(np.float32) / np.float32(7.0))
    tree = {0: {"max": ["mean", "median", "max", "min", "total", "std", "max2p5pc"]}}
    # float32 arithmetic inside NumPy: 1e-7 relative noise on values of ~1e4 (std of a constant object ~1e-3)
    _float_case(ab, pixels, labels, tree, 2e-5, atol=5e-3)
    poisoned = pixels.astype(np.float64)
    rr, cc = np.nonzero(labels == 3)
    poisoned[0, 0, 0, rr[0], cc[0]] = np.nan
    items, ga = _float_case(ab, poisoned, labels, tree, 1e-9)
    rows = np.array([it[0][1] == 3 for it in items])
    assert np.isnan(ga[rows]).all() and not np.isnan(ga[~rows]).any()


def test_div_reducer_on_integer_pixels(ab):
    """`div` = np.divide.reduce over Z: float64 values out of uint16 pixels, next to integer requests of the same tree."""
    from aliby_b200 import synth

    pixels, labels = synth.make_field(1013, (180, 240), 2, 20, n_z=3, semi_axes=(4, 12))
    pixels = np.maximum(pixels, 1)  # no division by zero
    tree = {0: {"div": ["mean", "median", "std", "total", "max2p5pc", "max"], "max": ["mean", "median", "total"]},
            1: {"add": ["median", "total"], "div": ["median"]}}
    _float_case(ab, pixels, labels, tree, 1e-9)


def test_config_c4_zstack_full_size(ab):
    """BASELINE.json configs[3]: 5 channels x 16 z x 2048^2, Z reduced inside the gather ('max' and one 'add' channel)."""
    from aliby_b200 import synth

    rng = np.random.default_rng(synth.CONFIG_SEEDS["C4"])
    labels = synth.ellipse_labels(rng, (2048, 2048), 1500)
    # cheap synthetic stack (the Poisson generator of synth needs minutes at this size)
    base = rng.integers(100, 4000, size=(5, 1, 2048, 2048), dtype=np.uint16)
    zmod = rng.integers(0, 3000, size=(5, 16, 1, 1), dtype=np.uint16)
    noise = rng.integers(0, 64, size=(5, 16, 2048, 2048), dtype=np.uint16)
    pixels = (base + zmod + noise)[None]
    tree = {"None": {"None": ["area", "centroid_x", "centroid_y"]}}
    for ch in range(4):
        tree[ch] = {"max": ["mean", "median", "max2p5pc", "std", "max"]}
    tree[4] = {"add": ["mean", "median", "total", "max5px_median"]}
    items, got = against_oracle(ab, tree, labels, pixels)
    a, _ = as_float_pairs(got)
    metrics = np.array([it[1][2] for it in items])
    chans = np.array([str(it[1][0]) for it in items])
    tot = a[(metrics == "total") & (chans == "4")].sum()
    assert tot == pixels[0, 4].astype(np.uint64).sum(axis=0)[labels > 0].sum()


def test_config_c3_trap_position_real_geometry(ab):
    """BASELINE.json configs[2]: 1200^2 frames, 40 tiles of 96^2, 5 channels; extraction fused with the tile crop,
    time point by time point, with per-tile background metrics (global_settings.py:37-53 feature set)."""
    from aliby_b200.tile import FusedTiler
    from aliby_b200 import synth
    from oracle import fast, port

    frames, centres, labels = synth.make_trap_position(synth.CONFIG_SEEDS["C3"], n_tp=3)
    tiler = FusedTiler(frames, centres, 96)
    tree = {"None": {"None": ["area", "volume", "eccentricity", "centroid_x", "centroid_y"]}}
    for ch in range(5):
        tree[ch] = {"max": ["mean", "median", "std", "imBackground", "max5px_median"]}
    for tp in range(3):
        view = tiler.run_tp(tp)["pixels"]
        masks = [m for m in labels[tp]]
        items, got = ab.process_tree_masks(tree, masks, view, ab.extract_tree)
        crop = port.crop_tiles(frames[tp], centres, (96, 96), (), tp)
        o_items, want = fast.run_tree(tree, masks, crop)
        assert [tuple(i[0]) for i in o_items] == [i[0] for i in items]
        check_items(items, got, as_float_pairs(want)[0])


def test_many_tiles_in_one_call_backgrounds_on_their_own_stream(ab):
    """Sixty time points x 40 tiles in ONE call (the batched form of C3: more than 4 096 objects, so the per-plane
    backgrounds run on the second helper stream next to the statistics and the shape chain) against the oracle."""
    from aliby_b200 import synth
    from oracle import fast, port

    T = 60
    frames, centres, labels = synth.make_trap_position(31, n_tp=T, n_channels=2, frame=(700, 720), n_tiles=40, tile_size=96)
    tree = {"None": {"None": ["area", "eccentricity"]}, 0: {"max": ["mean", "median", "imBackground", "background_max5"]},
            1: {"max": ["std", "imBackground"]}}
    masks = [labels[t, i] for t in range(T) for i in range(40)]
    crops = np.concatenate([port.crop_tiles(frames[t], centres, (96, 96), (), t) for t in range(T)])
    assert sum(int(m.max()) for m in masks) > 4096
    table = ab.extract_table(tree, masks, crops)
    o_items, want = fast.run_tree(tree, masks, crops)
    wa = as_float_pairs(want)[0].reshape(len(table.objects), -1)
    names = ["/".join(str(x) for x in i[1]) + "/" + i[1][-1] for i in o_items[: wa.shape[1]]]
    assert names == table.names
    loose = np.array([n.endswith("std") for n in names])
    assert_same(table.values[:, ~loose], wa[:, ~loose], 0.0, "exact metrics")
    assert_same(table.values[:, loose], wa[:, loose], RTOL_LOOSE, "fp64 metrics")


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_random_label_planes(ab, seed):
    """Randomised planes that stress the scan and the window classes: salt-and-pepper labels (several ids inside
    one 8-pixel strip), touching blocks, rings with holes, objects wider than 64 px, id gaps, several tiles with
    different label counts, odd plane sizes, uint8/uint16, Z in {1, 2, 3}, max/add."""
    rng = np.random.default_rng(1000 + seed)
    H, W = int(rng.integers(17, 150)), int(rng.integers(17, 300))
    n_tiles = int(rng.integers(1, 4))
    Z = int(rng.integers(1, 4))
    dtype = np.uint8 if seed % 3 == 0 else np.uint16
    masks = []
    for _ in range(n_tiles):
        lab = np.zeros((H, W), np.uint16)
        next_id = 1
        for _ in range(int(rng.integers(0, 12))):  # blocks, some overlapping (later ones overwrite), some with holes
            h, w = int(rng.integers(1, min(H, 90))), int(rng.integers(1, min(W, 120)))
            r, c = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
            lab[r : r + h, c : c + w] = next_id
            if h > 6 and w > 6 and rng.random() < 0.5:
                lab[r + 2 : r + h - 2, c + 2 : c + w - 2] = 0 if rng.random() < 0.5 else next_id + 1
                next_id += 1
            next_id += int(rng.integers(1, 3))  # id gaps
        speck = rng.random((H, W)) < 0.03  # salt and pepper: many ids per strip, 1-pixel objects
        lab[speck] = rng.integers(1, max(2, next_id + 3), size=int(speck.sum()))
        masks.append(lab)
    pixels = rng.integers(0, np.iinfo(dtype).max + 1, size=(n_tiles, 2, Z, H, W)).astype(dtype)
    tree = {"None": {"None": ["area", "centroid_x", "centroid_y", "eccentricity", "volume", "conical_volume",
                              "min_maj_approximation"] + BBOX},
            0: {"max": INTENSITY + ["max", "min", "imBackground", "background_max5"]},
            1: {"add": ["mean", "median", "total", "total_squared", "max2p5pc", "max5px_median", "std"]}}
    against_oracle(ab, tree, masks if n_tiles > 1 else masks[0], pixels)


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_tma_staged_windows(ab, seed):
    """Planes whose layout qualifies for the TMA kernels (16-byte aligned rows, Z = 1): windows of every width up to 64
    at every column alignment (58..64 columns wide at a bad alignment = the left-over path), cells that fill the slot
    (512-bin histogram), full-range values (refinement sweeps), rings and ellipses (rows with several runs), thin and
    square blocks (cone tops of 1, a few, more than 32 pixels), uint8 and uint16."""
    from aliby_b200 import synth

    rng = np.random.default_rng(7000 + seed)
    H, W = int(rng.integers(64, 170)), 16 * int(rng.integers(4, 20))
    n_tiles = int(rng.integers(1, 4))
    dtype = np.uint8 if seed % 2 else np.uint16
    top = np.iinfo(dtype).max
    masks = []
    for _ in range(n_tiles):
        lab = synth.ellipse_labels(rng, (H, W), int(rng.integers(0, 10)), semi_axes=(3, 31)).astype(np.uint16)
        next_id = int(lab.max()) + 1
        for _ in range(int(rng.integers(2, 10))):
            h, w = int(rng.integers(1, 65)), int(rng.choice([1, 2, 3, 7, 30, 33, 57, 58, 60, 63, 64, int(rng.integers(1, 65))]))
            h, w = min(h, H), min(w, W)
            r, c = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
            lab[r : r + h, c : c + w] = next_id
            if h > 8 and w > 8 and rng.random() < 0.4:  # ring
                lab[r + 3 : r + h - 3, c + 3 : c + w - 3] = 0
            next_id += int(rng.integers(1, 3))
        masks.append(lab)
    Z = 1 if seed < 6 else int(rng.integers(2, 5))  # Z stacks: max requests are reduced up front (zreduce.cu), add are not
    if seed % 3 == 0:   # noise around a per-pixel level: narrow ranges, one-pass histogram
        base = rng.integers(0, max(2, top // 2), size=(n_tiles, 3, 1, 1, 1))
        pixels = np.clip(base + rng.poisson(40, size=(n_tiles, 3, Z, H, W)), 0, top).astype(dtype)
    else:
        pixels = rng.integers(0, top + 1, size=(n_tiles, 3, Z, H, W)).astype(dtype)
    tree = {"None": {"None": ["area", "centroid_x", "centroid_y", "eccentricity", "volume", "conical_volume",
                              "min_maj_approximation"] + BBOX},
            0: {"max": INTENSITY + ["max", "min"]},
            1: {"max": ["median", "mean"], "add": ["total", "median"]},
            2: {"add": ["mean", "median", "total", "total_squared", "max2p5pc", "max5px_median", "std", "moment_of_inertia"]}}
    against_oracle(ab, tree, masks if n_tiles > 1 else masks[0], pixels)


def test_border_cut_cell_long_cone_top(ab):
    """A cell cut straight by the image border: the EDT maximum is a line of 34 pixels (> 32: the plateau path of the
    shape kernel, which once took 0.3 ms on one warp and set the step time of a whole batch)."""
    yy, xx = np.mgrid[0:96, 0:128]
    labels = np.zeros((96, 128), np.uint16)
    labels[14:70, 0:10] = 1                                               # a 56 x 10 cell on the left border ...
    labels[14:20, 6:10] = 0; labels[64:70, 5:10] = 0                      # ... with trimmed corners
    labels[((yy - 70) / 6.0) ** 2 + ((xx - 100) / 27.0) ** 2 <= 1.0] = 2  # flat ellipse: a horizontal line of maxima
    labels[10:14, 40:104] = 3                                             # 4 x 64 bar
    rng = np.random.default_rng(5)
    pixels = rng.integers(0, 4000, size=(1, 1, 1, 96, 128)).astype(np.uint16)
    tree = {"None": {"None": SHAPE}, 0: {"max": INTENSITY}}
    against_oracle(ab, tree, labels, pixels)


def test_out_of_frame_tiles_golden_and_extraction(ab):
    """if_out_of_bounds_pad (tiler.py:601-650): median-padded and NaN tiles, against the golden crops written by the
    real reference, then extraction through padded / NaN tiles against the oracle on the reference-shaped crop."""
    import torch

    from aliby_b200.tile import TileView, tile_origins
    from oracle import fast, port

    g = load_golden("tile_crop.npz")
    frame, centres, drifts = g["frame"], g["centres"], g["drifts"]  # frame: (Z, Y, X) of one channel
    for tp in (0, 1):
        view = TileView(frame[None], tile_origins(centres, 16, drifts, tp), 16)
        tiles = view.materialize().cpu().numpy()  # (tiles, 1, Z, 16, 16)
        for i in range(len(centres)):
            want = g[f"tp{tp}_tile{i}"]
            assert_same(tiles[i, 0], want, 0.0, f"tp{tp} tile{i}")
        assert (tiles.dtype == np.float64) == any(g[f"tp{tp}_tile{i}"].dtype == np.float64 for i in range(len(centres)))

    # extraction: uint16 frame, tiles hanging over every edge (<= 25 % -> median pad), then one NaN tile too
    rng = np.random.default_rng(21)
    fr = rng.integers(100, 4000, size=(2, 2, 90, 110)).astype(np.uint16)  # (C, Z, H, W)
    size = 32
    lab = np.zeros((size, size), np.uint16)
    lab[3:14, 2:12] = 1
    lab[18:30, 15:31] = 2
    lab[0:6, 24:32] = 3
    tree = {"None": {"None": ["area", "centroid_x"]}, 0: {"max": ["mean", "median", "std", "max2p5pc", "imBackground"]},
            1: {"add": ["total", "median"]}}
    for centres2 in ([(12, 12), (80, 100), (45, 10), (20, 104)], [(12, 12), (80, 100), (2, 50)]):
        centres2 = np.asarray(centres2)
        view = TileView(fr, tile_origins(centres2, size), size)
        crop = port.crop_tiles(fr, centres2, (size, size), (), 0)  # the reference's own (tiles, C, Z, h, w)
        assert view.out_of_frame.any() and (view.nan_tiles.any() == (crop.dtype == np.float64))
        masks = [lab.copy() for _ in centres2]
        items, got = ab.process_tree_masks(tree, masks, view, ab.extract_tree)
        o_items, want = fast.run_tree(tree, masks, crop)
        ga, _ = as_float_pairs(got)
        wa, _ = as_float_pairs(want)
        metrics = np.array([it[1][2] for it in items])
        loose = np.isin(metrics, ["std", "mean", "total", "max2p5pc"]) if crop.dtype == np.float64 else (metrics == "std")
        assert_same(ga[~loose], wa[~loose], 0.0, "padded tiles: exact metrics")
        assert_same(ga[loose], wa[loose], RTOL_LOOSE, "padded tiles: fp64 sums")


def test_overlap_original_ids_and_format(ab):
    """(tile, stack, label) extraction keyed by the ORIGINAL, non-sequential ids — what the reference's
    format_extraction_overlap (extract.py:602-682) was written for — against the oracle on each stack plane."""
    from functools import partial

    from oracle import fast

    rng = np.random.default_rng(8)
    H, W = 48, 64
    pixels = rng.integers(50, 3000, size=(2, 2, 2, H, W)).astype(np.uint16)
    masks = []
    for t in range(2):
        stacks = np.zeros((2, H, W), np.uint16)
        stacks[0, 4:20, 5:25] = 7
        stacks[0, 25:40, 30:60] = 3
        stacks[1, 10:30, 15:40] = 12 if t == 0 else 5  # overlaps stack 0, non-sequential ids
        masks.append(stacks)
    tree = {"None": {"None": ["area", "centroid_x"]}, 0: {"max": ["mean", "median"]}, 1: {"add": ["total"]}}
    items, got, inv = ab.process_tree_masks_overlap(tree, masks, pixels, partial(ab.extract_tree, overlap=True),
                                                    original_ids=True)
    assert [it[0] for it in items[::5]] == [(0, 0, 1), (0, 0, 2), (0, 1, 1), (1, 0, 1), (1, 0, 2), (1, 1, 1)]
    assert inv[(0, 0)].tolist() == [0, 3, 7] and inv[(0, 1)].tolist() == [0, 12] and inv[(1, 1)].tolist() == [0, 5]
    want = []
    for (t, s, j), inst in items:
        o_items, o_res = fast.run_tree({inst[0]: {inst[1]: [inst[2]]}}, masks[t][s], pixels[t : t + 1])
        want.append(float(o_res[int(inv[(t, s)][j]) - 1]))
    assert_same(np.asarray(got, dtype=float), np.asarray(want), 0.0, "overlap, original ids")
    table = ab.format_extraction_overlap((items, got, inv))
    assert table.column_names[:2] == ["metadata_tile", "metadata_label"]
    keys = sorted(zip(table.column("metadata_tile").to_pylist(), table.column("metadata_label").to_pylist()))
    assert keys == [(0, 3), (0, 7), (0, 12), (1, 3), (1, 5), (1, 7)]


def test_background_of_a_whole_field_streaming_path(ab):
    """imBackground / background_max5 (trap.py:6-43) on a whole 2160^2 field: the streaming 65 536-bin path."""
    import time

    from aliby_b200 import synth

    pixels, labels = synth.make_field(1020, (2160, 2160), 3, 400, n_z=2)
    tree = {0: {"max": ["imBackground", "background_max5", "median"]}, 2: {"max": ["background_max5"]},
            1: {"add": ["imBackground"]}}  # the Z-add request stays on the per-object path (wide values)
    t0 = time.perf_counter()
    tab = ab.extract_table(tree, [labels], pixels)
    dt = time.perf_counter() - t0
    bg = labels == 0
    for name, ch, red, fn in [("0/max/imBackground/imBackground", 0, np.maximum, np.median),
                              ("0/max/background_max5/background_max5", 0, np.maximum, lambda v: np.mean(np.sort(v)[-5:])),
                              ("2/max/background_max5/background_max5", 2, np.maximum, lambda v: np.mean(np.sort(v)[-5:])),
                              ("1/add/imBackground/imBackground", 1, np.add, np.median)]:
        want = float(fn(red.reduce(pixels[0, ch], axis=0)[bg]))
        col = tab.values[:, tab.names.index(name)]
        assert (col == want).all(), (name, col[:3], want)
    assert dt < 5.0


def test_stale_label_count_is_reported(ab):
    """A label above the plane's n_labels (a caller's stale plane_base) must not produce a silently wrong table: the
    device error flag travels back with the table (abx_extract_args.status) and raises."""
    import torch

    from aliby_b200 import engine, synth

    pixels, labels = synth.make_field(77, (96, 128), 1, 6, semi_axes=(4, 9))
    plan = engine.compile_tree({"None": {"None": ["area"]}, 0: {"max": ["mean"]}})
    dev = torch.device("cuda", 0)
    lab = torch.from_numpy(labels[None].astype(np.uint16)).to(dev)
    px = torch.from_numpy(pixels).to(dev)
    n_true = int(labels.max())
    for n_given, bad in ((n_true, False), (n_true - 2, True)):
        buf, table, status = engine.alloc_table(n_given, plan.n_columns, dev)
        engine.run_planes(plan, lab, np.zeros(1, np.int32), np.array([n_given]), px, np.zeros(1, np.int64),
                          96 * 128, 96 * 128, 128, 1, 1, out=table, status=status)
        word = int(status.cpu()[0])
        assert bool(word & 1) == bad
        if bad:
            with pytest.raises(IndexError, match="label id exceeds"):
                engine.raise_on_status(word)
    # and through the public API: masks whose shape differs from the pixels' (the reference raises IndexError too)
    with pytest.raises(IndexError):
        ab.process_tree_masks({0: {"max": ["mean"]}}, np.pad(labels, ((0, 8), (0, 0))), pixels, ab.extract_tree)


def test_concurrent_calls_two_threads_two_streams(ab):
    """The host engine is re-entrant: two threads, each on its own stream of one device, extract different fields at
    the same time and both get the oracle's numbers (per-call scratch, no shared item state)."""
    import threading

    import torch

    from aliby_b200 import synth
    from oracle import fast

    tree = {"None": {"None": ["area", "eccentricity"]}, 0: {"max": ["mean", "median", "max2p5pc", "std"]},
            1: {"max": ["total", "max5px_median"]}}
    fields = [synth.make_field(300 + k, (256, 320), 2, 40, semi_axes=(5, 20)) for k in range(2)]
    want = []
    for px, lab in fields:
        _, w = fast.run_tree(tree, lab, px)
        want.append(as_float_pairs(w)[0])
    errors, results = [], [None, None]

    def worker(k):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream()):
                for _ in range(20):
                    items, got = ab.process_tree_masks(tree, fields[k][1], fields[k][0], ab.extract_tree)
                    results[k] = (items, got)
                    check_items(items, got, want[k])
        except Exception:  # noqa: BLE001
            import traceback

            errors.append(traceback.format_exc())

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, "\n".join(errors)


def test_graph_replay_of_repeated_small_calls(ab):
    """The small-call regime of a time-lapse (same shapes at every time point): after two eager calls the extract step
    replays a captured CUDA graph (engine.GraphedExtract).  Different data and different label counts at every call —
    including more labels than the captured per-plane capacity, which rebuilds the graph — must give the oracle's
    numbers, for a dense (tiles, C, Z, Y, X) array and for the fused tile view."""
    from aliby_b200 import extract as ex
    from aliby_b200 import synth
    from aliby_b200.tile import TileView
    from oracle import fast, port

    tree = {"None": {"None": ["area", "eccentricity", "centroid_x"]},
            0: {"max": ["mean", "median", "max2p5pc", "std", "imBackground"]}, 1: {"max": ["total", "max5px_median", "max"]}}
    ex._graph_cache.clear(); ex._graph_seen.clear()
    counts = [6, 9, 4, 7, 40, 5, 12]  # 40 > 2 x 9: above the first graph's capacity
    for k, n in enumerate(counts):
        tiles = [synth.make_field(900 + 10 * k + t, (96, 128), 2, n if t == 0 else 3, semi_axes=(3, 7)) for t in range(3)]
        px = np.concatenate([t[0] for t in tiles])
        masks = [t[1] for t in tiles]
        against_oracle(ab, tree, masks, px)
    assert any(isinstance(g, object) for g in ex._graph_cache.values()) and len(ex._graph_cache) >= 1
    # fused tile crop: frames of one position, tile table fixed
    frames, centres, labels = synth.make_trap_position(77, n_tp=5, n_channels=2, frame=(400, 480), n_tiles=6, tile_size=64)
    org = centres - 32
    ex._graph_cache.clear(); ex._graph_seen.clear()
    for t in range(5):
        view = TileView(frames[t], org, 64)
        masks = [labels[t, i] for i in range(6)]
        items, got = ab.process_tree_masks(tree, masks, view, ab.extract_tree)
        _, want = fast.run_tree(tree, masks, port.crop_tiles(frames[t], centres, (64, 64)))
        check_items(items, got, as_float_pairs(want)[0])
    assert len(ex._graph_cache) == 1


def _cpm_reference(tree, masks, pixels, cp_kwargs):
    """Per (object, instruction) dicts of oracle/cpm.py, in the reference's item order (parity self-defined: the
    cp_measure source is not available)."""
    from oracle import cpm, port

    if not isinstance(masks, list):
        masks = [masks]
    insts = port.tree_instructions(tree)
    out = []
    for tile_i, lab in enumerate(masks):
        for k in range(1, int(lab.max()) + 1 if lab.size else 1):
            for ch, red, metric in insts:
                img = None if ch == "None" else port.project_z(pixels[tile_i, ch], port.Z_REDUCERS[red])
                out.append(cpm.FEATURES[metric](lab == k, img, **cp_kwargs.get(metric, {})))
    return out


def _check_cp(items, got, want):
    assert len(got) == len(want) == len(items)
    loose = {"Intensity_StdIntensity", "AreaShape_MeanRadius", "AreaShape_Eccentricity", "AreaShape_MajorAxisLength",
             "AreaShape_MinorAxisLength", "Intensity_MassDisplacement", "Location_CenterMassIntensity_X",
             "Location_CenterMassIntensity_Y", "AreaShape_Center_X", "AreaShape_Center_Y", "Intensity_MeanIntensity",
             "AreaShape_Extent", "AreaShape_EquivalentDiameter"}
    for it, g, w in zip(items, got, want):
        assert isinstance(g, dict) and list(g) == list(w), it
        for key in w:
            a, b = float(g[key][0]), float(w[key][0])
            if b != b:
                assert a != a, (it, key, a)
            elif key in loose:
                assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (it, key, a, b)
            else:
                assert a == b, (it, key, a, b)


def test_cp_measure_intensity_and_sizeshape(ab):
    """The trees the reference's stock builder emits (pipe_builder.py:115-120: sizeshape on the masks, intensity per
    channel, edge features switched off): dict-valued results with CellProfiler's feature names, checked against
    oracle/cpm.py — absent ids, 1-6 pixel cells, a constant and a saturated cell, narrow and full-range values (the
    radix-select path), uint8 and uint16 — and the wide table the reference's format_extraction builds from them."""
    from aliby_b200 import synth

    kw = {"intensity": {"edge_measurements": False}}
    tree = {"None": {"None": ("sizeshape",)}, 1: {"max": ("intensity",)}, 0: {"max": ("intensity",)}}
    pixels, labels = synth.make_field(4711, (320, 384), 2, 60, semi_axes=(3, 24))
    rng = np.random.default_rng(5)
    cases = [(pixels, labels), (rng.integers(0, 65536, size=pixels.shape).astype(np.uint16), labels),
             (rng.integers(0, 256, size=pixels.shape).astype(np.uint8), labels)]
    for px, lab in cases:
        items, got = ab.process_tree_masks(tree, lab, px, ab.extract_tree, cp_measure_kwargs=kw)
        _check_cp(items, got, _cpm_reference(tree, lab, px, kw))
    table = ab.format_extraction((items, got))
    assert table.num_rows == int(labels.max())
    assert "1/max/intensity/Intensity_MADIntensity" in table.column_names
    assert "None/None/sizeshape/AreaShape_BoundingBoxMaximum_X" in table.column_names
    assert table.column_names[2:] == sorted(table.column_names[2:])
    # the same numbers through the generic (per item) formatter of the reference's contract
    plain = ab.format_extraction((tuple(items), list(got)))
    assert plain.column_names == table.column_names
    for c in table.column_names:
        x, y = table.column(c).to_pylist(), plain.column(c).to_pylist()
        assert all((p == q) or (p != p and q != q) for p, q in zip(x, y)), c
    # mixed with the cell.py functions, several tiles, and the dense fast entry
    tree2 = {"None": {"None": ["area", "sizeshape"]}, 0: {"max": ["mean", "intensity", "median"]}}
    tiles = [synth.make_field(4800 + t, (128, 192), 1, 8, semi_axes=(3, 12)) for t in range(3)]
    px2 = np.concatenate([t[0] for t in tiles])
    masks2 = [t[1] for t in tiles]
    items2, got2 = ab.process_tree_masks(tree2, masks2, px2, ab.extract_tree, cp_measure_kwargs=kw)
    want2 = _cpm_reference({"None": {"None": ["sizeshape"]}, 0: {"max": ["intensity"]}}, masks2, px2, kw)
    dict_items = [(it, g) for it, g in zip(items2, got2) if isinstance(g, dict)]
    _check_cp([it for it, _ in dict_items], [g for _, g in dict_items], want2)
    fast_table = ab.extract_table(tree2, masks2, px2, cp_measure_kwargs=kw)
    assert "0/max/intensity/Intensity_UpperQuartileIntensity" in fast_table.names and len(fast_table.names) == 1 + 15 + 1 + 16 + 1
    # edge features and cp_measure functions without a kernel say so
    with pytest.raises(NotImplementedError, match="edge_measurements"):
        ab.process_tree_masks(tree, labels, pixels, ab.extract_tree)
    with pytest.raises(KeyError, match="radial_zernikes"):
        ab.process_tree_masks({0: {"max": ("radial_zernikes",)}}, labels, pixels, ab.extract_tree)
    # objects above the 64 x 64 window (the CTA-per-object kernel): narrow and full value ranges, odd and even areas, a
    # whole-plane object; plus sizeshape of the same masks
    wide = np.zeros((160, 200), np.uint16)
    wide[20:100, 30:130] = 1
    wide[100:159, 3:190] = 2
    wide[101, 5] = 0  # (an odd area)
    wide[4:12, 150:199] = 4  # id 3 absent
    tree_w = {"None": {"None": ("sizeshape",)}, 0: {"max": ("intensity",)}, 1: {"max": ("intensity", "median")}}
    for px in (pixels[:, :, :, :160, :200], rng.integers(0, 65536, size=(1, 2, 1, 160, 200)).astype(np.uint16),
               rng.integers(0, 256, size=(1, 2, 1, 160, 200)).astype(np.uint8)):
        items_w, got_w = ab.process_tree_masks(tree_w, wide, px, ab.extract_tree, cp_measure_kwargs=kw)
        want_w = _cpm_reference({"None": {"None": ("sizeshape",)}, 0: {"max": ("intensity",)}, 1: {"max": ("intensity",)}}, wide, px, kw)
        dict_w = [(it, g) for it, g in zip(items_w, got_w) if isinstance(g, dict)]
        _check_cp([it for it, _ in dict_w], [g for _, g in dict_w], want_w)
    whole = np.ones((96, 128), np.uint16)
    px_whole = rng.integers(0, 4000, size=(1, 1, 1, 96, 128)).astype(np.uint16)
    items_p, got_p = ab.process_tree_masks({0: {"max": ("intensity",)}}, whole, px_whole, ab.extract_tree, cp_measure_kwargs=kw)
    _check_cp(items_p, got_p, _cpm_reference({0: {"max": ("intensity",)}}, whole, px_whole, kw))
    # rank statistics of a Z stack exist for the `max` reduction only: `add` says so instead of returning them
    stack = rng.integers(0, 3000, size=(1, 1, 3, 96, 128)).astype(np.uint16)
    items_z, got_z = ab.process_tree_masks({0: {"max": ("intensity",)}}, whole, stack, ab.extract_tree, cp_measure_kwargs=kw)
    _check_cp(items_z, got_z, _cpm_reference({0: {"max": ("intensity",)}}, whole, stack, kw))
    with pytest.raises(NotImplementedError, match="add"):
        ab.process_tree_masks({0: {"add": ("intensity",)}}, whole, stack, ab.extract_tree, cp_measure_kwargs=kw)
    px_z, lab_z = synth.make_field(4900, (128, 192), 2, 12, n_z=3, semi_axes=(3, 14))  # window-sized cells of a Z stack
    items_z, got_z = ab.process_tree_masks({1: {"max": ("intensity",)}}, lab_z, px_z, ab.extract_tree, cp_measure_kwargs=kw)
    _check_cp(items_z, got_z, _cpm_reference({1: {"max": ("intensity",)}}, lab_z, px_z, kw))


def _colocalisation_reference(tree, masks, pixels, cp_kwargs=None):
    """Per (object, instruction) dicts of oracle/cpm.py for an extractmulti tree, in the reference's item order
    (extract.py:200-226: both channels Z-reduced, then ``fun(pixels1, pixels2, mask)``)."""
    from oracle import cpm, port

    if not isinstance(masks, list):
        masks = [masks]
    cp_kwargs = cp_kwargs or {}
    insts = port.tree_instructions(tree)
    out = []
    for tile_i, lab in enumerate(masks):
        for k in range(1, int(lab.max()) + 1 if lab.size else 1):
            for (ch0, ch1), red_ch, red, metric in insts:
                assert red_ch == "None"
                a = port.project_z(pixels[tile_i, ch0], port.Z_REDUCERS[red])
                b = port.project_z(pixels[tile_i, ch1], port.Z_REDUCERS[red])
                out.append(cpm.CORRELATIONS[metric](a, b, lab == k, **cp_kwargs.get(metric, {})))
    return out


def _check_colocalisation(items, got, want):
    assert len(got) == len(want) == len(items)
    for it, g, w in zip(items, got, want):
        assert isinstance(g, dict) and list(g) == list(w), it
        for key in w:
            a, b = float(g[key][0]), float(w[key][0])
            if b != b:
                assert a != a, (it, key, a)
            else:  # exact integer sums on the GPU, float64 sums in the oracle: 1e-9 (absolute for a correlation near 0)
                assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (it, key, a, b)


def test_extractmulti_colocalisation(ab):
    """``extractmulti_*`` trees (pipe_builder.py:19-43 without `costes`): Pearson, Manders, rank-weighted colocalisation
    and overlap of every channel pair per object against oracle/cpm.py (parity self-defined, like cp_measure's other
    features) — correlated and independent channels, absent ids, one-pixel and constant cells, a cell wider than the
    64 x 64 window, a Z stack under `max`, uint8 pixels, a non-default threshold — and the reference's column names."""
    from itertools import combinations

    from aliby_b200 import synth

    rng = np.random.default_rng(77)
    pixels, labels = synth.make_field(909, (320, 384), 3, 50, semi_axes=(3, 24))
    labels = labels.copy()
    labels[200:300, 20:140][labels[200:300, 20:140] == 0] = int(labels.max()) + 2  # wide cell, and one absent id below it
    pixels = pixels.copy()
    pixels[0, 1] = (pixels[0, 0] // 3 + rng.integers(0, 900, size=labels.shape)).astype(np.uint16)  # correlated with channel 0
    one = int(labels[labels > 0][0])
    pixels[0, 2, 0][labels == one] = 1234  # a constant cell in channel 2
    labels[5, 7] = int(labels.max()) + 1  # one-pixel cell
    tree = {pair: {"None": {"max": ["pearson", "manders_fold", "rwc", "overlap"]}} for pair in combinations(range(3), 2)}
    items, got = ab.process_tree_masks(tree, labels, pixels, ab.extract_tree_multi)
    assert items[0] == ((0, 1), ((0, 1), "None", "max", "pearson"))
    _check_colocalisation(items, got, _colocalisation_reference(tree, labels, pixels))
    table = ab.format_extraction((items, got))
    assert table.num_rows == int(labels.max())
    assert "(0, 1)/None/max/pearson/Correlation_Pearson" in table.column_names
    assert "(1, 2)/None/max/rwc/Correlation_RWC_2" in table.column_names
    plain = ab.format_extraction((tuple(items), list(got)))
    assert plain.column_names == table.column_names
    for c in table.column_names:
        x, y = table.column(c).to_pylist(), plain.column(c).to_pylist()
        assert all((p == q) or (p != p and q != q) for p, q in zip(x, y)), c
    # the dense fast entry names and fills the same columns
    fast_table = ab.extract_table(tree, labels, pixels)
    arrow = fast_table.to_arrow()
    assert arrow.column_names == table.column_names
    for c in table.column_names:
        x, y = table.column(c).to_pylist(), arrow.column(c).to_pylist()
        assert all((p == q) or (p != p and q != q) for p, q in zip(x, y)), c
    # repeated calls of one shape go through the captured graph from the third call on: same numbers
    for _ in range(3):
        again_items, again = ab.process_tree_masks(tree, labels, pixels, ab.extract_tree_multi)
        assert all(np.array_equal(a[k], b[k], equal_nan=True) for a, b in zip(again, got) for k in b)
    # a Z stack (max), several tiles, uint8, a different threshold, a subset of the features
    kw = {"manders_fold": {"thr": 40}}
    tiles = [synth.make_field(930 + t, (96, 128), 2, 6, n_z=3, semi_axes=(3, 12)) for t in range(3)]
    px8 = (np.concatenate([t[0] for t in tiles]) >> 8).astype(np.uint8)
    masks = [t[1] for t in tiles]
    tree2 = {(1, 0): {"None": {"max": ["manders_fold", "pearson"]}}}
    for px in (np.concatenate([t[0] for t in tiles]), px8):
        items2, got2 = ab.process_tree_masks(tree2, masks, px, ab.extract_tree_multi, cp_measure_kwargs=kw)
        _check_colocalisation(items2, got2, _colocalisation_reference(tree2, masks, px, kw))
    # what has no kernel says so: costes, a channel reduction, sums of a stack beyond 16 bits
    with pytest.raises(NotImplementedError, match="costes"):
        ab.process_tree_masks({(0, 1): {"None": {"max": ["costes"]}}}, labels, pixels, ab.extract_tree_multi)
    with pytest.raises(NotImplementedError, match="red_ch"):
        ab.process_tree_masks({(0, 1): {"add": {"max": ["pearson"]}}}, labels, pixels, ab.extract_tree_multi)
    with pytest.raises(KeyError, match="colocalise"):
        ab.process_tree_masks({(0, 1): {"None": {"max": ["colocalise"]}}}, labels, pixels, ab.extract_tree_multi)
    big = np.full((1, 2, 2, 64, 64), 40000, dtype=np.uint16)
    lab = np.zeros((64, 64), np.uint16)
    lab[10:20, 10:20] = 1
    with pytest.raises(NotImplementedError, match="65536"):
        ab.process_tree_masks({(0, 1): {"None": {"add": ["pearson"]}}}, lab, big, ab.extract_tree_multi)
    small = np.full((1, 2, 2, 64, 64), 300, dtype=np.uint16)
    small[0, 1, 1, 12:15] = 900
    small[0, 0, 0, 11:14, 11] = 70
    tree3 = {(0, 1): {"None": {"add": ["pearson", "rwc"]}}}
    items3, got3 = ab.process_tree_masks(tree3, lab, small, ab.extract_tree_multi)
    _check_colocalisation(items3, got3, _colocalisation_reference(tree3, lab, small))


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_cp_measure_and_pairs(ab, seed):
    """Randomised planes for the cp_measure features and the two-image features: odd plane sizes (rows that are not
    16-byte multiples: the CTA-per-object kernel serves every object), aligned ones (the sweep kernel), blocks, rings and
    ellipses up to 150 px, salt-and-pepper ids, id gaps, several tiles, uint8 / uint16, narrow and full value ranges,
    Z in {1, 2, 3} under `max`."""
    from aliby_b200 import synth

    rng = np.random.default_rng(9100 + seed)
    aligned = seed % 2 == 0
    H = int(rng.integers(64, 180))
    W = 16 * int(rng.integers(5, 14)) if aligned else int(rng.integers(65, 230))
    n_tiles = int(rng.integers(1, 4))
    Z = 1 if seed < 4 else int(rng.integers(2, 4))
    dtype = np.uint8 if seed % 4 == 1 else np.uint16
    top = np.iinfo(dtype).max
    masks = []
    for _ in range(n_tiles):
        lab = synth.ellipse_labels(rng, (H, W), int(rng.integers(0, 8)), semi_axes=(3, 28)).astype(np.uint16)
        next_id = int(lab.max()) + 1
        for _ in range(int(rng.integers(1, 6))):
            h, w = int(rng.integers(1, min(H, 150))), int(rng.integers(1, min(W, 150)))
            r, c = int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1))
            lab[r : r + h, c : c + w] = next_id
            if h > 8 and w > 8 and rng.random() < 0.4:
                lab[r + 3 : r + h - 3, c + 3 : c + w - 3] = 0
            next_id += int(rng.integers(1, 3))
        speck = rng.random((H, W)) < 0.01
        lab[speck] = rng.integers(1, next_id + 2, size=int(speck.sum()))
        masks.append(lab)
    if seed % 3 == 0:
        base = rng.integers(0, max(2, top // 2), size=(n_tiles, 3, 1, 1, 1))
        pixels = np.clip(base + rng.poisson(30, size=(n_tiles, 3, Z, H, W)), 0, top).astype(dtype)
    else:
        pixels = rng.integers(0, top + 1, size=(n_tiles, 3, Z, H, W)).astype(dtype)
    kw = {"intensity": {"edge_measurements": False}, "manders_fold": {"thr": int(rng.integers(5, 60))}}
    tree = {"None": {"None": ("sizeshape",)}, 2: {"max": ("intensity",)}, 0: {"max": ("intensity",)}}
    m = masks if n_tiles > 1 else masks[0]
    items, got = ab.process_tree_masks(tree, m, pixels, ab.extract_tree, cp_measure_kwargs=kw)
    _check_cp(items, got, _cpm_reference(tree, m, pixels, kw))
    multi = {(0, 1): {"None": {"max": ["pearson", "rwc"]}}, (2, 1): {"None": {"max": ["manders_fold", "overlap"]}}}
    items, got = ab.process_tree_masks(multi, m, pixels, ab.extract_tree_multi, cp_measure_kwargs=kw)
    _check_colocalisation(items, got, _colocalisation_reference(multi, m, pixels, kw))


def test_extractmulti_many_channels(ab):
    """Nine channels: more requests than the per-object kernel stages (8), so every pair takes the CTA-per-pair kernel;
    narrow values, so its rank tables are small — the route that no other test takes."""
    from aliby_b200 import synth

    rng = np.random.default_rng(4242)
    _, labels = synth.make_field(611, (160, 192), 1, 14, semi_axes=(3, 20))
    base = rng.integers(100, 3000, size=(1, 9, 1, 1, 1))
    pixels = (base + rng.poisson(25, size=(1, 9, 1, 160, 192))).astype(np.uint16)
    pixels[0, 3] = pixels[0, 2] + rng.integers(0, 3, size=(1, 160, 192)).astype(np.uint16)  # nearly identical channels
    tree = {(a, b): {"None": {"max": ["pearson", "rwc", "manders_fold"]}} for a, b in ((0, 8), (2, 3), (7, 1), (4, 5), (6, 0))}
    items, got = ab.process_tree_masks(tree, labels, pixels, ab.extract_tree_multi)
    _check_colocalisation(items, got, _colocalisation_reference(tree, labels, pixels))


def test_host_uploads_pinned_and_pageable(ab):
    """engine.host_to_device: page-locked memory is recognised through the driver (a NumPy view of a pinned tensor is
    pinned, a fresh array is not) and both routes — direct asynchronous copy, staging by threads — deliver the bytes."""
    import torch

    from aliby_b200 import _native as nat
    from aliby_b200 import engine

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    pageable = rng.integers(0, 65536, size=(3, 1500, 1501), dtype=np.uint16)  # 13.5 MB: several staging chunks
    pinned = torch.from_numpy(pageable).pin_memory().numpy()
    assert nat.lib().abx_host_is_pinned(pinned.ctypes.data) == 1 and nat.lib().abx_host_is_pinned(pinned[1:].ctypes.data) == 1
    assert nat.lib().abx_host_is_pinned(pageable.ctypes.data) == 0
    for src in (pageable, pinned, pageable[:, :64], pageable[1]):
        dst = torch.empty(src.shape, dtype=torch.uint16, device=dev)
        for _ in range(2):  # the second round reuses the staging buffer
            engine.host_to_device(dst, np.ascontiguousarray(src))
        assert np.array_equal(dst.cpu().numpy(), src)


def test_extractmulti_step_splits_with_the_reference(ab, monkeypatch):
    """``init_step('extractmulti_*')`` (pipe.py:65-66): the branches with a kernel run on the GPU, the others (costes)
    go to the reference's step, and the concatenated result pivots into one table with every requested column."""
    from aliby_b200 import pipe, synth

    pixels, labels = synth.make_field(941, (128, 160), 2, 9, semi_axes=(3, 12))
    seen = {}

    def fake_reference_init_step(step_name, parameters, other_steps, why):
        seen["tree"] = parameters["tree"]

        def step(masks, pixels, **kw):
            masks_ = masks if isinstance(masks, list) else [masks]
            insts = ab.kv(ab.flatten(parameters["tree"]))
            objs = [(t, k) for t, m in enumerate(masks_) for k in range(1, int(m.max()) + 1)]
            items = tuple((o, i) for o in objs for i in insts)
            return items, [{"Correlation_Costes_1": np.array([0.5]), "Correlation_Costes_2": np.array([0.25])} for _ in items]

        return step

    monkeypatch.setattr(pipe, "_reference_init_step", fake_reference_init_step)
    params = {"tree": {(0, 1): {"None": {"max": ["pearson", "costes", "manders_fold", "rwc"]}}}, "kwargs": {"ncores": None}}
    step = pipe.init_step("extractmulti_nuclei", params)
    assert seen["tree"] == {(0, 1): {"None": {"max": ["costes"]}}}
    items, results = step(masks=[labels], pixels=pixels)
    table = ab.format_extraction((items, results))
    assert table.num_rows == int(labels.max())
    cols = table.column_names
    assert "(0, 1)/None/max/costes/Correlation_Costes_2" in cols and "(0, 1)/None/max/rwc/Correlation_RWC_1" in cols
    want = _colocalisation_reference({(0, 1): {"None": {"max": ["pearson"]}}}, labels, pixels)
    got = table.column("(0, 1)/None/max/pearson/Correlation_Pearson").to_pylist()
    for g, w in zip(got, want):
        w = float(w["Correlation_Pearson"][0])
        assert (g != g and w != w) or abs(g - w) <= 1e-9
    # a tree the kernels cover completely never touches the reference
    monkeypatch.setattr(pipe, "_reference_init_step", lambda *a, **k: (_ for _ in ()).throw(AssertionError("not needed")))
    step2 = pipe.init_step("extractmulti_nuclei", {"tree": {(0, 1): {"None": {"max": ["pearson", "overlap"]}}}})
    items2, res2 = step2(masks=[labels], pixels=pixels)
    assert len(items2) == 2 * int(labels.max()) and list(res2[1]) == ["Correlation_Overlap", "Correlation_K_1", "Correlation_K_2"]


SHARD_WORKER = r'''
import sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from aliby_b200 import synth
from aliby_b200.sharding import extract_sharded
from aliby_b200.extract import extract_table

rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=rank, world_size=2)
torch.cuda.set_device(0)  # both ranks share the one GPU of the test box: the sharding logic is what is under test
tree = {"None": {"None": ["area", "eccentricity"]}, 0: {"max": ["mean", "median", "max5px_median"]}, 1: {"max": ["total", "std"]}}

def load(seed):
    px, lab = synth.make_field(seed, (192, 256), 2, 25, semi_axes=(4, 14))
    return lab, px

units = list(range(500, 509))
out = extract_sharded(tree, units, load)  # compute = the CUDA path
if rank == 0:
    assert [u for u, *_ in out] == units
    for u, objs, names, vals in out:
        t = extract_table(tree, *load(u))
        assert np.array_equal(objs, t.objects) and names == t.names
        assert np.array_equal(np.nan_to_num(vals, nan=-1), np.nan_to_num(t.values, nan=-1))
    print("SHARD_GPU_OK", len(out))
else:
    assert out is None
dist.destroy_process_group()
'''


def test_sharded_extraction_on_the_gpu(ab, tmp_path):
    """sharding.extract_sharded with its default compute (the CUDA path): one rank in-process, then two ranks (two
    processes over gloo, both on the test box's GPU) whose gathered tables equal the single-rank ones."""
    import os
    import subprocess
    import sys

    from aliby_b200 import sharding, synth
    from conftest import ROOT

    tree = {"None": {"None": ["area"]}, 0: {"max": ["mean", "median"]}}
    fields = {s: synth.make_field(s, (128, 160), 1, 10, semi_axes=(3, 10)) for s in (1, 2, 3)}
    out = sharding.extract_sharded(tree, [1, 2, 3], lambda s: (fields[s][1], fields[s][0]), rank=0, world=1)
    assert [u for u, *_ in out] == [1, 2, 3] and all(len(o) == int(fields[u][1].max()) for u, o, _, _ in out)
    script = tmp_path / "shard_worker.py"
    script.write_text(SHARD_WORKER)
    port = 29700 + (os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_GPU_OK 9" in outs[0]
