"""The CPU oracles against vectors produced by the real reference (oracle/make_golden.py)."""

import numpy as np
import pytest

from conftest import as_float_pairs, assert_same, golden_tree, load_golden
from oracle import fast, port


@pytest.mark.parametrize("impl", [port, fast], ids=["port", "fast"])
def test_field_small(impl):
    g = load_golden("field_small.npz")
    items, res = impl.run_tree(golden_tree(g), g["labels"], g["pixels"])
    assert len(items) == int(g["n_items"])
    a, b = as_float_pairs(res)
    # fast oracle sums moment_of_inertia / conical_volume over a bbox window instead of the
    # whole plane (different pairwise-summation tree) -> 1e-12; everything else identical
    rtol = 0.0 if impl is port else 1e-12
    assert_same(a, g["values"], rtol, "values")
    assert_same(b, g["values2"], rtol, "values2")


def test_fast_equals_port_bitwise_on_intensity():
    g = load_golden("field_small.npz")
    tree = {0: {"max": ["mean", "std", "median", "total", "total_squared", "max2p5pc", "max5px_median"]},
            1: {"add": ["mean", "std", "median", "total", "total_squared", "max2p5pc", "max5px_median"]}}
    _, r_port = port.run_tree(tree, g["labels"], g["pixels"])
    _, r_fast = fast.run_tree(tree, g["labels"], g["pixels"])
    assert_same(as_float_pairs(r_fast)[0], as_float_pairs(r_port)[0], 0.0, "intensity")


@pytest.mark.parametrize("impl", [port, fast], ids=["port", "fast"])
def test_tiles_list_and_table(impl):
    g = load_golden("tiles_list.npz")
    masks = [m for m in g["labels"]]
    items, res = impl.run_tree(golden_tree(g), masks, g["pixels"])
    assert [it[0][0] for it in items] == g["item_tile"].tolist()
    assert [it[0][1] for it in items] == g["item_label"].tolist()
    assert_same(as_float_pairs(res)[0], g["values"], 0.0 if impl is port else 1e-12, "values")
    import json

    wide = port.pivot_wide(items, [float(r) for r in res])
    assert list(wide) == json.loads(str(g["table_columns"]))
    assert wide["tile"] == g["table_tile"].tolist()
    assert wide["label"] == g["table_label"].tolist()
    got = np.stack([np.asarray(wide[c], dtype=float) for c in list(wide)[2:]], axis=1)
    assert_same(got, g["table_values"], 0.0 if impl is port else 1e-12, "table")


def test_volume_shapes():
    """tests/extraction/test_volume.py:32-74 — reference outputs on numpy-drawn disks/ellipses."""
    from oracle.make_golden import numpy_disk, numpy_ellipse

    g = load_golden("volume_shapes.npz")
    for kind, x, ecc, rot, want in zip(g["kind"], g["x"], g["ecc"], g["rot"], g["out"]):
        if kind == "disk":
            m = numpy_disk(int(x))
        else:
            y = int(np.round(np.sqrt(x**2 / (1 - ecc**2))))
            m = numpy_ellipse(int(x), y, int(rot))
        mn, mj = port.axes_estimate(m)
        got = [mn, mj, port.m_volume(m), port.m_eccentricity(m), port.m_conical_volume(m)]
        assert_same(got, want, 0.0, f"{kind} {x} {ecc} {rot}")
        idx = fast.PlaneIndex(m.astype(np.uint16))
        got_f = [*fast.shape_metric(idx, 1, "min_maj_approximation"), fast.shape_metric(idx, 1, "volume"),
                 fast.shape_metric(idx, 1, "eccentricity"), fast.shape_metric(idx, 1, "conical_volume")]
        assert_same(got_f, want, 1e-12, f"fast {kind} {x} {ecc} {rot}")
        if kind == "disk":
            # the reference's own analytic bound (1 %)
            real_v = 4 * np.pi * x**3 / 3
            assert abs(got[2] - real_v) / real_v < 0.01


@pytest.mark.parametrize("impl", ["port", "fast"])
def test_degenerate_shapes(impl):
    g = load_golden("degenerate_shapes.npz")
    lab = g["labels"]
    idx = fast.PlaneIndex(lab)
    for k in range(1, 10):
        if impl == "port":
            m = lab == k
            mn, mj = port.axes_estimate(m)
            got = [mn, mj, port.m_volume(m), port.m_eccentricity(m), port.m_conical_volume(m), port.m_area(m)]
        else:
            got = [*fast.shape_metric(idx, k, "min_maj_approximation"), fast.shape_metric(idx, k, "volume"),
                   fast.shape_metric(idx, k, "eccentricity"), fast.shape_metric(idx, k, "conical_volume"),
                   fast.shape_metric(idx, k, "area")]
        assert_same(got, g["out"][k - 1], 0.0 if impl == "port" else 1e-12, f"label {k}")


def test_background():
    g = load_golden("background.npz")
    assert port.t_background_median(g["labels"], g["image"]) == float(g["imBackground"])
    assert port.t_background_max5(g["labels"], g["image"]) == float(g["background_max5"])


def test_tile_crop():
    g = load_golden("tile_crop.npz")
    frame, centres, drifts = g["frame"], g["centres"], g["drifts"]
    for tp in (0, 1):
        for i, c in enumerate(centres):
            want = g[f"tp{tp}_tile{i}"]
            got = port.crop_with_padding(frame, port.tile_window(c, (16, 16), drifts, tp))
            assert got.dtype == want.dtype
            assert_same(got, want, 0.0, f"tp{tp} tile{i}")


def test_invalid_reducer_raises():
    with pytest.raises(Exception, match="invalid reducer"):
        port.project_z(np.zeros((2, 3, 3)), port.Z_REDUCERS["mean"])
    with pytest.raises(Exception, match="invalid reducer"):
        port.project_z(np.zeros((2, 3, 3)), port.Z_REDUCERS["None"])


def test_fast_extension_and_background_metrics():
    """max/min/imBackground/background_max5 of the fast oracle against one-line NumPy definitions
    (self-defined parity: the reference has no dispatched equivalent)."""
    from aliby_b200 import synth

    pixels, labels = synth.make_field(77, (80, 90), 2, 8, n_z=2, semi_axes=(4, 10))
    tree = {0: {"max": ["max", "min", "imBackground", "background_max5"]}, 1: {"add": ["max", "imBackground"]}}
    items, res = fast.run_tree(tree, labels, pixels)
    img0 = pixels[0, 0].max(axis=0)
    img1 = pixels[0, 1].sum(axis=0, dtype=np.uint64)
    for (obj, inst), r in zip(items, res):
        m = labels == obj[1]
        img = img0 if inst[0] == 0 else img1
        if inst[2] == "max":
            want = img[m].max() if m.any() else np.nan
        elif inst[2] == "min":
            want = img[m].min() if m.any() else np.nan
        elif inst[2] == "imBackground":
            want = np.median(img[labels == 0])
        else:
            want = np.mean(np.sort(img[labels == 0])[-5:])
        assert (np.isnan(want) and np.isnan(r)) or float(r) == float(want)


def test_fast_equals_port_on_c1_sample():
    """The quick oracle against the faithful port on a sample of a C1-sized field."""
    from aliby_b200 import synth

    pixels, labels = synth.make_field(1001, (270, 300), 2, 30)
    tree = {"None": {"None": ["area", "centroid_x", "centroid_y", "eccentricity", "volume", "conical_volume"]},
            0: {"max": ["mean", "std", "median", "total", "max2p5pc", "max5px_median", "moment_of_inertia"]}}
    _, a = port.run_tree(tree, labels, pixels)
    _, b = fast.run_tree(tree, labels, pixels)
    assert_same(as_float_pairs(b)[0], as_float_pairs(a)[0], 1e-12, "fast vs port")


def test_cpm_oracle_against_scipy_ndimage():
    """oracle/cpm.py is unpinned against cp_measure itself (its source is absent), but CellProfiler's
    MeasureObjectIntensity / MeasureColocalization obtain their per-object sums, extrema, positions and deviations from
    ``scipy.ndimage`` label functions — which ARE here: the restatement must agree with them on a labelled image, and
    its rank rule with a direct sort."""
    from scipy import ndimage

    from oracle import cpm

    rng = np.random.default_rng(11)
    labels = np.zeros((60, 70), np.int32)
    labels[5:25, 8:30] = 1
    labels[30:55, 10:22] = 2
    labels[28:40, 40:66] = 3
    labels[12, 50] = 4
    img = rng.integers(0, 4000, size=labels.shape).astype(np.uint16)
    img2 = (img // 2 + rng.integers(0, 500, size=labels.shape)).astype(np.uint16)
    idx = np.arange(1, 5)
    f = img.astype(np.float64)
    total, mean = ndimage.sum(f, labels, idx), ndimage.mean(f, labels, idx)
    std, lo, hi = ndimage.standard_deviation(f, labels, idx), ndimage.minimum(f, labels, idx), ndimage.maximum(f, labels, idx)
    maxpos = ndimage.maximum_position(f, labels, idx)
    com = ndimage.center_of_mass(f, labels, idx)
    geo = ndimage.center_of_mass(np.ones_like(f), labels, idx)
    med = ndimage.median(f, labels, idx)
    for k in idx:
        got = {key: float(v[0]) for key, v in cpm.get_intensity(labels == k, img).items()}
        assert got["Intensity_IntegratedIntensity"] == total[k - 1] and got["Intensity_MinIntensity"] == lo[k - 1]
        assert got["Intensity_MaxIntensity"] == hi[k - 1]
        assert abs(got["Intensity_MeanIntensity"] - mean[k - 1]) < 1e-9 and abs(got["Intensity_StdIntensity"] - std[k - 1]) < 1e-9
        assert (got["Location_MaxIntensity_Y"], got["Location_MaxIntensity_X"]) == tuple(float(x) for x in maxpos[k - 1])
        assert abs(got["Location_CenterMassIntensity_Y"] - com[k - 1][0]) < 1e-9
        assert abs(got["Location_CenterMassIntensity_X"] - com[k - 1][1]) < 1e-9
        want_disp = np.hypot(com[k - 1][0] - geo[k - 1][0], com[k - 1][1] - geo[k - 1][1])
        assert abs(got["Intensity_MassDisplacement"] - want_disp) < 1e-9
        # CellProfiler's rank rule against a direct sort; for an odd count it is the ordinary median shifted by half a step
        v = np.sort(f[labels == k])
        n = len(v)
        for frac, key in ((0.25, "Intensity_LowerQuartileIntensity"), (0.5, "Intensity_MedianIntensity"),
                          (0.75, "Intensity_UpperQuartileIntensity")):
            q = n * frac
            i = int(q)
            want = v[i] * (1 - (q - i)) + v[i + 1] * (q - i) if i < n - 1 else v[min(i, n - 1)]
            assert got[key] == want
        if n % 2 == 0:
            assert got["Intensity_MedianIntensity"] == v[n // 2] >= med[k - 1]  # upper of the two middle values
        # two-image features: Pearson against numpy's own correlation, Manders against explicit sums
        co = {key: float(x[0]) for key, x in cpm.get_correlation(img, img2, labels == k).items()}
        a, b = f[labels == k], img2[labels == k].astype(np.float64)
        if n > 1:
            assert abs(co["Correlation_Pearson"] - np.corrcoef(a, b)[0, 1]) < 1e-9
            both = (a >= 0.15 * a.max()) & (b >= 0.15 * b.max())
            assert abs(co["Correlation_Manders_1"] - a[both].sum() / a[a >= 0.15 * a.max()].sum()) < 1e-12
            from scipy.stats import rankdata

            ra, rb = rankdata(a, method="dense") - 1, rankdata(b, method="dense") - 1
            big_r = max(ra.max(), rb.max()) + 1
            w = (big_r - np.abs(ra - rb)) / big_r
            assert abs(co["Correlation_RWC_1"] - (a[both] * w[both]).sum() / a[a >= 0.15 * a.max()].sum()) < 1e-12
        # size / shape: bounding box from ndimage.find_objects, centre from center_of_mass, axes from the eigenvalues of
        # the coordinate covariance (numpy.linalg), radii from the distance transform of the padded mask
        ss = {key: float(x[0]) for key, x in cpm.get_sizeshape(labels == k).items()}
        sl = ndimage.find_objects((labels == k).astype(np.int32))[0]
        assert (ss["AreaShape_BoundingBoxMinimum_Y"], ss["AreaShape_BoundingBoxMaximum_Y"]) == (sl[0].start, sl[0].stop)
        assert (ss["AreaShape_BoundingBoxMinimum_X"], ss["AreaShape_BoundingBoxMaximum_X"]) == (sl[1].start, sl[1].stop)
        assert ss["AreaShape_Area"] == n and abs(ss["AreaShape_Center_Y"] - geo[k - 1][0]) < 1e-9
        rr, cc = np.nonzero(labels == k)
        if n > 1:
            lam = np.linalg.eigvalsh(np.cov(np.stack([rr, cc]).astype(np.float64), bias=True))
            assert abs(ss["AreaShape_MajorAxisLength"] - 4 * np.sqrt(lam[1])) < 1e-9
            assert abs(ss["AreaShape_MinorAxisLength"] - 4 * np.sqrt(max(lam[0], 0))) < 1e-9
            assert abs(ss["AreaShape_Eccentricity"] - np.sqrt(1 - max(lam[0], 0) / lam[1])) < 1e-9
        dist = ndimage.distance_transform_edt(np.pad(labels == k, 1))[1:-1, 1:-1][labels == k]
        assert abs(ss["AreaShape_MaximumRadius"] - dist.max()) < 1e-12 and abs(ss["AreaShape_MeanRadius"] - dist.mean()) < 1e-12
