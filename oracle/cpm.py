"""CPU restatement of the cp_measure features the reference's stock pipelines request — TEST INFRASTRUCTURE.

**PARITY UNPINNED.**  The reference obtains ``intensity`` and ``sizeshape`` from the third-party package
``cp-measure 0.1.17`` (``uv.lock:441-442``; call sites ``src/extraction/core/functions/loaders.py:6,71-77,135-150``,
selected by ``src/aliby/pipe_builder.py:49-56,115-120``).  Its source is not under ``/root/reference`` and it is not
installed here; the reference's tests pin no value at that boundary.  What follows restates the PUBLISHED
CellProfiler definitions that cp_measure wraps — MeasureObjectIntensity and MeasureObjectSizeShape — for one binary
mask at a time, the way ``wrap_cp_measure_features`` calls them (``fun(mask.astype(uint16), pixels, **kw)`` ->
``{feature: ndarray of length 1}``).  Parity of the CUDA path with THIS file is self-defined; it is not parity with
cp_measure.  What can be checked here is checked: ``tests/test_oracle_golden.py::test_cpm_oracle_against_scipy_ndimage``
holds this file against the ``scipy.ndimage`` label functions CellProfiler itself calls (sum, mean, standard_deviation,
minimum, maximum, maximum_position, center_of_mass), against ``numpy.corrcoef`` and ``scipy.stats.rankdata(dense)``.

Conventions taken from CellProfiler:

* ``Intensity_*``: sums over the object's pixels of the image as given (no rescaling); ``StdIntensity`` is the
  population standard deviation; quartiles, median and MAD use CellProfiler's rule, NOT ``np.median``: with the
  object's ``n`` values sorted, ``q = n * fraction``, ``i = floor(q)``, ``f = q - i``; the result is
  ``v[i] * (1 - f) + v[i + 1] * f`` when ``i < n - 1`` and ``v[i]`` otherwise.  MAD applies the same rule (fraction
  0.5) to ``|v - median|``.
* ``Location_CenterMassIntensity_X/Y``: intensity-weighted mean of the 0-based column / row index;
  ``Location_MaxIntensity_X/Y``: position of the first maximum in row-major order; ``Intensity_MassDisplacement``:
  distance between the intensity-weighted and the unweighted centroid.  ``*_Z`` is 0 for a 2-D image.
* edge features (``*Edge``) need the object's outline and are not restated: the reference can switch them off with
  ``cp_measure_kwargs={"intensity": {"edge_measurements": False}}`` (``pipe_builder.py:84-91``), which is what the
  CUDA path requires.
* ``AreaShape_*`` (subset): ``Area``; ``BoundingBoxMinimum/Maximum_X/Y`` (maximum exclusive, as skimage's ``bbox``);
  ``BoundingBoxArea``; ``Center_X/Y`` (0-based centroid); ``EquivalentDiameter = sqrt(4 A / pi)``; ``Extent = A /
  bbox area``; ``MaximumRadius`` / ``MeanRadius`` = maximum / mean of the Euclidean distance of the object's pixels
  to the background; ``Eccentricity``, ``MajorAxisLength``, ``MinorAxisLength`` from the second
  central moments of the pixel coordinates as ``skimage.measure.regionprops`` defines them (inertia-tensor
  eigenvalues ``l1 >= l2``: axes ``4 sqrt(l)``, eccentricity ``sqrt(1 - l2 / l1)``).
  Orientation, Perimeter, FormFactor, Compactness, ConvexArea, Solidity, EulerNumber, Feret diameters, MedianRadius
  and the Zernike moments are not restated.
* ``Correlation_*`` (the two-image features of ``extractmulti_*`` steps, ``extract.py:200-237``,
  ``loaders.py:75-77,153-168``: ``fun(pixels1, pixels2, mask)``), after CellProfiler's MeasureColocalization for
  objects, one binary mask at a time: ``pearson`` = correlation of the two images over the object's pixels;
  ``manders_fold`` = with ``t1 = thr/100 * max(first)``, ``t2 = thr/100 * max(second)`` over the object (``thr`` = 15)
  and ``both = (first >= t1) & (second >= t2)``: ``M1 = sum(first[both]) / sum(first[first >= t1])``, M2 alike;
  ``rwc`` = the same sums weighted by ``(R - |rank1 - rank2|) / R`` with dense 0-based ranks of the object's values in
  each image and ``R = max(rank) + 1``; ``overlap`` = ``sum(f s) / sqrt(sum(f^2) sum(s^2))`` over ``both`` with
  ``K1 = sum(f s) / sum(f^2)``, ``K2 = sum(f s) / sum(s^2)``.  An object without a pixel in ``both`` has 0 for the
  thresholded features.  Integer images are taken as they are, in float64 (no wrap-around of squares).
  ``costes`` (an iterative search for a threshold whose scale depends on how cp_measure normalises integer images) is
  not restated.  The key names are CellProfiler's feature stems; cp_measure's own spelling could not be checked.
"""

from __future__ import annotations

import numpy as np
from scipy import ndimage

INTENSITY_FEATURES = (
    "Intensity_IntegratedIntensity",
    "Intensity_MeanIntensity",
    "Intensity_StdIntensity",
    "Intensity_MinIntensity",
    "Intensity_MaxIntensity",
    "Intensity_MassDisplacement",
    "Intensity_LowerQuartileIntensity",
    "Intensity_MedianIntensity",
    "Intensity_MADIntensity",
    "Intensity_UpperQuartileIntensity",
    "Location_CenterMassIntensity_X",
    "Location_CenterMassIntensity_Y",
    "Location_CenterMassIntensity_Z",
    "Location_MaxIntensity_X",
    "Location_MaxIntensity_Y",
    "Location_MaxIntensity_Z",
)

SIZESHAPE_FEATURES = (
    "AreaShape_Area",
    "AreaShape_BoundingBoxArea",
    "AreaShape_BoundingBoxMaximum_X",
    "AreaShape_BoundingBoxMaximum_Y",
    "AreaShape_BoundingBoxMinimum_X",
    "AreaShape_BoundingBoxMinimum_Y",
    "AreaShape_Center_X",
    "AreaShape_Center_Y",
    "AreaShape_Eccentricity",
    "AreaShape_EquivalentDiameter",
    "AreaShape_Extent",
    "AreaShape_MajorAxisLength",
    "AreaShape_MaximumRadius",
    "AreaShape_MeanRadius",
    "AreaShape_MinorAxisLength",
)


def cp_quantile(sorted_values: np.ndarray, fraction: float) -> float:
    """CellProfiler's order-statistic interpolation (MeasureObjectIntensity, quartile section)."""
    n = len(sorted_values)
    if n == 0:
        return float("nan")
    q = n * fraction
    i = int(np.floor(q))
    f = q - i
    if i < n - 1:
        return float(sorted_values[i]) * (1.0 - f) + float(sorted_values[i + 1]) * f
    return float(sorted_values[min(i, n - 1)])


def get_intensity(mask: np.ndarray, pixels: np.ndarray, edge_measurements: bool = False) -> dict:
    """``{feature: ndarray(1)}`` of one binary mask on one 2-D image."""
    if edge_measurements:
        raise NotImplementedError("edge intensities are not restated (parity unpinned): pass edge_measurements=False")
    m = np.asarray(mask) > 0
    img = np.asarray(pixels)
    rows, cols = np.nonzero(m)  # row-major order
    v = img[rows, cols].astype(np.float64)
    n = len(v)
    nan = float("nan")
    out = {k: nan for k in INTENSITY_FEATURES}
    out["Intensity_IntegratedIntensity"] = 0.0  # (an absent label: sums over no pixels are 0, everything else is NaN)
    if n:
        total = float(v.sum())
        mean = total / n
        out["Intensity_IntegratedIntensity"] = total
        out["Intensity_MeanIntensity"] = mean
        out["Intensity_StdIntensity"] = float(np.sqrt(np.mean((v - mean) ** 2)))
        out["Intensity_MinIntensity"] = float(v.min())
        out["Intensity_MaxIntensity"] = float(v.max())
        s = np.sort(v)
        med = cp_quantile(s, 0.5)
        out["Intensity_LowerQuartileIntensity"] = cp_quantile(s, 0.25)
        out["Intensity_MedianIntensity"] = med
        out["Intensity_UpperQuartileIntensity"] = cp_quantile(s, 0.75)
        out["Intensity_MADIntensity"] = cp_quantile(np.sort(np.abs(v - med)), 0.5)
        k = int(np.argmax(v))  # first maximum in row-major order
        out["Location_MaxIntensity_X"] = float(cols[k])
        out["Location_MaxIntensity_Y"] = float(rows[k])
        out["Location_MaxIntensity_Z"] = 0.0
        with np.errstate(invalid="ignore", divide="ignore"):
            cmx = float((v * cols).sum() / total) if total else nan
            cmy = float((v * rows).sum() / total) if total else nan
        out["Location_CenterMassIntensity_X"] = cmx
        out["Location_CenterMassIntensity_Y"] = cmy
        out["Location_CenterMassIntensity_Z"] = 0.0 if total else nan
        out["Intensity_MassDisplacement"] = float(np.hypot(cmx - cols.mean(), cmy - rows.mean()))
    return {k: np.array([val]) for k, val in out.items()}


def get_sizeshape(mask: np.ndarray, pixels=None) -> dict:
    """``{feature: ndarray(1)}`` (subset, see the module docstring) of one binary mask."""
    m = np.asarray(mask) > 0
    rows, cols = np.nonzero(m)
    n = len(rows)
    nan = float("nan")
    out = {k: nan for k in SIZESHAPE_FEATURES}
    out["AreaShape_Area"] = 0.0  # (an absent label)
    if n:
        r0, r1, c0, c1 = rows.min(), rows.max() + 1, cols.min(), cols.max() + 1
        out["AreaShape_Area"] = float(n)
        out["AreaShape_BoundingBoxMinimum_X"], out["AreaShape_BoundingBoxMaximum_X"] = float(c0), float(c1)
        out["AreaShape_BoundingBoxMinimum_Y"], out["AreaShape_BoundingBoxMaximum_Y"] = float(r0), float(r1)
        bbox_area = float((r1 - r0) * (c1 - c0))
        out["AreaShape_BoundingBoxArea"] = bbox_area
        out["AreaShape_Center_X"], out["AreaShape_Center_Y"] = float(cols.mean()), float(rows.mean())
        out["AreaShape_EquivalentDiameter"] = float(np.sqrt(4.0 * n / np.pi))
        out["AreaShape_Extent"] = n / bbox_area
        # distance of every object pixel to the background, the object padded by one pixel like a plane border is not
        win = np.pad(m[r0:r1, c0:c1], 1)
        dist = ndimage.distance_transform_edt(win)[1:-1, 1:-1][m[r0:r1, c0:c1]]
        out["AreaShape_MaximumRadius"] = float(dist.max())
        out["AreaShape_MeanRadius"] = float(dist.mean())
        # second central moments of the coordinates (skimage regionprops: inertia_tensor = [[mu02, -mu11], [-mu11, mu20]] / n
        # with axis 0 = rows; its eigenvalues are those of the covariance matrix)
        dr, dc = rows - rows.mean(), cols - cols.mean()
        mu_rr, mu_cc, mu_rc = float((dr * dr).sum() / n), float((dc * dc).sum() / n), float((dr * dc).sum() / n)
        half_tr, det_term = (mu_rr + mu_cc) / 2.0, np.sqrt(((mu_rr - mu_cc) / 2.0) ** 2 + mu_rc**2)
        l1, l2 = half_tr + det_term, max(half_tr - det_term, 0.0)
        out["AreaShape_MajorAxisLength"] = float(4.0 * np.sqrt(l1))
        out["AreaShape_MinorAxisLength"] = float(4.0 * np.sqrt(l2))
        out["AreaShape_Eccentricity"] = float(np.sqrt(1.0 - l2 / l1)) if l1 > 0 else 0.0
    return {k: np.array([val]) for k, val in out.items()}


CORRELATION_FEATURES = {
    "pearson": ("Correlation_Pearson",),
    "manders_fold": ("Correlation_Manders_1", "Correlation_Manders_2"),
    "rwc": ("Correlation_RWC_1", "Correlation_RWC_2"),
    "overlap": ("Correlation_Overlap", "Correlation_K_1", "Correlation_K_2"),
}


def _dense_rank(v: np.ndarray) -> np.ndarray:
    """0-based dense ranks (equal values share a rank), MeasureColocalization's lexsort / cumsum construction."""
    order = np.argsort(v, kind="stable")
    step = np.concatenate([[False], v[order[:-1]] != v[order[1:]]])
    rank = np.zeros(len(v), dtype=np.int64)
    rank[order] = np.cumsum(step)
    return rank


def get_correlation(pixels1: np.ndarray, pixels2: np.ndarray, mask: np.ndarray, thr: float = 15) -> dict:
    """Every restated two-image feature of one binary mask: ``{key: ndarray(1)}``."""
    m = np.asarray(mask) > 0
    fi = np.asarray(pixels1)[m].astype(np.float64)
    si = np.asarray(pixels2)[m].astype(np.float64)
    nan = float("nan")
    out = {k: nan for keys in CORRELATION_FEATURES.values() for k in keys}
    if len(fi):
        with np.errstate(invalid="ignore", divide="ignore"):
            x, y = fi - fi.mean(), si - si.mean()
            std1, std2 = np.sqrt((x * x).sum()), np.sqrt((y * y).sum())
            out["Correlation_Pearson"] = float((x * y / (std1 * std2)).sum())
            t1, t2 = (thr / 100) * fi.max(), (thr / 100) * si.max()
            both = (fi >= t1) & (si >= t2)
            tot1, tot2 = fi[fi >= t1].sum(), si[si >= t2].sum()
            thresholded = [k for name in ("manders_fold", "rwc", "overlap") for k in CORRELATION_FEATURES[name]]
            if both.any():
                f, s = fi[both], si[both]
                out["Correlation_Manders_1"] = float(f.sum() / tot1)
                out["Correlation_Manders_2"] = float(s.sum() / tot2)
                r1, r2 = _dense_rank(fi), _dense_rank(si)
                big_r = max(r1.max(), r2.max()) + 1
                weight = ((big_r - np.abs(r1 - r2)) / big_r)[both]
                out["Correlation_RWC_1"] = float((f * weight).sum() / tot1)
                out["Correlation_RWC_2"] = float((s * weight).sum() / tot2)
                fs, ff, ss = (f * s).sum(), (f * f).sum(), (s * s).sum()
                out["Correlation_Overlap"] = float(fs / np.sqrt(ff * ss))
                out["Correlation_K_1"] = float(fs / ff)
                out["Correlation_K_2"] = float(fs / ss)
            else:
                for k in thresholded:
                    out[k] = 0.0
    return {k: np.array([v]) for k, v in out.items()}


def _correlation_subset(name):
    def fun(pixels1, pixels2, mask, **kw):
        full = get_correlation(pixels1, pixels2, mask, **kw)
        return {k: full[k] for k in CORRELATION_FEATURES[name]}

    fun.__name__ = f"get_correlation_{name}"
    return fun


FEATURES = {"intensity": get_intensity, "sizeshape": get_sizeshape}
CORRELATIONS = {name: _correlation_subset(name) for name in CORRELATION_FEATURES}
