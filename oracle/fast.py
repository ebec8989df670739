"""Quick CPU oracle: same numbers as ``oracle.port`` in O(Y·X log) (TEST INFRASTRUCTURE).

Pixels are grouped by label with one stable argsort per label plane, so the
values of an object arrive in the same row-major order as the reference's
``img[mask]`` gather (``src/extraction/core/functions/cell.py``), and the very
same NumPy reductions are then applied to them: intensity metrics come out
bit-identical to ``oracle.port``.  Shape metrics run the reference's three
chained EDTs (``cell.py:207-229``) on the object's bounding box plus the
one-pixel frame, which is exact (nearest zero never lies outside that window);
the one case where SciPy's answer depends on the whole plane — an input without
any zero, for which ``distance_transform_edt`` measures the distance to index
(-1, 0) — is restated explicitly.
"""

from __future__ import annotations

import math

import numpy as np
from scipy import ndimage

from . import port


def _axes_from_window(win: np.ndarray, r0: int, c0: int):
    """``cell.py:207-229`` on a bbox crop; ``(r0, c0)`` = plane coords of win[0, 0]."""
    framed = np.pad(win, 1, mode="constant", constant_values=0)
    from_edge = ndimage.distance_transform_edt(framed == 1) * framed
    peak = from_edge.max()
    from_top = ndimage.distance_transform_edt(from_edge - peak) * framed
    flat = from_top == 0
    if flat.all():
        # no zero anywhere: SciPy measures to index (-1, 0) of the *padded plane*
        rr, cc = np.mgrid[0 : framed.shape[0], 0 : framed.shape[1]]
        plateau = np.sqrt((rr + r0 + 1.0) ** 2 + (cc + c0 + 0.0) ** 2) * framed
    else:
        plateau = ndimage.distance_transform_edt(flat) * framed
    minor = np.round(np.max(from_edge))
    major = np.round(np.max(from_top) + np.sum(plateau) / 2)
    return minor, major, from_edge


BBOX_METRICS = ("bbox_rmin", "bbox_rmax", "bbox_cmin", "bbox_cmax")


class PlaneIndex:
    """Sort-by-label index of one label plane."""

    def __init__(self, labels: np.ndarray):
        self.labels = labels
        self.n_labels = int(labels.max()) if labels.size else 0
        flat = labels.ravel()
        self.order = np.argsort(flat, kind="stable")
        counts = np.bincount(flat, minlength=self.n_labels + 1)
        self.counts = counts
        self.starts = np.concatenate([[0], np.cumsum(counts)])
        H, W = labels.shape
        self.rows = (self.order // W).astype(np.int64)
        self.cols = (self.order % W).astype(np.int64)

    def span(self, lab: int):
        return slice(self.starts[lab], self.starts[lab + 1])

    def bbox(self, lab: int):
        s = self.span(lab)
        r, c = self.rows[s], self.cols[s]
        return int(r.min()), int(r.max()), int(c.min()), int(c.max())


def shape_metric(index: PlaneIndex, lab: int, metric: str):
    n = int(index.counts[lab])
    s = index.span(lab)
    if metric == "area":
        return np.int64(n)
    if metric in ("centroid_x", "centroid_y", "centroid"):
        with np.errstate(invalid="ignore", divide="ignore"):
            x = np.int64(np.sum(index.cols[s] + 1)) / np.int64(n)
            y = np.int64(np.sum(index.rows[s] + 1)) / np.int64(n)
        return {"centroid_x": x, "centroid_y": y, "centroid": (x, y)}[metric]
    if metric == "spherical_volume":
        radius = math.sqrt(n / np.pi)
        return (4 * np.pi * radius**3) / 3
    if metric == "bbox":
        return index.bbox(lab) if n else None
    if metric in BBOX_METRICS:  # extension (AreaShape_BoundingBox* territory): inclusive bbox, NaN for absent ids
        return np.float64(index.bbox(lab)[BBOX_METRICS.index(metric)]) if n else np.float64(np.nan)
    # EDT family
    if n == 0:
        return {
            "eccentricity": np.float64(np.nan),
            "volume": np.float64(0.0),
            "conical_volume": np.float64(0.0),
            "min_maj_approximation": (np.float64(0.0), np.float64(0.0)),
        }[metric]
    r0, r1, c0, c1 = index.bbox(lab)
    win = index.labels[r0 : r1 + 1, c0 : c1 + 1] == lab
    minor, major, from_edge = _axes_from_window(win, r0, c0)
    if metric == "min_maj_approximation":
        return minor, major
    if metric == "eccentricity":
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.sqrt(major**2 - minor**2) / major
    if metric == "volume":
        return (4 * np.pi * minor**2 * major) / 3
    if metric == "conical_volume":
        return 4 * np.sum(from_edge)
    raise KeyError(metric)


def intensity_metric(index: PlaneIndex, lab: int, img: np.ndarray, metric: str):
    s = index.span(lab)
    v = img.ravel()[index.order[s]]
    n = len(v)
    with np.errstate(invalid="ignore", divide="ignore"):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if metric == "mean":
                return np.mean(v)
            if metric == "total":
                return np.sum(v)
            if metric == "total_squared":
                return np.sum(v**2)
            if metric == "median":
                return np.median(v)
            if metric == "std":
                return np.std(v)
            if metric == "max2p5pc":
                k = int(np.ceil(n * 0.025))
                return np.mean(np.partition(v, n - k)[-k:]) if n else np.float64(np.nan)
            if metric == "max5px_median":
                if n <= 5:
                    return np.nan
                top5 = np.partition(v, n - 5)[-5:]
                mid = np.median(v)
                return np.nan if mid == 0 else np.mean(top5) / np.median(v)
            if metric == "max":
                return v.max() if n else np.nan
            if metric == "min":
                return v.min() if n else np.nan
            if metric == "ratio":
                return np.nan
            if metric == "moment_of_inertia":
                if n == 0 or not np.any(v):
                    return np.nan
                r0, r1, c0, c1 = index.bbox(lab)
                crop = img[r0 : r1 + 1, c0 : c1 + 1].copy()
                crop[index.labels[r0 : r1 + 1, c0 : c1 + 1] != lab] = 0
                cols = np.arange(c0 + 1, c1 + 2)[None, :]
                rows = np.arange(r0 + 1, r1 + 2)[:, None]
                m00 = np.sum(crop)
                xm = np.sum(np.multiply(crop, cols)) / m00
                ym = np.sum(np.multiply(crop, rows)) / m00
                mu20 = np.sum(np.multiply(crop, (cols - xm) ** 2))
                mu02 = np.sum(np.multiply(crop, (rows - ym) ** 2))
                return mu20 / m00 ** (1 + (2 + 0) / 2) + mu02 / m00 ** (1 + (0 + 2) / 2)
    raise KeyError(metric)


SHAPE_METRICS = set(port.MASK_ONLY) | {"bbox"} | set(BBOX_METRICS)
BACKGROUND_METRICS = {"imBackground", "background_max5"}


def run_tree(tree: dict, masks, pixels: np.ndarray):
    """Same contract as :func:`oracle.port.run_tree` (object-major results)."""
    if not isinstance(masks, list):
        masks = [masks]
    instructions = port.tree_instructions(tree)
    objects = port.enumerate_objects(masks)
    items = tuple((o, i) for o in objects for i in instructions)
    results = [None] * len(items)
    n_inst = len(instructions)
    obj_pos = {o: k for k, o in enumerate(objects)}
    for tile_i, lab_plane in enumerate(masks):
        if not len(lab_plane) or lab_plane.max() == 0:
            continue
        index = PlaneIndex(np.asarray(lab_plane))
        projected: dict = {}
        for j, (ch, red, metric) in enumerate(instructions):
            if metric not in port.CELL_METRICS and metric not in {"max", "min"} | BACKGROUND_METRICS | set(BBOX_METRICS):
                raise KeyError(metric)
            img = None
            if ch != "None":
                key = (ch, red)
                if key not in projected:
                    projected[key] = port.project_z(pixels[tile_i, ch], port.Z_REDUCERS[red])
                img = projected[key]
            for lab in range(1, index.n_labels + 1):
                k = obj_pos[(tile_i, lab)] * n_inst + j
                if metric in SHAPE_METRICS:
                    results[k] = shape_metric(index, lab, metric)
                elif metric == "imBackground":
                    results[k] = np.median(img.ravel()[index.order[index.span(0)]])
                elif metric == "background_max5":
                    results[k] = np.mean(np.sort(img.ravel()[index.order[index.span(0)]])[-5:])
                else:
                    results[k] = intensity_metric(index, lab, img, metric)
    return items, results
