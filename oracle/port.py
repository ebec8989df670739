"""Faithful CPU restatement of the reference extraction path (TEST INFRASTRUCTURE).

Every function names the reference lines it restates (paths relative to
``/root/reference/``).  The cost model is deliberately the reference's own:
one-hot label planes plus a full-plane boolean gather per (object, instruction).
It is the slow oracle; ``oracle.fast`` is the quick one.
"""

from __future__ import annotations

import itertools
import math

import numpy as np
from scipy import ndimage

# --------------------------------------------------------------------------
# Z reduction  — src/extraction/core/functions/distributors.py:6-24,
#                src/extraction/core/functions/loaders.py:110-127
# --------------------------------------------------------------------------
Z_REDUCERS = {
    "max": np.maximum,
    "mean": np.mean,
    "median": np.median,
    "div": np.divide,
    "add": np.add,
    "None": None,
}


def project_z(stack: np.ndarray, reducer, axis: int = 0) -> np.ndarray:
    """Only ufuncs are legal reducers; anything else raises (distributors.py:19-24)."""
    if isinstance(reducer, np.ufunc):
        return reducer.reduce(stack, axis=axis)
    raise Exception(f"{reducer} is an invalid reducer.")


# --------------------------------------------------------------------------
# label plane -> one-hot planes — src/agora/utils/masks.py:35-37
# --------------------------------------------------------------------------
def one_hot_planes(labels: np.ndarray) -> np.ndarray:
    ids = np.arange(1, labels.max() + 1)
    return np.equal.outer(ids, labels)


# --------------------------------------------------------------------------
# per-object metrics — src/extraction/core/functions/cell.py
# --------------------------------------------------------------------------
def m_area(mask):  # cell.py:18-27
    return np.sum(mask)


def m_mean(mask, img):  # cell.py:43-53
    return np.mean(img[mask])


def m_total(mask, img):  # cell.py:56-66
    return np.sum(img[mask])


def m_total_squared(mask, img):  # cell.py:69-83 (square taken in the image dtype)
    return np.sum(img[mask] ** 2)


def m_median(mask, img):  # cell.py:86-96
    return np.median(img[mask])


def m_max2p5pc(mask, img):  # cell.py:99-116
    count = np.sum(mask)
    k = int(np.ceil(count * 0.025))
    vals = img[mask]
    brightest = np.partition(vals, len(vals) - k)[-k:]
    return np.mean(brightest)


def m_max5px_median(mask, img):  # cell.py:119-144
    vals = img[mask]
    if len(vals) <= 5:
        return np.nan
    top5 = np.partition(vals, len(vals) - 5)[-5:]
    mid = np.median(vals)
    if mid == 0:
        return np.nan
    return np.mean(top5) / np.median(vals)


def m_std(mask, img):  # cell.py:147-157
    return np.std(img[mask])


def axes_estimate(mask):  # cell.py:207-229
    framed = np.pad(mask, 1, mode="constant", constant_values=0)
    from_edge = ndimage.distance_transform_edt(framed == 1) * framed
    from_top = ndimage.distance_transform_edt(from_edge - from_edge.max()) * framed
    plateau = ndimage.distance_transform_edt(from_top == 0) * framed
    minor = np.round(np.max(from_edge))
    major = np.round(np.max(from_top) + np.sum(plateau) / 2)
    return minor, major


def m_eccentricity(mask):  # cell.py:30-40
    minor, major = axes_estimate(mask)
    return np.sqrt(major**2 - minor**2) / major


def m_volume(mask):  # cell.py:160-173
    minor, major = axes_estimate(mask)
    return (4 * np.pi * minor**2 * major) / 3


def m_conical_volume(mask):  # cell.py:176-187
    framed = np.pad(mask, 1, mode="constant", constant_values=0)
    from_edge = ndimage.distance_transform_edt(framed == 1) * framed
    return 4 * np.sum(from_edge)


def m_spherical_volume(mask):  # cell.py:190-204
    radius = math.sqrt(m_area(mask) / np.pi)
    return (4 * np.pi * radius**3) / 3


def m_moment_of_inertia(mask, img):  # cell.py:232-265 (mutates img like the reference)
    img[~mask] = 0
    if not np.any(img):
        return np.nan
    cols = np.arange(1, img.shape[1] + 1, 1)[:, None].T
    rows = np.arange(1, img.shape[0] + 1, 1)[:, None]
    m00 = np.sum(img)
    m10 = np.sum(np.multiply(img, cols))
    m01 = np.sum(np.multiply(img, rows))
    xm = m10 / m00
    ym = m01 / m00
    mu20 = np.sum(np.multiply(img, (cols - xm) ** 2))
    mu02 = np.sum(np.multiply(img, (rows - ym) ** 2))
    eta20 = mu20 / m00 ** (1 + (2 + 0) / 2)
    eta02 = mu02 / m00 ** (1 + (0 + 2) / 2)
    return eta20 + eta02


def m_ratio(mask, img):  # cell.py:268-279
    if img.ndim == 3 and img.shape[-1] == 2:
        a = img[..., 0][mask]
        b = img[..., 1][mask]
        return np.nan if np.any(b == 0) else np.median(a / b)
    return np.nan


def m_centroid(mask):  # cell.py:282-293 (1-based coordinates)
    wc = np.arange(1, mask.shape[1] + 1, 1).reshape(1, mask.shape[1])
    wr = np.arange(1, mask.shape[0] + 1, 1).reshape(mask.shape[0], 1)
    m00 = np.sum(mask)
    m10 = np.sum(np.multiply(mask, wc))
    m01 = np.sum(np.multiply(mask, wr))
    return (m10 / m00, m01 / m00)


def m_centroid_x(mask):  # cell.py:296-298
    return m_centroid(mask)[0]


def m_centroid_y(mask):  # cell.py:301-303
    return m_centroid(mask)[1]


# per-tile background metrics — src/extraction/core/functions/trap.py:6-43.
# The reference expects (Y, X, N) masks; restated here on a label plane because
# nothing in the reference dispatches them.
def t_background_median(labels, img):
    return np.median(img[labels == 0])


def t_background_max5(labels, img):
    return np.mean(np.sort(img[labels == 0])[-5:])


MASK_ONLY = {
    "area": m_area,
    "centroid": m_centroid,
    "centroid_x": m_centroid_x,
    "centroid_y": m_centroid_y,
    "conical_volume": m_conical_volume,
    "eccentricity": m_eccentricity,
    "min_maj_approximation": axes_estimate,
    "spherical_volume": m_spherical_volume,
    "volume": m_volume,
}
MASK_AND_IMAGE = {
    "max2p5pc": m_max2p5pc,
    "max5px_median": m_max5px_median,
    "mean": m_mean,
    "median": m_median,
    "moment_of_inertia": m_moment_of_inertia,
    "ratio": m_ratio,
    "std": m_std,
    "total": m_total,
    "total_squared": m_total_squared,
}


def cell_metric_table():
    """Registry with the uniform ``(mask, pixels)`` call — loaders.py:28-79,170-171."""
    table = {name: (lambda m, _p, _f=f: _f(m)) for name, f in MASK_ONLY.items()}
    table.update(MASK_AND_IMAGE)
    return table


CELL_METRICS = cell_metric_table()


# --------------------------------------------------------------------------
# tree handling and the object × instruction loop — src/extraction/extract.py
# --------------------------------------------------------------------------
def tree_instructions(tree: dict) -> list[tuple]:
    """Nested ``{ch: {red: [metrics]}}`` -> ``[(ch, red, metric)]`` in insertion order
    (extract.py:33-74)."""
    out: list[tuple] = []

    def walk(node, prefix):
        for key, val in node.items():
            if isinstance(val, dict):
                walk(val, (*prefix, key))
            else:
                out.extend((*prefix, key, leaf) for leaf in val)

    walk(tree, ())
    return out


def enumerate_objects(masks: list) -> list[tuple[int, int]]:
    """Every id 1..max per tile, absent ids included (extract.py:276-281)."""
    objs = []
    for tile_i, tile_labels in enumerate(masks):
        if len(tile_labels):
            objs.extend((tile_i, lab) for lab in range(1, int(tile_labels.max()) + 1))
    return objs


def measure_one(one_hot, pixels, item, metrics=CELL_METRICS):
    """extract.py:77-153 — Z projection is redone for every call, like the reference."""
    (tile_i, lab), (ch, red, metric) = item
    plane = one_hot[tile_i][lab - 1]
    img = None
    if ch != "None":
        img = project_z(pixels[tile_i, ch], Z_REDUCERS[red])
    return metrics[metric](plane, img)


def run_tree(tree: dict, masks, pixels: np.ndarray, ncores=None):
    """``process_tree_masks`` + ``extract_tree`` (extract.py:240-375).

    Returns ``(items, results)`` with items = product(objects, instructions),
    object-major.  ``ncores`` selects the reference's joblib fan-out.
    """
    if not isinstance(masks, list):
        masks = [masks]
    instructions = tree_instructions(tree)
    items = tuple(itertools.product(enumerate_objects(masks), instructions))
    results = []
    if len(items):
        one_hot = [one_hot_planes(m) if len(m) else None for m in masks]
        if ncores is None:
            results = [measure_one(one_hot, pixels, it) for it in items]
        else:
            from joblib import Parallel, delayed

            results = list(
                Parallel(n_jobs=min(len(items), ncores))(
                    delayed(measure_one)(one_hot, pixels, it) for it in items
                )
            )
    return items, results


def run_tree_sample(tree: dict, masks, pixels: np.ndarray, objects: list[tuple[int, int]]):
    """Same per-item work as :func:`run_tree` restricted to ``objects`` (bench sampling).

    Builds one-hot planes only for the sampled ids so that a 2160² field does not
    need the reference's 9 GB ``(L, Y, X)`` array; the per-item cost (full-plane
    gather + metric) is unchanged.
    """
    if not isinstance(masks, list):
        masks = [masks]
    instructions = tree_instructions(tree)
    items = tuple(itertools.product(objects, instructions))
    results = []
    for (tile_i, lab), (ch, red, metric) in items:
        plane = masks[tile_i] == lab
        img = None
        if ch != "None":
            img = project_z(pixels[tile_i, ch], Z_REDUCERS[red])
        results.append(CELL_METRICS[metric](plane, img))
    return items, results


# --------------------------------------------------------------------------
# long -> wide pivot — src/extraction/extract.py:520-599
# --------------------------------------------------------------------------
def pivot_wide(items, results) -> dict[str, list]:
    """Column dict of the reference's Arrow table (names, order, null filling)."""
    rows: dict[tuple, dict] = {}
    names = set()
    for (obj, inst), val in zip(items, results, strict=True):
        if not isinstance(val, (int, float)):
            raise Exception(f"the metrics are in an invalid value: {type(val)}.")
        name = "/".join(str(x) for x in inst) + f"/{inst[-1]}"
        names.add(name)
        rows.setdefault((obj[0], obj[-1]), {})[name] = val
    ordered = sorted(names)
    wide = {"tile": [], "label": []}
    wide.update({n: [] for n in ordered})
    for (tile_i, lab), cells in rows.items():
        wide["tile"].append(tile_i)
        wide["label"].append(lab)
        for n in ordered:
            wide[n].append(cells.get(n))
    return wide


# --------------------------------------------------------------------------
# tile crop — src/aliby/tile/tiles.py:109-166, src/aliby/tile/tiler.py:601-650
# --------------------------------------------------------------------------
def tile_window(centre, size, drifts, tp):
    """(first-axis slice, second-axis slice) of a tile at time ``tp``."""
    c = (np.asarray(centre) - np.sum(np.asarray(drifts, dtype=float).reshape(-1, 2)[: tp + 1], axis=0)).astype(int)
    a0 = int(c[0] - size[0] // 2)
    a1 = int(c[1] - size[1] // 2)
    return slice(a0, a0 + size[0]), slice(a1, a1 + size[1])


def crop_with_padding(stack: np.ndarray, window) -> np.ndarray:
    """``(Z, Y, X)`` -> ``(Z, h, w)`` with the reference's out-of-bounds rules."""
    bounds = stack.shape[-2:]
    inner = [slice(max(0, s.start), min(ub, s.stop)) for s, ub in zip(window, bounds)]
    pad = np.array([(-min(0, s.start), -min(0, ub - s.stop)) for s, ub in zip(window, bounds)])
    tile = stack[:, inner[0], inner[1]]
    if pad.any():
        want = [s.stop - s.start for s in window]
        if (pad / 0.25 > want).any():
            tile = np.full((stack.shape[0], *want), np.nan)
        else:
            tile = np.pad(tile, [[0, 0]] + pad.tolist(), "median")
    return tile


def crop_tiles(frame: np.ndarray, centres, size, drifts=(), tp: int = 0) -> np.ndarray:
    """``(C, Z, Y, X)`` -> ``(tiles, C, Z, h, w)`` (tiler.py:309-366)."""
    per_channel = []
    for ch in range(frame.shape[0]):
        per_channel.append(
            np.stack([crop_with_padding(frame[ch], tile_window(c, size, drifts, tp)) for c in centres])
        )
    return np.swapaxes(np.array(per_channel), 0, 1)
