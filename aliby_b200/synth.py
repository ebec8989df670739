"""Deterministic synthetic fields for the extraction hot path.

Follows the recipe in SURVEY.md §8(d): uint16 pixels with Poisson background and
per-object Poisson signal, uint16 label planes of non-overlapping rotated
ellipses with a few deleted ids (absent-label NaN rows), a handful of 1-6 pixel
objects (NaN rule of ``max5px_median``), one saturated and one constant object
and some objects touching the border.

Pure NumPy on purpose: the same arrays feed the CUDA path, the oracle and the
bench, on this CPU container and on the GPU box.
"""

from __future__ import annotations

import numpy as np

CONFIG_SEEDS = {"C1": 1001, "C2": 1002, "C3": 1003, "C4": 1004, "C5": 1005}


def ellipse_labels(
    rng: np.random.Generator,
    shape: tuple[int, int],
    n_objects: int,
    semi_axes: tuple[float, float] = (8.0, 30.0),
    n_tiny: int = 6,
    delete_frac: float = 0.01,
    max_tries: int = 200,
) -> np.ndarray:
    """Label plane (uint16) of non-overlapping filled, rotated ellipses.

    Ids run 1..N in placement order; ``delete_frac`` of them are erased again so
    that the id range has holes; ``n_tiny`` objects of 1-6 pixels are added.
    """
    H, W = shape
    lab = np.zeros((H, W), dtype=np.uint16)
    next_id = 1
    fails = 0
    while next_id <= n_objects and fails < max_tries * n_objects:
        a = rng.uniform(*semi_axes)
        b = rng.uniform(*semi_axes)
        th = rng.uniform(0.0, np.pi)
        # some objects are allowed to stick out of the field (border touching)
        cy = rng.uniform(-0.3 * a, H - 1 + 0.3 * a) if rng.random() < 0.05 else rng.uniform(a, H - 1 - a)
        cx = rng.uniform(-0.3 * a, W - 1 + 0.3 * a) if rng.random() < 0.05 else rng.uniform(a, W - 1 - a)
        R = int(np.ceil(max(a, b))) + 1
        r0, r1 = max(0, int(cy) - R), min(H, int(cy) + R + 1)
        c0, c1 = max(0, int(cx) - R), min(W, int(cx) + R + 1)
        if r1 <= r0 or c1 <= c0:
            fails += 1
            continue
        yy, xx = np.mgrid[r0:r1, c0:c1]
        dy, dx = yy - cy, xx - cx
        u = (dx * np.cos(th) + dy * np.sin(th)) / a
        v = (-dx * np.sin(th) + dy * np.cos(th)) / b
        inside = (u * u + v * v) <= 1.0
        win = lab[r0:r1, c0:c1]
        if not inside.any() or (win[inside] != 0).any():
            fails += 1
            continue
        win[inside] = next_id
        next_id += 1
    n_placed = next_id - 1
    # tiny objects: 1..6 pixels, straight segments
    for k in range(n_tiny):
        for _ in range(max_tries):
            npx = 1 + (k % 6)
            r = int(rng.integers(0, H))
            c = int(rng.integers(0, max(1, W - npx)))
            if (lab[r, c : c + npx] == 0).all() and next_id < 65535:
                lab[r, c : c + npx] = next_id
                next_id += 1
                break
    n_total = next_id - 1
    # delete ~1 % of ids (never the largest id, so that max() keeps the row count)
    n_del = int(round(delete_frac * n_placed))
    if n_del and n_total > 2:
        dead = rng.choice(np.arange(1, n_total), size=min(n_del, n_total - 1), replace=False)
        lab[np.isin(lab, dead)] = 0
    return lab


def pixels_for_labels(
    rng: np.random.Generator,
    labels: np.ndarray,
    n_channels: int,
    n_z: int = 1,
    special: bool = True,
) -> np.ndarray:
    """uint16 pixels ``(C, Z, Y, X)`` for one label plane."""
    H, W = labels.shape
    n_lab = int(labels.max())
    out = np.empty((n_channels, n_z, H, W), dtype=np.uint16)
    for ch in range(n_channels):
        gains = np.exp(rng.uniform(np.log(200.0), np.log(20000.0), size=n_lab + 1))
        gains[0] = 0.0
        lam = gains[labels]
        for z in range(n_z):
            img = rng.poisson(300.0, size=(H, W)).astype(np.int64) + 100
            fg = labels > 0
            img[fg] += rng.poisson(lam[fg] * (1.0 - 0.5 * z / max(1, n_z)))
            out[ch, z] = np.clip(img, 0, 65535).astype(np.uint16)
    if special and n_lab >= 2:
        present = np.unique(labels)
        present = present[present > 0]
        if len(present) >= 2:
            out[:, :, labels == present[0]] = 65535  # saturated object
            out[:, :, labels == present[1]] = 1234  # constant object
        if len(present) >= 3:
            out[:, :, labels == present[2]] = 0  # all-zero object (median == 0 -> NaN rule)
    return out


def make_field(
    seed: int,
    shape: tuple[int, int] = (1080, 1080),
    n_channels: int = 2,
    n_objects: int = 300,
    n_z: int = 1,
    semi_axes: tuple[float, float] = (8.0, 30.0),
):
    """One field: ``pixels (1, C, Z, Y, X) uint16`` and ``labels (Y, X) uint16``."""
    rng = np.random.default_rng(seed)
    labels = ellipse_labels(rng, shape, n_objects, semi_axes=semi_axes)
    pixels = pixels_for_labels(rng, labels, n_channels, n_z)[None]
    return pixels, labels


def make_trap_position(
    seed: int,
    n_tp: int = 4,
    n_channels: int = 5,
    frame: tuple[int, int] = (1200, 1200),
    n_tiles: int = 40,
    tile_size: int = 96,
    max_cells: int = 8,
):
    """Yeast-style position (config C3): frames, tile centres and per-tile labels.

    Returns ``frames (T, C, 1, H, W) uint16``, ``centres (n_tiles, 2) int`` as
    (row, col) and ``labels (T, n_tiles, tile, tile) uint16``.
    """
    rng = np.random.default_rng(seed)
    H, W = frame
    half = tile_size // 2
    g = int(np.ceil(np.sqrt(n_tiles)))
    step_r = (H - 2 * half - 8) // g
    step_c = (W - 2 * half - 8) // g
    centres = []
    for i in range(n_tiles):
        gr, gc = divmod(i, g)
        r = half + 4 + gr * step_r + int(rng.integers(0, max(1, step_r - tile_size)))
        c = half + 4 + gc * step_c + int(rng.integers(0, max(1, step_c - tile_size)))
        centres.append((min(r, H - half - 1), min(c, W - half - 1)))
    centres = np.asarray(centres, dtype=np.int64)
    labels = np.zeros((n_tp, n_tiles, tile_size, tile_size), dtype=np.uint16)
    frames = np.empty((n_tp, n_channels, 1, H, W), dtype=np.uint16)
    for t in range(n_tp):
        full_lab = np.zeros((H, W), dtype=np.uint16)
        for i in range(n_tiles):
            n_cells = int(rng.integers(0, max_cells + 1))
            lab = ellipse_labels(
                rng, (tile_size, tile_size), n_cells, semi_axes=(5.0, 14.0), n_tiny=0, delete_frac=0.0
            )
            labels[t, i] = lab
            r0, c0 = centres[i, 0] - half, centres[i, 1] - half
            # unique ids in the full frame only matter for drawing the pixels
            full_lab[r0 : r0 + tile_size, c0 : c0 + tile_size] = np.where(lab > 0, lab + 10 * i, 0)
        frames[t] = pixels_for_labels(rng, full_lab, n_channels, 1, special=False)
    return frames, centres, labels
