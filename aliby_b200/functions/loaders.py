"""Registry views with the reference's loader API (``src/extraction/core/functions/loaders.py``).

``load_funs()`` returns ``(CELL_FUNS, TRAP_FUNS, ALL_FUNS)`` like loaders.py:92-107.  Every
entry keeps the uniform ``f(mask, pixels)`` call of the reference (loaders.py:56-66) but is
served by the CUDA path: the boolean mask becomes a one-object label plane and the metric is
evaluated by the same kernels that serve whole trees.  The 18 cell names and 2 trap names
are exactly the reference's registry; cp_measure entries are absent
(no kernel, parity unpinned) and ``max``/``min``/``bbox_*`` are extensions.
"""

from __future__ import annotations

import typing as t

import numpy as np


class DeviceMetric:
    """``f(mask, pixels)`` evaluated on the GPU for a single boolean mask."""

    def __init__(self, name: str, needs_pixels: bool, background: bool = False):
        self.name = name
        self.needs_pixels = needs_pixels
        self.background = background
        self.__name__ = name

    def __call__(self, mask: np.ndarray, pixels: np.ndarray | None = None):
        from .. import engine
        from ..extract import _results_from_dense, _run_dense

        mask = np.asarray(mask)
        if self.background:
            # trap.py:6-43: masks are (Y, X, N) one-hot planes, pixels the tile image
            occupied = mask.sum(axis=2).astype(bool) if mask.size else np.zeros_like(pixels, dtype=bool)
            plane = occupied.astype(np.uint16)
            # a phantom object keeps one table row alive when the tile has no cell at all
            n_lab = 1
        else:
            plane = mask.astype(bool).astype(np.uint16)
            n_lab = 1
        if self.needs_pixels:
            if pixels is None:
                raise TypeError("'NoneType' object is not subscriptable")
            px = np.asarray(pixels)
            if self.name == "ratio":  # cell.py:268-279: NaN unless the image is (Y, X, 2)
                if px.ndim == 3 and px.shape[-1] == 2:
                    _ratio_unsupported()
                return np.nan
            inst = [(0, "max", self.name)]
            px5 = px[None, None, None]
        else:
            inst = [("None", "None", self.name)]
            px5 = np.zeros((1, 1, 1, *plane.shape), dtype=np.uint16)
        plan = engine.compile_instructions(inst)
        dense = _run_dense(plan, [plane], np.zeros(1, np.int32), np.array([n_lab]), px5)
        return _results_from_dense(plan, dense, np.zeros(1, np.int64), np.zeros(1, np.int64))[0]


def _ratio_unsupported():
    raise NotImplementedError("cell.ratio on a two-channel (Y, X, 2) image has no CUDA kernel (never produced by the pipelines)")


def load_cellfuns_core() -> dict:
    from ..engine import CELL_FUN_NAMES, SHAPE_METRICS

    return {name: DeviceMetric(name, needs_pixels=name not in SHAPE_METRICS) for name in CELL_FUN_NAMES}


def load_cellfuns(cp_measure_kwargs: t.Mapping[str, t.Mapping[str, t.Any]] | None = None) -> dict:
    """loaders.py:28-79 without the cp_measure entries (``cp_measure_kwargs`` is accepted and unused)."""
    from ..engine import EXTENSION_NAMES, SHAPE_METRICS

    funs = load_cellfuns_core()
    funs.update({name: DeviceMetric(name, needs_pixels=name not in SHAPE_METRICS) for name in EXTENSION_NAMES})
    return funs


def load_trapfuns() -> dict:
    from ..engine import TRAP_FUN_NAMES

    return {name: DeviceMetric(name, needs_pixels=True, background=True) for name in TRAP_FUN_NAMES}


def load_funs(cp_measure_kwargs: t.Mapping[str, t.Mapping[str, t.Any]] | None = None):
    CELL_FUNS = load_cellfuns(cp_measure_kwargs=cp_measure_kwargs)
    TRAP_FUNS = load_trapfuns()
    return CELL_FUNS, TRAP_FUNS, {**TRAP_FUNS, **CELL_FUNS}


def load_redfuns() -> dict:
    """REDUCTION_FUNS of loaders.py:110-127 (only the ufuncs are legal reducers)."""
    return {"max": np.maximum, "mean": np.mean, "median": np.median, "div": np.divide, "add": np.add, "None": None}
