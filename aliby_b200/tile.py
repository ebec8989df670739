"""Fused tile crop: tile windows addressed inside the full frame instead of copied out of it.

The reference's Tiler (``src/aliby/tile/tiler.py:309-366``) loads each channel of a time
point and slices one ``(Z, h, w)`` window per tile (``tiles.py:109-166`` gives the window:
``start = int(centre - sum(drifts[:tp+1])) - size // 2`` on each axis), stacking them into a
``(tiles, C, Z, h, w)`` copy.  Here the frame is uploaded once and a :class:`TileView`
records only the per-tile origins; the extraction kernels read the windows in place (each
pixel leaves HBM once), and :meth:`TileView.materialize` produces the reference's dense
array on the device for consumers that want it (a segmenter).

Windows that leave the frame follow ``if_out_of_bounds_pad`` (``tiler.py:601-650``): the tiles
are then materialised on the device by ``abx_crop_tiles_padded`` (per-line median padding,
integers rounded half to even), and a tile with more than 25 % padding becomes a NaN tile, which
— exactly like ``np.stack`` of a float64 NaN tile with uint16 tiles in the reference — promotes
the whole ``pixels`` array to float64 (served by the float kernel).  This is the cold path: the
reference filters edge tiles at time point 0 (``tiler.py:685-690``) and its drift is pinned to
zero in this fork.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


def tile_origins(centres, size, drifts=(), tp: int = 0) -> np.ndarray:
    """First/second-axis start of every tile window at time ``tp`` (tiles.py:109-166)."""
    size = (size, size) if isinstance(size, int) else tuple(size)
    centres = np.asarray(centres)
    shift = np.sum(np.asarray(drifts, dtype=float).reshape(-1, 2)[: tp + 1], axis=0) if len(drifts) else np.zeros(2)
    at_tp = (centres - shift).astype(int)  # truncation toward zero, like Tile.centre_at_time
    half = np.array([size[0] // 2, size[1] // 2])
    return (at_tp - half).astype(np.int64)


class TileView:
    """Lazy ``(tiles, C, Z, h, w)`` view of windows inside one ``(C, Z, H, W)`` frame."""

    def __init__(self, frame, origins, size):
        self.frame = frame  # numpy array or torch tensor (host or device)
        self.origins = np.asarray(origins, dtype=np.int64).reshape(-1, 2)
        self.size = (size, size) if isinstance(size, int) else tuple(size)
        C_, Z_, H, W = frame.shape
        h, w = self.size
        r0, c0 = self.origins[:, 0], self.origins[:, 1]
        # padding before / after on both axes (tiler.py:634-639)
        pad = np.stack([np.maximum(0, -r0), np.maximum(0, r0 + h - H), np.maximum(0, -c0), np.maximum(0, c0 + w - W)], axis=1)
        self.out_of_frame = pad.any(axis=1)
        # tiler.py:644: ``(padding / 0.25 > tile_shape).any()`` broadcasts the (2, 2) padding against [h, w]
        # column-wise: the "before" pads are compared with h, the "after" pads with w (reproduced as written)
        self.nan_tiles = ((pad[:, [0, 2]] / 0.25 > h) | (pad[:, [1, 3]] / 0.25 > w)).any(axis=1)
        self._dev = None

    @property
    def shape(self):
        C_, Z_, _, _ = self.frame.shape
        return (len(self.origins), C_, Z_, *self.size)

    @property
    def dtype(self):
        return np.dtype(np.float64) if self.nan_tiles.any() else self.frame.dtype

    def __len__(self):
        return len(self.origins)

    def device_frame(self, device=None):
        import torch

        if self._dev is None or (device is not None and self._dev.device != torch.device(device)):
            if device is None:
                device = torch.device("cuda", torch.cuda.current_device())
            f = self.frame
            if isinstance(f, np.ndarray):
                if str(f.dtype) not in ("uint8", "uint16", "float32", "float64"):
                    raise NotImplementedError(f"pixel dtype {f.dtype} has no CUDA kernel in aliby_b200")
                f = torch.from_numpy(np.ascontiguousarray(f))
            self._dev = f.to(device, non_blocking=True).contiguous()
        return self._dev

    def addressing(self, device=None):
        """(device pixels, tile element offsets, chan/z/row strides, C, Z) for the kernels."""
        if self.out_of_frame.any():  # cold path: materialised tiles with the reference's padding rules
            t = self.materialize(device)
            n, C_, Z_, h, w = t.shape
            return t, np.arange(n, dtype=np.int64) * (C_ * Z_ * h * w), Z_ * h * w, h * w, w, C_, Z_
        f = self.device_frame(device)
        C_, Z_, H, W = f.shape
        offs = self.origins[:, 0] * W + self.origins[:, 1]
        return f, offs.astype(np.int64), Z_ * H * W, H * W, W, C_, Z_

    def materialize(self, device=None):
        """Dense ``(tiles, C, Z, h, w)`` device tensor (what ``Tiler.get_fczyx`` returns, tiler.py:309-366),
        including the median padding / NaN tiles of windows that leave the frame (tiler.py:601-650)."""
        import torch

        f = self.device_frame(device)
        C_, Z_, H, W = f.shape
        h, w = self.size
        out = torch.empty((len(self.origins), C_, Z_, h, w), dtype=f.dtype, device=f.device)
        if len(self.origins) == 0:
            return out
        org = torch.from_numpy(self.origins.astype(np.int32)).to(f.device)
        dt = {torch.uint8: nat.U8, torch.uint16: nat.U16, torch.float32: nat.F32, torch.float64: nat.F64}[f.dtype]
        stream = C.c_void_p(torch.cuda.current_stream(f.device).cuda_stream)
        with torch.cuda.device(f.device):
            if self.out_of_frame.any():
                nat.check(
                    nat.lib().abx_crop_tiles_padded(
                        f.data_ptr(), dt, C_, Z_, Z_ * H * W, H * W, W, H, W, org.data_ptr(), len(self.origins), h, w,
                        out.data_ptr(), stream,
                    ),
                    "abx_crop_tiles_padded",
                )
            else:
                nat.check(
                    nat.lib().abx_crop_tiles(
                        f.data_ptr(), dt, C_, Z_, Z_ * H * W, H * W, W, org.data_ptr(), len(self.origins), h, w,
                        out.data_ptr(), stream,
                    ),
                    "abx_crop_tiles",
                )
        if self.nan_tiles.any():  # np.stack of a float64 NaN tile with the others promotes everything to float64
            out = out.to(torch.float64)
            out[torch.from_numpy(np.flatnonzero(self.nan_tiles)).to(f.device)] = float("nan")
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.materialize().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a


class FusedTiler:
    """Step object with the reference Tiler's ``run_tp`` contract (tiler.py:393-448) for
    pre-located tiles: returns ``{"drift": [0.0, 0.0], "pixels": TileView}``."""

    def __init__(self, pixels_tczyx, centres, tile_size):
        self.pixels = pixels_tczyx
        self.centres = np.asarray(centres)
        self.tile_size = tile_size

    def run_tp(self, tp: int):
        frame = self.pixels[tp]
        if hasattr(frame, "compute"):
            frame = frame.compute(scheduler="synchronous")
        view = TileView(np.asarray(frame), tile_origins(self.centres, self.tile_size), self.tile_size)
        return {"drift": [0.0, 0.0], "pixels": view}


def fuse_reference_tiler(tiler):
    """Swap the crop of a reference ``Tiler`` (tiler.py:309-366 ``get_fczyx``) for the fused :class:`TileView`.

    Everything else of the step stays the reference's: image reading, trap detection at time point 0
    (``set_areas_of_interest``), drift bookkeeping and the returned ``{"drift", "pixels"}`` dict (tiler.py:393-448).
    Only what ``pixels`` IS changes — a view that the extraction kernels address through tile offsets; a consumer that
    wants the dense ``(tiles, C, Z, h, w)`` array (a segmenter receiving ``tile.get_fczyx`` as passed method,
    pipe_builder.py:155-157) gets it from ``np.asarray(view)``, cropped on the device with the reference's padding
    rules."""

    def get_fczyx(tp: int):
        frame = tiler.pixels[tp]
        if hasattr(frame, "compute"):  # dask: fetch the frame once (the reference fetches it once per channel)
            frame = frame.compute(scheduler="synchronous")
        locs = tiler.tile_locs
        size = locs.tile_size if hasattr(locs, "tile_size") else tiler.tile_size
        origins = []
        for tile in locs.tiles:  # tiles.py:109-166: as_range(tp) = the two slices of the window at this time point
            rows, cols = tile.as_range(tp)
            origins.append((rows.start, cols.start))
        return TileView(np.asarray(frame), np.asarray(origins, dtype=np.int64).reshape(-1, 2), size)

    tiler.get_fczyx = get_fczyx  # instance attribute: shadows the method for run_tp and for passed methods
    return tiler
