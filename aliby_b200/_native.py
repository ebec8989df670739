"""ctypes binding of ``libaliby_b200.so`` (the C-ABI in ``include/aliby_b200.h``).

There is no CPU fallback: if the shared library is missing, :func:`lib` raises with the
command that builds it.  Importing this module never touches CUDA.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libaliby_b200.so")

# enums of include/aliby_b200.h
U8, U16, U32, F32, F64 = 0, 1, 2, 3, 4
RED_MAX, RED_ADD, RED_DIV = 0, 1, 2

METRIC = {
    "area": 0,
    "centroid_x": 1,
    "centroid_y": 2,
    "spherical_volume": 3,
    "eccentricity": 4,
    "volume": 5,
    "conical_volume": 6,
    "minor_axis": 7,
    "major_axis": 8,
    "bbox_rmin": 9,
    "bbox_rmax": 10,
    "bbox_cmin": 11,
    "bbox_cmax": 12,
    "mean": 16,
    "total": 17,
    "total_squared": 18,
    "std": 19,
    "median": 20,
    "max2p5pc": 21,
    "max5px_median": 22,
    "moment_of_inertia": 23,
    "ratio": 24,
    "max": 25,
    "min": 26,
    "imBackground": 27,
    "background_max5": 28,
    # cp_measure `intensity` / `sizeshape` building blocks (ABX_M_CP_*)
    "cp_lower_quartile": 32,
    "cp_median": 33,
    "cp_upper_quartile": 34,
    "cp_mad": 35,
    "cp_mass_displacement": 36,
    "cp_center_mass_x": 37,
    "cp_center_mass_y": 38,
    "cp_max_pos_x": 39,
    "cp_max_pos_y": 40,
    "cp_zero": 41,
    "cp_center_mass_z": 42,
    "cp_bbox_area": 48,
    "cp_bbox_max_x": 49,
    "cp_bbox_max_y": 50,
    "cp_center_x": 51,
    "cp_center_y": 52,
    "cp_equivalent_diameter": 53,
    "cp_extent": 54,
    "cp_maximum_radius": 55,
    "cp_mean_radius": 56,
    "cp_eccentricity": 57,
    "cp_major_axis_length": 58,
    "cp_minor_axis_length": 59,
    # two-image features of extractmulti steps (ABX_M_CO_*): the column's request indexes pairs[]
    "co_pearson": 64,
    "co_manders_1": 65,
    "co_manders_2": 66,
    "co_rwc_1": 67,
    "co_rwc_2": 68,
    "co_overlap": 69,
    "co_k_1": 70,
    "co_k_2": 71,
}
EDT_METRICS = {4, 5, 6, 7, 8, 55}       # need_edt bit 0 (three chained EDTs; 55: the maximum of the first)
CONICAL_METRICS = {6, 56}               # need_edt bit 1 (sum of the first EDT)
MOMENT_METRICS = {57, 58, 59}           # need_edt bit 2 (second coordinate moments)
CONICAL_METRIC = 6

F_MEDIAN, F_TOP2P5, F_TOP5, F_WRAPSQ, F_MOI, F_CPQ, F_CPMAD = 1, 2, 4, 8, 16, 32, 64
F_HAS_DIV = 0x40000000
PF_THRESHOLDED, PF_RWC = 1, 2

EXPORTS = (
    "abx_version",
    "abx_last_error",
    "abx_extract_workspace_bytes",
    "abx_extract",
    "abx_label_scan",
    "abx_label_max",
    "abx_crop_tiles",
    "abx_crop_tiles_padded",
    "abx_host_is_pinned",
    "abx_event_create",
    "abx_event_destroy",
    "abx_event_elapsed_ms",
)


class Request(C.Structure):
    _fields_ = [("channel", C.c_int32), ("reduction", C.c_int32), ("features", C.c_uint32), ("bg_features", C.c_uint32)]


class Pair(C.Structure):
    _fields_ = [("request_a", C.c_int32), ("request_b", C.c_int32), ("features", C.c_uint32), ("pad_", C.c_uint32),
                ("threshold_fraction", C.c_double)]


class Column(C.Structure):
    _fields_ = [("request", C.c_int32), ("metric", C.c_int32)]


class ObjectRec(C.Structure):
    _fields_ = [
        ("sum_row", C.c_uint64),
        ("sum_col", C.c_uint64),
        ("rmin", C.c_uint32),
        ("rmax", C.c_uint32),
        ("cmin", C.c_uint32),
        ("cmax", C.c_uint32),
        ("n", C.c_uint32),
        ("pad_", C.c_uint32 * 3),
    ]


class ExtractArgs(C.Structure):
    _fields_ = [
        ("labels", C.c_void_p),
        ("label_dtype", C.c_int32),
        ("n_planes", C.c_int32),
        ("H", C.c_int32),
        ("W", C.c_int32),
        ("label_plane_stride", C.c_int64),
        ("label_row_stride", C.c_int64),
        ("plane_tile", C.c_void_p),
        ("plane_base", C.c_void_p),
        ("n_objects", C.c_int32),
        ("with_background", C.c_int32),
        ("pixels", C.c_void_p),
        ("pixel_dtype", C.c_int32),
        ("n_tiles", C.c_int32),
        ("C", C.c_int32),
        ("Z", C.c_int32),
        ("tile_offset", C.c_void_p),
        ("chan_stride", C.c_int64),
        ("z_stride", C.c_int64),
        ("row_stride", C.c_int64),
        ("requests", C.c_void_p),
        ("n_requests", C.c_int32),
        ("columns", C.c_void_p),
        ("n_columns", C.c_int32),
        ("need_edt", C.c_int32),
        ("request_feature_union", C.c_int32),
        ("table", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("stream", C.c_void_p),
        ("stage_events", C.POINTER(C.c_void_p)),
        ("pixel_elems", C.c_int64),
        ("status", C.c_void_p),
        ("pairs", C.c_void_p),
        ("n_pairs", C.c_int32),
        ("pad_", C.c_int32),
    ]


ABI_VERSION = 4  # include/aliby_b200.h ABX_VERSION


class NativeError(RuntimeError):
    """Non-zero status from the C-ABI."""


_lib = None


def lib() -> C.CDLL:
    """Load the shared library once; fail loudly if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. aliby_b200 has no CPU fallback: build the CUDA library with "
            "`python -m aliby_b200.build` (needs nvcc, targets sm_100a)."
        )
    handle = C.CDLL(LIB_PATH)
    handle.abx_version.restype = C.c_int
    handle.abx_last_error.restype = C.c_char_p
    handle.abx_extract_workspace_bytes.argtypes = [C.POINTER(ExtractArgs), C.POINTER(C.c_size_t)]
    handle.abx_extract.argtypes = [C.POINTER(ExtractArgs)]
    handle.abx_label_scan.argtypes = [C.POINTER(ExtractArgs), C.c_void_p]
    handle.abx_label_max.argtypes = [
        C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
    ]
    handle.abx_crop_tiles.argtypes = [
        C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
        C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
    ]
    handle.abx_crop_tiles_padded.argtypes = [
        C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
        C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
    ]
    handle.abx_host_is_pinned.argtypes = [C.c_void_p]
    handle.abx_event_create.argtypes = [C.POINTER(C.c_void_p)]
    handle.abx_event_destroy.argtypes = [C.c_void_p]
    handle.abx_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    for name in EXPORTS:
        getattr(handle, name)  # AttributeError if the header and the library drifted apart
    if handle.abx_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI {handle.abx_version()}, this package needs {ABI_VERSION}: rebuild it "
                          "(python -m aliby_b200.build --force)")
    _lib = handle
    return handle


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().abx_last_error().decode("utf-8", "replace")
        if status == -2:
            raise NotImplementedError(f"{what}: {msg}")
        raise NativeError(f"{what} failed with status {status}: {msg}")
