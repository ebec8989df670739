/*
 * aliby_b200 — C-ABI of the B200 (sm_100a) per-object feature-extraction library.
 *
 * This is the drop-in boundary for ALIBY's extraction hot path.  The reference
 * (afermg/aliby) is pure Python and has no FFI of its own; each entry point below
 * names the reference code it replaces (paths relative to the reference root):
 *
 *   abx_label_scan     src/agora/utils/masks.py:35-37 (one-hot expansion, deleted) +
 *                      src/extraction/core/functions/cell.py:18-27,282-303 (area, centroid)
 *   abx_object_stats   src/extraction/extract.py:77-153 (measure/measure_mono loop),
 *                      src/extraction/core/functions/distributors.py:6-24 (Z reduction, fused),
 *                      src/extraction/core/functions/cell.py:43-157,232-265 (mean, total,
 *                      total_squared, median, max2p5pc, max5px_median, std, moment_of_inertia),
 *                      src/extraction/core/functions/trap.py:6-43 (background = label 0),
 *                      src/aliby/tile/tiler.py:309-366 (tile crop, fused through tile offsets)
 *                      Requests whose values are floating point (float32/float64 pixels, e.g. CropTiler's
 *                      standard_scale tiler.py:95-102, or the `div` reducer) take a generic fp64 kernel.
 *   abx_shape_edt      src/extraction/core/functions/cell.py:30-40,160-229 (eccentricity,
 *                      volume, conical_volume, min_maj_approximation: three chained EDTs)
 *   abx_extract (pairs) src/extraction/extract.py:200-237 (measure_multi: two channels, one mask) with the two-image
 *                      features that loaders.py:75-77,153-168 takes from cp_measure (CellProfiler MeasureColocalization)
 *   abx_finalize       the scalar arithmetic of the functions above + the dense
 *                      [objects x columns] table that replaces the long->wide pivot of
 *                      src/extraction/extract.py:574-598
 *   abx_extract        all four, stream-ordered (one call per extract step and timepoint,
 *                      i.e. what src/aliby/pipe_core.py:217 invokes)
 *   abx_crop_tiles     src/aliby/tile/tiler.py:309-366 materialised (tiles, C, Z, h, w)
 *   abx_crop_tiles_padded  the same with if_out_of_bounds_pad (tiler.py:601-650) for windows leaving the frame
 *
 * Conventions: plain C, no exceptions; every call returns 0 on success or a negative
 * abx_status, with a thread-local message behind abx_last_error().  The caller owns
 * every buffer (device memory unless stated otherwise); nothing is allocated on the
 * hot call — size the scratch with abx_extract_workspace_bytes().  All strides and
 * offsets are in ELEMENTS of the addressed array.  Calls are re-entrant across
 * streams and devices as long as every call in flight has its own workspace.  State the library keeps: the
 * thread-local error string, and per (host thread, device) two helper streams with three events (created on first use,
 * kept for the life of the thread) and the cached encodings of the last call's TMA descriptors.
 */
#ifndef ALIBY_B200_H
#define ALIBY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABX_VERSION 4

typedef enum abx_status {
  ABX_OK = 0,
  ABX_ERR_INVALID = -1,   /* bad argument (message says which) */
  ABX_ERR_UNSUPPORTED = -2, /* dtype / reduction without a kernel: no CPU fallback exists */
  ABX_ERR_CUDA = -3,      /* CUDA runtime error */
  ABX_ERR_WORKSPACE = -4  /* workspace too small */
} abx_status;

typedef enum abx_dtype { ABX_U8 = 0, ABX_U16 = 1, ABX_U32 = 2, ABX_F32 = 3, ABX_F64 = 4 } abx_dtype;

/* Z reductions of REDUCTION_FUNS (loaders.py:110-127) that are legal ufuncs.  DIV is np.divide.reduce: a
 * left fold of true divisions whose result is float64 whatever the pixel dtype (distributors.py:19-21). */
typedef enum abx_reduction { ABX_RED_MAX = 0, ABX_RED_ADD = 1, ABX_RED_DIV = 2 } abx_reduction;

/* Dense-table column kinds. 0-15 and 48-63 need only the label plane, 16-47 a pixel request, 64-79 a pair of requests. */
typedef enum abx_metric {
  ABX_M_AREA = 0,
  ABX_M_CENTROID_X = 1,
  ABX_M_CENTROID_Y = 2,
  ABX_M_SPHERICAL_VOLUME = 3,
  ABX_M_ECCENTRICITY = 4,
  ABX_M_VOLUME = 5,
  ABX_M_CONICAL_VOLUME = 6,
  ABX_M_MINOR_AXIS = 7,
  ABX_M_MAJOR_AXIS = 8,
  ABX_M_BBOX_RMIN = 9,
  ABX_M_BBOX_RMAX = 10,
  ABX_M_BBOX_CMIN = 11,
  ABX_M_BBOX_CMAX = 12,
  ABX_M_MEAN = 16,
  ABX_M_TOTAL = 17,
  ABX_M_TOTAL_SQUARED = 18,
  ABX_M_STD = 19,
  ABX_M_MEDIAN = 20,
  ABX_M_MAX2P5PC = 21,
  ABX_M_MAX5PX_MEDIAN = 22,
  ABX_M_MOMENT_OF_INERTIA = 23,
  ABX_M_RATIO = 24,
  ABX_M_MAX = 25,
  ABX_M_MIN = 26,
  ABX_M_IMBACKGROUND = 27,
  ABX_M_BACKGROUND_MAX5 = 28,
  /* cp_measure `intensity` (loaders.py:71-73,135-150; CellProfiler MeasureObjectIntensity without the edge features).
   * Integrated / Mean / Std / Min / Max are ABX_M_TOTAL / MEAN / STD / MIN / MAX.  Quartiles, median and MAD follow
   * CellProfiler's rank rule (i = floor(n f), linear interpolation to i + 1), not np.median.  Positions are 0-based
   * plane coordinates.  Needs a pixel request of uint8 / uint16 pixels (ABX_ERR_UNSUPPORTED otherwise).  Objects of any
   * size and any layout: the sweep kernel for windows <= 64 x 64 in a TMA-addressable layout, the CTA-per-object kernel
   * for everything else.  A request under the `div` reducer, or under `add` on the reduced planes of a Z stack, is served
   * by kernels without these statistics: status bit 1. */
  ABX_M_CP_LOWER_QUARTILE = 32,
  ABX_M_CP_MEDIAN = 33,
  ABX_M_CP_UPPER_QUARTILE = 34,
  ABX_M_CP_MAD = 35,
  ABX_M_CP_MASS_DISPLACEMENT = 36,
  ABX_M_CP_CENTER_MASS_X = 37,
  ABX_M_CP_CENTER_MASS_Y = 38,
  ABX_M_CP_MAX_POS_X = 39,
  ABX_M_CP_MAX_POS_Y = 40,
  ABX_M_CP_ZERO = 41, /* Location_MaxIntensity_Z of a 2-D image: 0 (NaN for an absent label) */
  ABX_M_CP_CENTER_MASS_Z = 42, /* 0, NaN when the intensities sum to 0 (0 / 0 in CellProfiler) */
  /* cp_measure `sizeshape` subset (label plane only, like 0-15): bounding box with exclusive maxima, 0-based centroid,
   * equivalent diameter, extent, maximum / mean radius (distance to the background: need_edt bits 0 / 1), and
   * eccentricity / axis lengths from the second central moments of the coordinates (need_edt bit 2) */
  ABX_M_CP_BBOX_AREA = 48,
  ABX_M_CP_BBOX_MAX_X = 49,
  ABX_M_CP_BBOX_MAX_Y = 50,
  ABX_M_CP_CENTER_X = 51,
  ABX_M_CP_CENTER_Y = 52,
  ABX_M_CP_EQUIVALENT_DIAMETER = 53,
  ABX_M_CP_EXTENT = 54,
  ABX_M_CP_MAXIMUM_RADIUS = 55,
  ABX_M_CP_MEAN_RADIUS = 56,
  ABX_M_CP_ECCENTRICITY = 57,
  ABX_M_CP_MAJOR_AXIS_LENGTH = 58,
  ABX_M_CP_MINOR_AXIS_LENGTH = 59,
  /* Two-image features of `extractmulti_*` steps (extract.py:200-237; CellProfiler MeasureColocalization for objects,
   * one object at a time).  abx_column.request indexes pairs[].  x, y: the object's values in the pair's two requests;
   * tx = threshold_fraction * max x, ty alike; "both" = pixels with x >= tx and y >= ty (an object without such a pixel
   * has 0 in 65-71).  Integer pixels only (uint8 / uint16, values below 65536 after the Z reduction: otherwise status
   * bit 2). */
  ABX_M_CO_PEARSON = 64,   /* correlation of x and y over the object (NaN when either is constant) */
  ABX_M_CO_MANDERS_1 = 65, /* sum x [both] / sum x [x >= tx] */
  ABX_M_CO_MANDERS_2 = 66,
  ABX_M_CO_RWC_1 = 67,     /* sum x w [both] / sum x [x >= tx], w = (R - |rank x - rank y|) / R, dense ranks in the object */
  ABX_M_CO_RWC_2 = 68,
  ABX_M_CO_OVERLAP = 69,   /* sum x y / sqrt(sum x^2 sum y^2) over both */
  ABX_M_CO_K_1 = 70,       /* sum x y / sum x^2 over both */
  ABX_M_CO_K_2 = 71
} abx_metric;

/* What a pixel request has to compute (bit mask). Sums/min/max are always produced. */
#define ABX_F_MEDIAN 1u
#define ABX_F_TOP2P5 2u
#define ABX_F_TOP5 4u
#define ABX_F_WRAPSQ 8u   /* total_squared with the square wrapped in the pixel dtype */
#define ABX_F_MOI 16u
#define ABX_F_CPQ 32u     /* CellProfiler quartiles + median (six order statistics) */
#define ABX_F_CPMAD 64u   /* CellProfiler MAD and the position of the maximum (a second pass over the window) */
/* abx_extract_args.request_feature_union only: some request uses ABX_RED_DIV (the float kernel must run
 * even though the pixels are integers) */
#define ABX_F_HAS_DIV 0x40000000u

/* One (channel, Z-reduction) pair of the extraction tree. */
typedef struct abx_request {
  int32_t channel;
  int32_t reduction;   /* abx_reduction */
  uint32_t features;   /* ABX_F_* needed for cell objects */
  uint32_t bg_features; /* ABX_F_* needed for the per-plane background object (0 = none) */
} abx_request;

/* Two requests measured together on every object (one (channel, channel) branch of an extractmulti tree). */
#define ABX_PF_THRESHOLDED 1u /* the sums over "both" (Manders, overlap, K) */
#define ABX_PF_RWC 2u         /* the rank-weighted sums (a second pass and two rank tables) */
typedef struct abx_pair {
  int32_t request_a, request_b; /* indices into requests[]; both requests exist there (their sums, minima and maxima are used) */
  uint32_t features;            /* ABX_PF_*; sum x y is always produced */
  uint32_t pad_;
  double threshold_fraction;    /* thr / 100 of CellProfiler's "threshold as percentage of maximum intensity" (15 -> 0.15) */
} abx_pair; /* 24 bytes */

/* One column of the dense output table. */
typedef struct abx_column {
  int32_t request; /* index into requests[] (into pairs[] for ABX_M_CO_*), -1 for label-only metrics */
  int32_t metric;  /* abx_metric */
} abx_column;

/* Per-object record produced by abx_label_scan (also usable on its own). */
typedef struct abx_object_rec {
  uint64_t sum_row; /* sum of (row + 1) over the object's pixels */
  uint64_t sum_col; /* sum of (col + 1) */
  uint32_t rmin, rmax, cmin, cmax; /* inclusive bbox, plane coordinates (16-byte aligned) */
  uint32_t n;       /* area in pixels */
  uint32_t pad_[3];
} abx_object_rec; /* 48 bytes */

typedef struct abx_extract_args {
  /* label planes: [n_planes][H][W] */
  const void* labels;
  int32_t label_dtype; /* ABX_U16 (segment/dispatch.py:14-19 guarantees ids < 65536) */
  int32_t n_planes;
  int32_t H, W;
  int64_t label_plane_stride, label_row_stride;
  const int32_t* plane_tile; /* [n_planes] pixel tile read by each plane */
  const int32_t* plane_base; /* [n_planes + 1] exclusive prefix of max label per plane */
  int32_t n_objects;         /* == plane_base[n_planes] (host value) */
  int32_t with_background;   /* also reduce label 0 of every plane (per-tile background) */
  /* pixels: element (tile, ch, z, r, c) lives at
   *   pixels[tile_offset[tile] + ch*chan_stride + z*z_stride + r*row_stride + c]
   * which covers a dense (tiles,C,Z,h,w) array and a tile crop fused straight out of
   * full frames (tile_offset = frame base + row0*row_stride + col0). */
  const void* pixels;
  int32_t pixel_dtype; /* ABX_U8 | ABX_U16 (integer kernels) | ABX_F32 | ABX_F64 (float kernel: CropTiler / NaN tiles) */
  int32_t n_tiles;
  int32_t C, Z;
  const int64_t* tile_offset; /* [n_tiles] */
  int64_t chan_stride, z_stride, row_stride;
  /* plan */
  const abx_request* requests; /* device, [n_requests] */
  int32_t n_requests;
  const abx_column* columns;   /* device, [n_columns] */
  int32_t n_columns;
  int32_t need_edt;            /* bit 0: ECCENTRICITY/VOLUME/MINOR/MAJOR/CP_MAXIMUM_RADIUS requested, bit 1: CONICAL_VOLUME /
                                * CP_MEAN_RADIUS, bit 2: second coordinate moments (CP_ECCENTRICITY / CP_*_AXIS_LENGTH) */
  int32_t request_feature_union; /* OR of features|bg_features over requests (host copy), | ABX_F_HAS_DIV */
  /* output: [n_objects][n_columns] float64, row-major */
  double* table;
  /* scratch */
  void* workspace;
  size_t workspace_bytes;
  void* stream; /* cudaStream_t */
  /* optional: 6 events made by abx_event_create, recorded on `stream` before the label scan and after the
   * label scan / object_stats_warp / object_edt_warp / large-object kernels / finalisation */
  void* const* stage_events;
  /* elements of the caller's pixel buffer counted from `pixels` (0 = unknown).  Lets the library describe the buffer
   * to the TMA unit (whole rows of row_stride elements, out-of-buffer parts of a box zero-filled); without it, or for
   * layouts TMA cannot address (unaligned base / strides, Z stacks), the statistics kernel gathers with plain loads. */
  int64_t pixel_elems;
  /* optional (ABI 3): device uint32 that receives the call's error flags when the kernels have run, stream-ordered like
   * the table (copy it back together with the table).  Bit 0: a label above its plane's n_labels
   * (plane_base[p + 1] - plane_base[p]) was met — those pixels belong to no row of the table and the background
   * statistics of that plane are not meaningful; the caller passed a stale or wrong plane_base.  Bit 1: an
   * ABX_M_CP_* rank statistic (quartiles, MAD, maximum position) was requested of a request that a kernel without them
   * serves — the `div` reducer, or the `add` reducer of a Z stack: those cells are not meaningful.  Bit 2: an
   * ABX_M_CO_* column met an object with values of 65536 or more (a Z-add of a stack): that cell is NaN. */
  uint32_t* status;
  /* optional (ABI 4): pairs of requests for the ABX_M_CO_* columns */
  const abx_pair* pairs; /* device, [n_pairs] */
  int32_t n_pairs;
  int32_t pad_;
} abx_extract_args;

int abx_version(void);
const char* abx_last_error(void);

/* Bytes of scratch abx_extract needs for these shapes (labels/pixels pointers are not read). */
int abx_extract_workspace_bytes(const abx_extract_args* args, size_t* bytes);

/* Whole hot path: label scan -> per-object statistics -> EDT shape metrics -> dense table. */
int abx_extract(const abx_extract_args* args);

/* Stages, callable on their own (records: [n_objects + n_planes], background records last). */
int abx_label_scan(const abx_extract_args* args, abx_object_rec* records);

/* Per-plane maximum label (what masks.max() is to extract.py:279), out: device int32 [n_planes]. */
int abx_label_max(const void* labels, int32_t label_dtype, int32_t n_planes, int32_t H, int32_t W,
                  int64_t plane_stride, int64_t row_stride, int32_t* out_max, void* stream);

/* Materialised tile crop (tiler.py:309-366) for in-bounds windows:
 * out[(t, c, z, r, x)] = frame[c*chan_stride + z*z_stride + (row0[t]+r)*row_stride + col0[t]+x]. */
int abx_crop_tiles(const void* frame, int32_t dtype, int32_t C, int32_t Z, int64_t chan_stride,
                   int64_t z_stride, int64_t row_stride, const int32_t* tile_origin /* [n_tiles][2] */,
                   int32_t n_tiles, int32_t h, int32_t w, void* out, void* stream);

/* Tile crop with the reference's out-of-frame rule for windows that leave the H x W frame (tiler.py:601-650):
 * the part inside the frame is copied, missing rows then missing columns are filled with np.pad's per-line
 * "median" (integers rounded half to even).  Origins may be negative or beyond the frame.  Whether a tile has
 * too much padding (> 25 %: a NaN tile in the reference) is the caller's decision. */
int abx_crop_tiles_padded(const void* frame, int32_t dtype, int32_t C, int32_t Z, int64_t chan_stride,
                          int64_t z_stride, int64_t row_stride, int32_t H, int32_t W,
                          const int32_t* tile_origin /* [n_tiles][2] */, int32_t n_tiles, int32_t h, int32_t w,
                          void* out, void* stream);

/* 1 when `ptr` points into page-locked (pinned or registered) host memory, 0 for pageable host memory, < 0 on error.
 * The host side uses it to decide between a direct asynchronous copy and staging through its own pinned buffers (a
 * NumPy view of a pinned allocation is pinned whatever its wrapper reports). */
int abx_host_is_pinned(const void* ptr);

/* Timing events for abx_extract_args.stage_events (thin wrappers over cudaEvent_t). */
int abx_event_create(void** event);
int abx_event_destroy(void* event);
int abx_event_elapsed_ms(void* start, void* end, float* ms); /* both events must have completed */

#ifdef __cplusplus
}
#endif
#endif /* ALIBY_B200_H */
