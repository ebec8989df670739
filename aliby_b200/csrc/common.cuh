// Internal declarations shared by the kernels of libaliby_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "aliby_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;
typedef long long i64;

// Raw per-(object, request) statistics written by object_stats, read by finalize.
struct ChanStats {
  u64 sum;         // sum of x
  u64 sumsq;       // sum of x*x (mod 2^64)
  u64 wrapsq;      // sum of (x*x mod 2^bits(pixel dtype)) — NumPy's v**2 in the image dtype
  u64 m10, m01;    // sum x*c', sum x*r' with (r', c') relative to the bbox origin
  u64 m20, m02;    // sum x*c'^2, sum x*r'^2
  u64 top2p5_sum;  // sum of the ceil(0.025 n) largest values
  u64 top5_sum;    // sum of the min(5, n) largest values
  u32 vmin, vmax;
  u32 med_lo, med_hi;  // the two middle order statistics (equal for odd n)
  // cp_measure `intensity` (object_sweep.cu only): order statistics i = floor(n f) and i + 1 for f = 1/4, 1/2, 3/4;
  // the same pair for floor(|2 v - 2 median| / 2) with f = 1/2 (MAD; bit 31 of mad_hi: the deviations are half-integers);
  // position of the first maximum, (row << 16) | column relative to the bounding box
  u32 q[6];
  u32 mad_lo, mad_hi;
  u32 maxpos;
  u32 pad_;
};
static_assert(sizeof(ChanStats) == 128, "ChanStats layout");

// The same 96-byte slot for a request whose values are floating point (float pixels or the `div` reducer).
struct FloatStats {
  double sum, sumsq;      // sum x, sum x^2
  double css;             // sum (x - mean)^2  (two-pass, like np.std)
  double moi;             // moment_of_inertia, final value (cell.py:232-265)
  double top2p5_sum, top5_sum;
  double med_lo, med_hi;  // the two middle order statistics
  double vmin, vmax;
  u32 has_nan;            // any NaN among the object's values: every statistic is NaN, like NumPy
  u32 pad_;
};
static_assert(sizeof(FloatStats) <= sizeof(ChanStats), "FloatStats must fit the ChanStats slot");

__host__ __device__ __forceinline__ bool request_is_float(int pixel_dtype, int reduction) {
  return pixel_dtype == ABX_F32 || pixel_dtype == ABX_F64 || reduction == ABX_RED_DIV;
}

// Raw per-object output of the three chained EDTs (cell.py:207-229).
struct ShapeStats {
  double sum_nn;   // sum of sqrt(nn^2) over the object   (conical_volume / 4)
  double sum_top;  // sum of the plateau EDT over the cone top
  u32 max_nn2;     // max squared distance to the background
  u32 max_dn2;     // max squared distance to the cone top
};

// What the plan kernel (object_sweep.cu) hands the statistics kernel for one window-sized object: where its pixel window
// starts in the caller's buffer seen as rows of row_stride elements (the TMA coordinates of channel 0), and the window
// geometry.  geom = (h - 1) | (w - 1) << 6 | (rmin & 63) << 12 | (cmin & 63) << 18 | s_px << 24, with s_px the number
// of columns between the 16-byte aligned start of the TMA box and the bounding box.
struct ObjPlan {
  int32_t tma_x, tma_y;
  u32 n;
  u32 geom;
};

constexpr int kCounterWords = 32;  // err[0] + list_counts[31], zeroed by the label scan
// indices into Workspace::list_counts
constexpr int kCntStatsList = 0, kCntEdtList = 1, kCntGather = 2, kCntLeftover = 3, kCntEdtWork = 4, kCntLeftoverWork = 6,
              kCntRest = 8, kCntOrderBig = 12, kCntOrderSmall = 13, kCntSweepWork = 14, kCntEdtBig = 15, kCntEdtSmall = 16,
              kCntPairWork = 17, kCntPairWide = 19;

// Raw second moments of an object's pixel coordinates relative to its bounding box origin (plan kernel, from the bitmap)
struct MaskMoments {
  u64 s_rr, s_cc, s_rc;  // sum r^2, sum c^2, sum r c
  u64 pad_;
};

// Raw per-(object, pair of requests) sums of the two-image features (object_pair.cu), all exact integers.
// x, y: the object's values in the two requests; "both": pixels with x >= tx and y >= ty.
struct PairStats {
  u64 sxy;           // sum x y over the object
  u64 tot_x, tot_y;  // sum x [x >= tx], sum y [y >= ty]
  u64 cx, cy;        // sum x, sum y over both
  u64 cxy, cxx, cyy; // sum x y, x^2, y^2 over both
  u64 wx, wy;        // sum x (R - |rank x - rank y|), the same for y, over both
  u32 big_r;         // R = max(number of distinct x, number of distinct y)
  u32 n_both;        // pixels in both
  u32 flags;         // bit 0: not computed (values of 65536 or more)
  u32 pad_;
};
static_assert(sizeof(PairStats) == 96, "PairStats layout");

struct Workspace {
  abx_object_rec* recs;  // [n_objects + n_planes]
  PairStats* pairs;      // [n_objects * n_pairs]
  int* pair_wide;        // [n_objects * n_pairs] items the warp kernel of object_pair.cu leaves to its CTA kernel
  MaskMoments* mom;      // [n_objects] when need_edt bit 2
  u64* bitmaps;          // [n_objects + 1][64] torus bitmaps written by the label scan (label_scan.cu)
  ObjPlan* plan;         // [n_objects + n_planes]
  int* order_stats;      // [n_objects + n_planes] work order of the statistics kernel: big objects from the front, others from the back
  int* order_edt;        // [n_objects] the same for the shape kernel
  int* pair_list;        // [(n_objects + n_planes) * n_requests] (object * n_requests + request) pairs left to object_stats_warp
  ChanStats* chan;       // [(n_objects + n_planes) * n_requests]
  ShapeStats* shape;     // [n_objects]
  unsigned char* edt_scratch;
  size_t edt_scratch_per_cta;
  u32* err;          // [0] error flags, [1] stats_list length, [2] edt_list length
  int* stats_list;   // objects the warp kernel handed to the CTA statistics kernel
  int* edt_list;     // objects the warp kernel handed to the CTA EDT kernel
  u32* list_counts;  // = err + 1, indexed by the kCnt* constants above
  double* sqrt_tab;  // sqrt(d2) of every squared distance the warp EDT can produce
  u32* bg_hist;      // [n_planes][n_requests][65536] value histograms of large-plane backgrounds (background.cu)
  // Z stacks reduced up front (zreduce.cu); all null / unused when abx_zreduce_ok() is false
  void* zplanes;            // [n_tiles][n_requests] slots of H x W x 4 bytes: the Z-max plane (pixel dtype) or the Z-add
                            // plane (uint32) of request q of every tile
  abx_request* req_tma;     // [n_requests] requests as the kernels on the reduced planes see them
  abx_request* req_rest;    // [n_requests] requests as the gather pass over the original stack sees them
  i64* ztile_offset;        // [2][n_tiles] tile offsets inside zplanes, in pixel-dtype elements and in uint32 elements
  u32* zflags;              // [0]: some request is a Z-add (the gather pass on the uint32 planes has work)
  size_t total;
};

int abx_set_error(int code, const char* fmt, ...);
int abx_check_cuda(cudaError_t e, const char* what);
int abx_plan_workspace(const abx_extract_args* a, void* base, Workspace* ws);
int abx_validate(const abx_extract_args* a);

int launch_label_scan(const abx_extract_args* a, abx_object_rec* recs, u32* err /* [kCounterWords] */, u64* bitmaps,
                      double* sqrt_tab, int n_sqrt, cudaStream_t st);
int launch_plan(const abx_extract_args* a, const Workspace& ws, cudaStream_t st, bool sweep_ok);
int launch_object_sweep(const abx_extract_args* a, const Workspace& ws, cudaStream_t st);
bool abx_sweep_ok(const abx_extract_args* a);
int launch_object_stats_warp(const abx_extract_args* a, const Workspace& ws, cudaStream_t st, bool todo);
int launch_object_edt_warp(const abx_extract_args* a, const Workspace& ws, cudaStream_t st);
int launch_object_stats(const abx_extract_args* a, const Workspace& ws, cudaStream_t st, bool all_objects);
int launch_shape_edt(const abx_extract_args* a, const Workspace& ws, cudaStream_t st);
int launch_object_float(const abx_extract_args* a, const Workspace& ws, cudaStream_t st);
int launch_object_pair(const abx_extract_args* a, const Workspace& ws, cudaStream_t st);  // after every statistics kernel
int launch_big_background(const abx_extract_args* a, const Workspace& ws, cudaStream_t st);
bool abx_big_background(const abx_extract_args* a);
size_t abx_big_background_bytes(const abx_extract_args* a);
int launch_finalize(const abx_extract_args* a, const Workspace& ws, cudaStream_t st);  // also copies the error flags to a->status
bool abx_zreduce_ok(const abx_extract_args* a);
size_t abx_zreduce_bytes(const abx_extract_args* a);
int launch_zreduce(const abx_extract_args* a, const Workspace& ws, cudaStream_t st);
int launch_object_stats_rest(const abx_extract_args* a, const Workspace& ws, cudaStream_t st);
int abx_sqrt_table_entries();

constexpr int kEdtLargeCtas = 8;           // CTAs that own a whole-plane EDT scratch slot
constexpr int kEdtSmemWindow = 96 * 96;    // padded window (pixels) that is handled in shared memory
constexpr int kEdtBytesPerPixel = 8;       // mask(1) + flags(1) + g(2) + d2(4)
constexpr long long kBigBackground = 128 * 128;  // planes above this many pixels take the streaming background path

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }

// plane of an object row: largest p with plane_base[p] <= obj
__device__ __forceinline__ int find_plane(const int32_t* __restrict__ plane_base, int n_planes, int obj) {
  int lo = 0, hi = n_planes;  // invariant: base[lo] <= obj < base[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (plane_base[mid] <= obj) lo = mid; else hi = mid;
  }
  return lo;
}
