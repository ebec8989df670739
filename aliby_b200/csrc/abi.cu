// C-ABI entry points of libaliby_b200 (see include/aliby_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "common.cuh"

namespace {
thread_local char g_error[512] = "";

constexpr size_t kAlign = 256;
size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }
}  // namespace

int abx_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int abx_check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return ABX_OK;
  return abx_set_error(ABX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int abx_validate(const abx_extract_args* a) {
  if (!a) return abx_set_error(ABX_ERR_INVALID, "args is NULL");
  if (a->label_dtype != ABX_U16)
    return abx_set_error(ABX_ERR_UNSUPPORTED, "label dtype %d has no kernel (uint16 only; ids < 65536)", a->label_dtype);
  if (a->n_planes < 0 || a->n_objects < 0 || a->n_requests < 0 || a->n_columns < 0)
    return abx_set_error(ABX_ERR_INVALID, "negative count");
  if (a->n_planes > 0 && (a->H <= 0 || a->W <= 0 || a->H > 32767 || a->W > 32767))
    return abx_set_error(ABX_ERR_INVALID, "plane size %d x %d outside [1, 32767]", a->H, a->W);
  if (a->n_requests > 0) {
    if (a->pixel_dtype != ABX_U8 && a->pixel_dtype != ABX_U16 && a->pixel_dtype != ABX_F32 && a->pixel_dtype != ABX_F64)
      return abx_set_error(ABX_ERR_UNSUPPORTED,
                           "pixel dtype %d has no kernel (uint8/uint16/float32/float64); there is no CPU fallback",
                           a->pixel_dtype);
    if (a->Z < 1 || a->C < 1) return abx_set_error(ABX_ERR_INVALID, "C and Z must be >= 1");
    if (a->row_stride < 1 || a->row_stride >= (1LL << 25))
      return abx_set_error(ABX_ERR_INVALID, "pixel row stride %lld outside [1, 2^25)", (long long)a->row_stride);
    if (a->pixel_dtype == ABX_U16 && a->Z > 65536) return abx_set_error(ABX_ERR_INVALID, "Z too large for 32-bit sums");
    if ((a->request_feature_union & (int)(ABX_F_CPQ | ABX_F_CPMAD)) && a->pixel_dtype != ABX_U8 && a->pixel_dtype != ABX_U16)
      return abx_set_error(ABX_ERR_UNSUPPORTED,
                           "cp_measure intensity statistics have a kernel for uint8/uint16 pixels only (pixel dtype %d); there is no "
                           "CPU fallback", a->pixel_dtype);
  }
  if (a->n_pairs < 0) return abx_set_error(ABX_ERR_INVALID, "negative count");
  if (a->n_pairs > 0) {
    if (a->n_requests < 1) return abx_set_error(ABX_ERR_INVALID, "pairs need requests");
    if (a->pixel_dtype != ABX_U8 && a->pixel_dtype != ABX_U16)
      return abx_set_error(ABX_ERR_UNSUPPORTED,
                           "two-image features have a kernel for uint8/uint16 pixels only (pixel dtype %d); there is no CPU fallback",
                           a->pixel_dtype);
  }
  return ABX_OK;
}

int abx_plan_workspace(const abx_extract_args* a, void* base, Workspace* ws) {
  size_t off = 0;
  unsigned char* b = static_cast<unsigned char*>(base);
  const size_t n_rec = (size_t)a->n_objects + (size_t)a->n_planes;
  ws->recs = reinterpret_cast<abx_object_rec*>(b + off);
  off = align_up(off + n_rec * sizeof(abx_object_rec));
  ws->chan = reinterpret_cast<ChanStats*>(b + off);
  off = align_up(off + n_rec * (size_t)a->n_requests * sizeof(ChanStats));
  ws->pairs = reinterpret_cast<PairStats*>(b + off);
  off = align_up(off + (size_t)a->n_objects * (size_t)a->n_pairs * sizeof(PairStats));
  ws->pair_wide = reinterpret_cast<int*>(b + off);
  off = align_up(off + (size_t)a->n_objects * (size_t)a->n_pairs * sizeof(int));
  ws->shape = reinterpret_cast<ShapeStats*>(b + off);
  off = align_up(off + (a->need_edt ? (size_t)a->n_objects * sizeof(ShapeStats) : 0));
  ws->err = reinterpret_cast<u32*>(b + off);
  ws->list_counts = ws->err + 1;
  off = align_up(off + kCounterWords * sizeof(u32));
  // torus bitmaps of the label scan (+ one dummy for out-of-range labels), routing of the plan kernel
  ws->bitmaps = reinterpret_cast<u64*>(b + off);
  off = align_up(off + ((size_t)a->n_objects + 1) * 512);
  ws->mom = reinterpret_cast<MaskMoments*>(b + off);
  off = align_up(off + ((a->need_edt & 4) ? (size_t)a->n_objects * sizeof(MaskMoments) : 0));
  ws->plan = reinterpret_cast<ObjPlan*>(b + off);
  off = align_up(off + n_rec * sizeof(ObjPlan));
  ws->order_stats = reinterpret_cast<int*>(b + off);
  off = align_up(off + n_rec * sizeof(int));
  ws->order_edt = reinterpret_cast<int*>(b + off);
  off = align_up(off + n_rec * sizeof(int));
  ws->pair_list = reinterpret_cast<int*>(b + off);
  off = align_up(off + n_rec * (size_t)(a->n_requests > 0 ? a->n_requests : 1) * sizeof(int));
  ws->sqrt_tab = reinterpret_cast<double*>(b + off);
  off = align_up(off + (a->need_edt ? (size_t)abx_sqrt_table_entries() * sizeof(double) : 0));
  ws->bg_hist = reinterpret_cast<u32*>(b + off);
  off = align_up(off + abx_big_background_bytes(a));
  ws->zplanes = b + off;
  off = align_up(off + abx_zreduce_bytes(a));
  const bool zr = abx_zreduce_ok(a);
  ws->req_tma = reinterpret_cast<abx_request*>(b + off);
  off = align_up(off + (zr ? (size_t)a->n_requests * sizeof(abx_request) : 0));
  ws->req_rest = reinterpret_cast<abx_request*>(b + off);
  off = align_up(off + (zr ? (size_t)a->n_requests * sizeof(abx_request) : 0));
  ws->ztile_offset = reinterpret_cast<i64*>(b + off);
  off = align_up(off + (zr ? 2 * (size_t)a->n_tiles * sizeof(i64) : 0));
  ws->zflags = reinterpret_cast<u32*>(b + off);
  off = align_up(off + (zr ? 16 : 0));
  ws->stats_list = reinterpret_cast<int*>(b + off);
  off = align_up(off + n_rec * sizeof(int));
  ws->edt_list = reinterpret_cast<int*>(b + off);
  off = align_up(off + n_rec * sizeof(int));
  ws->edt_scratch = b + off;
  ws->edt_scratch_per_cta = 0;
  if (a->need_edt) {
    const size_t plane_px = (size_t)(a->H + 2) * (size_t)(a->W + 2);
    if (plane_px > (size_t)kEdtSmemWindow) {
      ws->edt_scratch_per_cta = align_up(plane_px * kEdtBytesPerPixel);
      off = align_up(off + ws->edt_scratch_per_cta * kEdtLargeCtas);
    }
  }
  ws->total = off;
  return ABX_OK;
}

extern "C" int abx_version(void) { return ABX_VERSION; }

extern "C" const char* abx_last_error(void) { return g_error; }

extern "C" int abx_extract_workspace_bytes(const abx_extract_args* args, size_t* bytes) {
  if (!bytes) return abx_set_error(ABX_ERR_INVALID, "bytes is NULL");
  int rc = abx_validate(args);
  if (rc) return rc;
  Workspace ws;
  abx_plan_workspace(args, nullptr, &ws);
  *bytes = ws.total;
  return ABX_OK;
}

static int check_pointers(const abx_extract_args* a, const Workspace& ws) {
  if (a->n_planes > 0 && (!a->labels || !a->plane_tile || !a->plane_base))
    return abx_set_error(ABX_ERR_INVALID, "labels / plane_tile / plane_base is NULL");
  if (a->n_requests > 0 && (!a->pixels || !a->tile_offset || !a->requests))
    return abx_set_error(ABX_ERR_INVALID, "pixels / tile_offset / requests is NULL");
  if (a->n_pairs > 0 && !a->pairs) return abx_set_error(ABX_ERR_INVALID, "pairs is NULL");
  if (!a->workspace || a->workspace_bytes < ws.total)
    return abx_set_error(ABX_ERR_WORKSPACE, "workspace has %zu bytes, %zu needed", a->workspace_bytes, ws.total);
  if ((reinterpret_cast<uintptr_t>(a->workspace) & 15u) != 0)
    return abx_set_error(ABX_ERR_INVALID, "workspace must be 16-byte aligned");
  return ABX_OK;
}

extern "C" int abx_label_scan(const abx_extract_args* args, abx_object_rec* records) {
  int rc = abx_validate(args);
  if (rc) return rc;
  if (!records) return abx_set_error(ABX_ERR_INVALID, "records is NULL");
  if (!args->workspace || args->workspace_bytes < kCounterWords * sizeof(u32))
    return abx_set_error(ABX_ERR_WORKSPACE, "abx_label_scan needs a 128-byte workspace for its flags");
  return launch_label_scan(args, records, static_cast<u32*>(args->workspace), nullptr, nullptr, 0,
                           static_cast<cudaStream_t>(args->stream));
}

// Two helper streams and their fork / join events per (host thread, device), created on first use and kept.
struct Helper {
  cudaStream_t stream;   // the shape chain
  cudaStream_t stream2;  // the per-plane backgrounds
  cudaEvent_t fork, join, join2;
};
static Helper* helper_stream() {
  static thread_local Helper helpers[64];
  static thread_local bool made[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!made[dev]) {
    Helper h;
    if (cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h.stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h.join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h.join2, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    helpers[dev] = h;
    made[dev] = true;
  }
  return &helpers[dev];
}

extern "C" int abx_extract(const abx_extract_args* args) {
  int rc = abx_validate(args);
  if (rc) return rc;
  Workspace ws;
  abx_plan_workspace(args, args->workspace, &ws);
  rc = check_pointers(args, ws);
  if (rc) return rc;
  if (args->n_objects > 0 && args->n_columns > 0 && (!args->table || !args->columns))
    return abx_set_error(ABX_ERR_INVALID, "table / columns is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(args->stream);
  void* const* ev = args->stage_events;
  auto mark = [&](int i) { if (ev && ev[i]) cudaEventRecord(static_cast<cudaEvent_t>(ev[i]), st); };
  mark(0);
  const bool edt = (args->need_edt & 3) && args->n_objects > 0;
  if ((rc = launch_label_scan(args, ws.recs, ws.err, ws.bitmaps, edt ? ws.sqrt_tab : nullptr, abx_sqrt_table_entries(), st)))
    return rc;
  mark(1);
  // Z stacks: every requested (tile, channel) stack is reduced once, streaming, into planes of the workspace (max: pixel
  // dtype, add: uint32), and the window-sized objects then see a Z = 1 problem on those planes (zreduce.cu).
  abx_extract_args red = *args;  // what the per-object statistics kernels of window-sized objects work on
  const bool zred = abx_zreduce_ok(args);
  if (zred) {
    if ((rc = launch_zreduce(args, ws, st))) return rc;
    red.pixels = ws.zplanes;
    red.tile_offset = reinterpret_cast<const int64_t*>(ws.ztile_offset);
    red.requests = ws.req_tma;
    red.C = args->n_requests;
    red.Z = 1;
    const i64 slot = (i64)args->H * args->W * (args->pixel_dtype == ABX_U8 ? 4 : 2);  // a uint32-sized slot in pixel elements
    red.chan_stride = red.z_stride = slot;
    red.row_stride = args->W;
    red.pixel_elems = (i64)args->n_tiles * args->n_requests * slot;
  }
  // objects with a window <= 64 x 64: the sweep kernel (TMA-staged windows, lists from the label scan's bitmaps) when
  // the pixel layout allows, plain gathers otherwise.  The plan kernel routes every object first.
  const bool sweep = abx_sweep_ok(&red);
  if ((rc = launch_plan(&red, ws, st, sweep))) return rc;
  // From here the shape metrics (label planes and bitmaps only) and the intensity statistics are independent: the
  // shape kernels go to a helper stream.  A big launch gains the overlap of one chain's tail with the other's start; a
  // small one (a time point of a yeast position: a few hundred objects, every kernel a few microseconds) halves its
  // chain of dependent launches.  With stage events (a profiling run) the two chains stay in line on the caller's
  // stream so that each stage can be timed on its own.
  // The per-plane backgrounds (label 0: labels, pixels and the scan's counts only) are a third independent chain, on a
  // second helper stream: for a time point of a trap position they are the longest kernel of the call.
  Helper* hp = !ev ? helper_stream() : nullptr;
  const bool side_edt = hp && edt;
  // (a small call that is launched eagerly is bound by the host's launch calls, not by its kernels: the extra stream
  // operations would cost it more than the overlap returns — 0.089 -> 0.112 ms per C3 time point; replayed from a
  // graph the same call gains, 0.075 -> 0.058 ms.  So: while the caller is capturing, or when the call is large.)
  cudaStreamCaptureStatus capture = cudaStreamCaptureStatusNone;
  if (hp && cudaStreamIsCapturing(st, &capture) != cudaSuccess) { cudaGetLastError(); capture = cudaStreamCaptureStatusNone; }
  const bool side_bg = hp && abx_big_background(args) && args->n_planes > 0 &&
                       (capture == cudaStreamCaptureStatusActive || args->n_objects > 4096);
  if (side_edt || side_bg) cudaEventRecord(hp->fork, st);
  if (side_edt) {
    cudaStreamWaitEvent(hp->stream, hp->fork, 0);
    if ((rc = launch_object_edt_warp(args, ws, hp->stream))) return rc;
    if ((rc = launch_shape_edt(args, ws, hp->stream))) return rc;
  }
  if (side_bg) {
    cudaStreamWaitEvent(hp->stream2, hp->fork, 0);
    if ((rc = launch_big_background(args, ws, hp->stream2))) return rc;
    cudaEventRecord(hp->join2, hp->stream2);
  }
  if (sweep) {
    if ((rc = launch_object_sweep(&red, ws, st))) return rc;
  } else if ((rc = launch_object_stats_warp(&red, ws, st, false))) {
    return rc;
  }
  if (zred && (rc = launch_object_stats_rest(args, ws, st))) return rc;  // the Z-add requests, from their uint32 sum planes
  mark(2);
  if (!side_edt && edt && (rc = launch_object_edt_warp(args, ws, st))) return rc;
  // the few (object, request) pairs the sweep kernel left over (a window that starts in front of the buffer): the
  // gather kernel, one warp per CTA — behind the shape kernels on the helper stream when there is one (their cost is a
  // few warps' latency: off the caller's chain)
  if (sweep) {
    if (side_edt) {
      cudaEventRecord(hp->fork, st);  // the sweep kernel has written the pair list
      cudaStreamWaitEvent(hp->stream, hp->fork, 0);
    }
    if ((rc = launch_object_stats_warp(&red, ws, side_edt ? hp->stream : st, true))) return rc;
  }
  if (side_edt) cudaEventRecord(hp->join, hp->stream);
  mark(3);
  // the rest (large objects, Z-add backgrounds) — every object when cp_measure rank statistics are wanted in a layout the
  // sweep kernel cannot address (the gather kernel has none; this one reads with plain loads, any layout, any reduction)
  const bool cp_all = !sweep && (args->request_feature_union & (int)(ABX_F_CPQ | ABX_F_CPMAD)) != 0;
  if ((rc = launch_object_stats(args, ws, st, cp_all))) return rc;
  if (!side_bg && (rc = launch_big_background(args, ws, st))) return rc;  // per-plane backgrounds: tile or streaming histogram
  if ((rc = launch_object_float(args, ws, st))) return rc;  // floating-point requests (float pixels, `div`)
  if (args->n_pairs > 0) {  // two-image features: they start from the minima / maxima of both requests
    if (side_edt) cudaStreamWaitEvent(st, hp->join, 0);  // (the left-over pairs of the sweep kernel ran on the helper stream)
    if ((rc = launch_object_pair(args, ws, st))) return rc;
  }
  if (!side_edt && (rc = launch_shape_edt(args, ws, st))) return rc;
  mark(4);
  if (side_edt) cudaStreamWaitEvent(st, hp->join, 0);
  if (side_bg) cudaStreamWaitEvent(st, hp->join2, 0);
  if ((rc = launch_finalize(args, ws, st))) return rc;  // (also copies the error flags to args->status)
  mark(5);
  return ABX_OK;
}

extern "C" int abx_host_is_pinned(const void* ptr) {
  cudaPointerAttributes attr;
  const cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return abx_check_cuda(e, "cudaPointerGetAttributes");
  }
  return attr.type == cudaMemoryTypeHost ? 1 : 0;
}

extern "C" int abx_event_create(void** event) {
  if (!event) return abx_set_error(ABX_ERR_INVALID, "event is NULL");
  cudaEvent_t e;
  cudaError_t rc = cudaEventCreate(&e);
  if (rc != cudaSuccess) return abx_check_cuda(rc, "cudaEventCreate");
  *event = e;
  return ABX_OK;
}

extern "C" int abx_event_destroy(void* event) {
  return abx_check_cuda(cudaEventDestroy(static_cast<cudaEvent_t>(event)), "cudaEventDestroy");
}

extern "C" int abx_event_elapsed_ms(void* start, void* end, float* ms) {
  if (!ms) return abx_set_error(ABX_ERR_INVALID, "ms is NULL");
  return abx_check_cuda(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(end)),
                        "cudaEventElapsedTime");
}

// ---- materialised tile crop -------------------------------------------------------------------
namespace {
template <typename T>
__global__ void crop_tiles_kernel(const T* __restrict__ frame, int C, int Z, i64 chan_stride, i64 z_stride,
                                  i64 row_stride, const int32_t* __restrict__ origin, int n_tiles, int h, int w,
                                  T* __restrict__ out) {
  // grid.y enumerates (tile, c, z, row); threads walk the row
  const i64 line = blockIdx.x;
  const int r = (int)(line % h);
  i64 t = line / h;
  const int z = (int)(t % Z); t /= Z;
  const int c = (int)(t % C); t /= C;
  const int tile = (int)t;
  const T* src = frame + (i64)c * chan_stride + (i64)z * z_stride + (i64)(origin[2 * tile] + r) * row_stride +
                 origin[2 * tile + 1];
  T* dst = out + line * w;
  for (int x = threadIdx.x; x < w; x += blockDim.x) dst[x] = __ldg(src + x);
}
}  // namespace

extern "C" int abx_crop_tiles(const void* frame, int32_t dtype, int32_t C, int32_t Z, int64_t chan_stride,
                              int64_t z_stride, int64_t row_stride, const int32_t* tile_origin, int32_t n_tiles,
                              int32_t h, int32_t w, void* out, void* stream) {
  if (!frame || !tile_origin || !out || C < 1 || Z < 1 || n_tiles < 0 || h < 1 || w < 1)
    return abx_set_error(ABX_ERR_INVALID, "abx_crop_tiles: bad arguments");
  const i64 lines = (i64)n_tiles * C * Z * h;
  if (lines == 0) return ABX_OK;
  if (lines > 2147483647LL) return abx_set_error(ABX_ERR_INVALID, "abx_crop_tiles: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = w >= 128 ? 128 : (w >= 64 ? 64 : 32);
  if (dtype == ABX_U16)
    crop_tiles_kernel<uint16_t><<<(unsigned)lines, threads, 0, st>>>(static_cast<const uint16_t*>(frame), C, Z,
                                                                     chan_stride, z_stride, row_stride, tile_origin,
                                                                     n_tiles, h, w, static_cast<uint16_t*>(out));
  else if (dtype == ABX_U8)
    crop_tiles_kernel<uint8_t><<<(unsigned)lines, threads, 0, st>>>(static_cast<const uint8_t*>(frame), C, Z,
                                                                    chan_stride, z_stride, row_stride, tile_origin,
                                                                    n_tiles, h, w, static_cast<uint8_t*>(out));
  else if (dtype == ABX_F32)
    crop_tiles_kernel<float><<<(unsigned)lines, threads, 0, st>>>(static_cast<const float*>(frame), C, Z, chan_stride,
                                                                  z_stride, row_stride, tile_origin, n_tiles, h, w,
                                                                  static_cast<float*>(out));
  else if (dtype == ABX_F64)
    crop_tiles_kernel<double><<<(unsigned)lines, threads, 0, st>>>(static_cast<const double*>(frame), C, Z, chan_stride,
                                                                   z_stride, row_stride, tile_origin, n_tiles, h, w,
                                                                   static_cast<double*>(out));
  else
    return abx_set_error(ABX_ERR_UNSUPPORTED, "abx_crop_tiles: dtype %d has no kernel", dtype);
  return abx_check_cuda(cudaGetLastError(), "crop_tiles");
}

// ---- tile crop with the reference's out-of-frame rules (tiler.py:601-650) -----------------------------------
// np.pad(tile, [[0, 0], [top, bottom], [left, right]], "median") pads axis by axis: first the missing ROWS of
// every column get that column's median over the rows that exist, then the missing COLUMNS of every row (the new
// rows included) get that row's median over the columns that exist; integer medians are rounded half to even
// (numpy/lib/arraypad.py: _get_stats + _round_if_needed).  Tiles with more than 25 % padding become NaN tiles in
// the reference: that decision (and the float64 promotion it implies) is the caller's, see aliby_b200/tile.py.
namespace {
template <typename T>
__global__ void crop_clip_kernel(const T* __restrict__ frame, int C, int Z, i64 chan_stride, i64 z_stride, i64 row_stride,
                                 int H, int W, const int32_t* __restrict__ origin, int h, int w, T* __restrict__ out) {
  const i64 line = blockIdx.x;  // (tile, c, z, row)
  const int r = (int)(line % h);
  i64 t = line / h;
  const int z = (int)(t % Z); t /= Z;
  const int c = (int)(t % C); t /= C;
  const int tile = (int)t;
  const int fr = origin[2 * tile] + r, fc0 = origin[2 * tile + 1];
  T* dst = out + line * w;
  for (int x = threadIdx.x; x < w; x += blockDim.x) {
    const int fc = fc0 + x;
    dst[x] = (fr >= 0 && fr < H && fc >= 0 && fc < W)
                 ? frame[(i64)c * chan_stride + (i64)z * z_stride + (i64)fr * row_stride + fc] : T(0);
  }
}

template <typename T>
__device__ __forceinline__ T median_of_line(const T* __restrict__ base, i64 stride, int n) {
  // n <= a tile side: rank every element by counting (ties broken by position), pick the middle one(s)
  const int k_lo = (n - 1) / 2, k_hi = n / 2;
  double lo = 0, hi = 0;
  for (int i = 0; i < n; ++i) {
    const T xi = base[(i64)i * stride];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const T xj = base[(i64)j * stride];
      rank += (xj < xi) || (xj == xi && j < i);
    }
    if (rank == k_lo) lo = (double)xi;
    if (rank == k_hi) hi = (double)xi;
  }
  if (std::is_integral<T>::value) return (T)rint((lo + hi) / 2.0);  // np.round: half to even
  if (sizeof(T) == 4) return (T)(((float)lo + (float)hi) / 2.0f);    // float32 mean of the two middle values
  return (T)((lo + hi) / 2.0);
}

// axis 0: one thread per (plane, existing column): fill the missing rows with the column median
// axis 1: one thread per (plane, row): fill the missing columns with the row median over the existing columns
template <typename T>
__global__ void pad_median_kernel(T* __restrict__ out, const int32_t* __restrict__ origin, int n_tiles, int CZ, int H, int W,
                                  int h, int w, int axis) {
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_plane = axis == 0 ? w : h;
  if (idx >= (i64)n_tiles * CZ * per_plane) return;
  const int k = (int)(idx % per_plane);
  const i64 plane = idx / per_plane;
  const int tile = (int)(plane / CZ);
  const int r0 = origin[2 * tile], c0 = origin[2 * tile + 1];
  const int rv0 = max(0, -r0), rv1 = min(h, H - r0);  // rows of the tile that exist in the frame
  const int cv0 = max(0, -c0), cv1 = min(w, W - c0);
  if (rv1 <= rv0 || cv1 <= cv0) return;  // entirely outside: a NaN tile for the caller
  T* p = out + plane * (i64)h * w;
  if (axis == 0) {
    if ((rv0 == 0 && rv1 == h) || k < cv0 || k >= cv1) return;
    const T m = median_of_line(p + (i64)rv0 * w + k, (i64)w, rv1 - rv0);
    for (int r = 0; r < rv0; ++r) p[(i64)r * w + k] = m;
    for (int r = rv1; r < h; ++r) p[(i64)r * w + k] = m;
  } else {
    if (cv0 == 0 && cv1 == w) return;
    const T m = median_of_line(p + (i64)k * w + cv0, (i64)1, cv1 - cv0);
    for (int c = 0; c < cv0; ++c) p[(i64)k * w + c] = m;
    for (int c = cv1; c < w; ++c) p[(i64)k * w + c] = m;
  }
}

template <typename T>
int crop_padded(const void* frame, int C, int Z, i64 chan_stride, i64 z_stride, i64 row_stride, int H, int W,
                const int32_t* origin, int n_tiles, int h, int w, void* out, cudaStream_t st) {
  const i64 lines = (i64)n_tiles * C * Z * h;
  if (lines > 2147483647LL) return abx_set_error(ABX_ERR_INVALID, "abx_crop_tiles_padded: too many rows");
  const int threads = w >= 128 ? 128 : (w >= 64 ? 64 : 32);
  crop_clip_kernel<T><<<(unsigned)lines, threads, 0, st>>>(static_cast<const T*>(frame), C, Z, chan_stride, z_stride,
                                                          row_stride, H, W, origin, h, w, static_cast<T*>(out));
  for (int axis = 0; axis < 2; ++axis) {
    const i64 n = (i64)n_tiles * C * Z * (axis == 0 ? w : h);
    pad_median_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(static_cast<T*>(out), origin, n_tiles, C * Z, H, W, h,
                                                                       w, axis);
  }
  return abx_check_cuda(cudaGetLastError(), "crop_tiles_padded");
}
}  // namespace

extern "C" int abx_crop_tiles_padded(const void* frame, int32_t dtype, int32_t C, int32_t Z, int64_t chan_stride,
                                     int64_t z_stride, int64_t row_stride, int32_t H, int32_t W,
                                     const int32_t* tile_origin, int32_t n_tiles, int32_t h, int32_t w, void* out,
                                     void* stream) {
  if (!frame || !tile_origin || !out || C < 1 || Z < 1 || n_tiles < 0 || h < 1 || w < 1 || H < 1 || W < 1)
    return abx_set_error(ABX_ERR_INVALID, "abx_crop_tiles_padded: bad arguments");
  if (n_tiles == 0) return ABX_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case ABX_U8: return crop_padded<uint8_t>(frame, C, Z, chan_stride, z_stride, row_stride, H, W, tile_origin, n_tiles, h, w, out, st);
    case ABX_U16: return crop_padded<uint16_t>(frame, C, Z, chan_stride, z_stride, row_stride, H, W, tile_origin, n_tiles, h, w, out, st);
    case ABX_F32: return crop_padded<float>(frame, C, Z, chan_stride, z_stride, row_stride, H, W, tile_origin, n_tiles, h, w, out, st);
    case ABX_F64: return crop_padded<double>(frame, C, Z, chan_stride, z_stride, row_stride, H, W, tile_origin, n_tiles, h, w, out, st);
    default: return abx_set_error(ABX_ERR_UNSUPPORTED, "abx_crop_tiles_padded: dtype %d has no kernel", dtype);
  }
}
