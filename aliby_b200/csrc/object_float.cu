// Per-object statistics of FLOATING-POINT requests: float32 / float64 pixels (the CropTiler's standard_scale and
// clip_outliers yield float64, src/aliby/tile/tiler.py:75-102; NaN tiles of tiler.py:644-646 are float64) and the
// `div` Z-reducer, whose result is float64 whatever the pixel dtype (np.divide.reduce,
// src/extraction/core/functions/distributors.py:19-21).
//
// Same functions as the integer path (cell.py:43-157,232-265; trap.py:6-43) under NumPy's float semantics:
//   * sums in fp64 (NumPy's pairwise order differs in the last bits: parity is "within 1e-6", tested at 1e-9);
//     std from the two-pass centred sum like np.std; moment_of_inertia with cell.py's formula
//   * median / top-2.5 % / top-5 are EXACT order statistics: the values are mapped to order-preserving 64-bit
//     keys and selected by an MSB-first radix select, 8 bits per window sweep, skipping the leading bits that all
//     keys share; -0.0 and +0.0 compare equal in NumPy and may come back with either sign
//   * one NaN among an object's values makes every statistic NaN (np.sort puts NaNs last, np.median / np.mean /
//     np.partition-based top-k all return NaN)
//
// One CTA per object, any window size (the window is swept, nothing is staged): this is the generic, colder path —
// about ten sweeps of the bounding box per request — and also serves the per-plane background (label 0).
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr u64 kSignBit = 0x8000000000000000ull;
static_assert(kWarps == 4, "one warp per selection target");

__device__ __forceinline__ u64 key_of(double x) {  // order-preserving; -0.0 is folded onto +0.0
  if (x == 0.0) x = 0.0;
  const u64 b = (u64)__double_as_longlong(x);
  return (b & kSignBit) ? ~b : (b | kSignBit);
}
__device__ __forceinline__ double value_of(u64 k) {
  const u64 b = (k & kSignBit) ? (k & ~kSignBit) : ~k;
  return __longlong_as_double((long long)b);
}

// One Z-reduced value.  Integer pixels only get here for `div`.  float32 folds in float32 like NumPy does.
template <typename PX>
__device__ __forceinline__ double load_value(const PX* __restrict__ p, int Z, i64 zs, int red) {
  if (sizeof(PX) == 4 && !std::is_integral<PX>::value) {  // float32
    float x = (float)__ldg(p);
    for (int z = 1; z < Z; ++z) {
      const float y = (float)__ldg(p + (i64)z * zs);
      if (red == ABX_RED_MAX) x = (x != x || y != y) ? __int_as_float(0x7FC00000) : fmaxf(x, y);  // np.maximum propagates NaN
      else if (red == ABX_RED_ADD) x = x + y;
      else x = x / y;
    }
    return (double)x;
  }
  double x = (double)__ldg(p);
  for (int z = 1; z < Z; ++z) {
    const double y = (double)__ldg(p + (i64)z * zs);
    if (red == ABX_RED_MAX) x = (x != x || y != y) ? __longlong_as_double(0x7FF8000000000000ll) : fmax(x, y);
    else if (red == ABX_RED_ADD) x = x + y;
    else x = x / y;
  }
  return x;
}

struct Smem {
  u32 hist[4][256];
  double red[kWarps][8];
  u64 redk[kWarps][2];
  u32 redc[kWarps][2];
  u64 prefix[4];  // bits of the four target keys found so far (relative to kmin), aligned at bit 0
  u32 rank[4];    // rank of each target among the keys that share its prefix
  abx_object_rec rec;
  int plane, tile;
  u32 label;
};

template <typename PX>
struct Window {
  const uint16_t* lab;
  const PX* px;
  i64 lab_rs, px_rs, zs;
  u32 rmin, rmax, cmin, cmax, label;
  int Z, red;
};

// f(x, r, c) for every pixel of the object; every thread of the CTA takes part in the walk.
template <typename PX, class F>
__device__ __forceinline__ void sweep(const Window<PX>& w, F&& f) {
  const u32 warp = threadIdx.x >> 5, lane = lane_id();
  for (u32 r = w.rmin + warp; r <= w.rmax; r += kWarps) {
    const uint16_t* lrow = w.lab + (i64)r * w.lab_rs;
    const PX* prow = w.px + (i64)r * w.px_rs;
    for (u32 c = w.cmin + lane; c <= w.cmax; c += 32)
      if ((u32)__ldg(lrow + c) == w.label) f(load_value(prow + c, w.Z, w.zs, w.red), r, c);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// sums of up to 8 doubles over the CTA; result in every thread
template <int N>
__device__ __forceinline__ void block_sum(Smem& s, double (&v)[N]) {
  const u32 warp = threadIdx.x >> 5, lane = lane_id();
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) s.red[warp][k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double t = 0;
    for (int w = 0; w < kWarps; ++w) t += s.red[w][k];
    v[k] = t;
  }
}

template <typename PX>
__global__ void __launch_bounds__(kThreads)
object_float_kernel(const uint16_t* __restrict__ labels, i64 lab_plane_stride, i64 lab_row_stride,
                    const int32_t* __restrict__ plane_tile, const int32_t* __restrict__ plane_base, int n_planes,
                    int n_objects, int n_total, const PX* __restrict__ pixels, const i64* __restrict__ tile_offset,
                    i64 chan_stride, i64 z_stride, i64 px_row_stride, int Z, int pixel_dtype,
                    const abx_request* __restrict__ requests, int n_requests, const abx_object_rec* __restrict__ recs,
                    ChanStats* __restrict__ out) {
  __shared__ Smem s;
  const u32 lane = lane_id(), warp = threadIdx.x >> 5;
  const double kNaN = __longlong_as_double(0x7FF8000000000000ll);

  for (int obj = blockIdx.x; obj < n_total; obj += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const bool bg = obj >= n_objects;
      const int p = bg ? (obj - n_objects) : find_plane(plane_base, n_planes, obj);
      s.plane = p;
      s.tile = plane_tile[p];
      s.label = bg ? 0u : (u32)(obj - plane_base[p] + 1);
      s.rec = recs[obj];
    }
    __syncthreads();
    const bool is_bg = obj >= n_objects;
    const u32 n = s.rec.n;

    for (int q = 0; q < n_requests; ++q) {
      const abx_request rq = requests[q];
      if (!request_is_float(pixel_dtype, rq.reduction)) continue;  // integer kernels own it (block-uniform)
      const u32 feats = is_bg ? rq.bg_features : rq.features;
      FloatStats* dst = reinterpret_cast<FloatStats*>(out + (i64)obj * n_requests + q);
      FloatStats fs;
      fs.sum = fs.sumsq = fs.css = 0; fs.moi = kNaN;
      fs.top2p5_sum = fs.top5_sum = fs.med_lo = fs.med_hi = fs.vmin = fs.vmax = 0;
      fs.has_nan = 0; fs.pad_ = 0;
      if (n == 0 || (is_bg && feats == 0)) {
        if (threadIdx.x == 0) *dst = fs;
        continue;
      }
      Window<PX> w;
      w.lab = labels + (i64)s.plane * lab_plane_stride;
      w.px = pixels + tile_offset[s.tile] + (i64)rq.channel * chan_stride;
      w.lab_rs = lab_row_stride; w.px_rs = px_row_stride; w.zs = z_stride;
      w.rmin = s.rec.rmin; w.rmax = s.rec.rmax; w.cmin = s.rec.cmin; w.cmax = s.rec.cmax;
      w.label = s.label; w.Z = Z; w.red = rq.reduction;

      // ---- sweep 1: NaN count, raw moments, extrema of the keys ----
      double a[5] = {0, 0, 0, 0, 0};  // nan count, sum, m10, m01, any nonzero
      u64 kmin = ~0ull, kmax = 0;
      sweep(w, [&](double x, u32 r, u32 c) {
        if (x != x) { a[0] += 1.0; return; }
        a[1] += x;
        a[2] += x * (double)(c + 1u);  // 1-based plane coordinates, cell.py:243-250
        a[3] += x * (double)(r + 1u);
        if (x != 0.0) a[4] += 1.0;
        const u64 k = key_of(x);
        kmin = min(kmin, k); kmax = max(kmax, k);
      });
      block_sum<5>(s, a);
      {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          kmin = min(kmin, (u64)__shfl_xor_sync(0xFFFFFFFFu, kmin, o));
          kmax = max(kmax, (u64)__shfl_xor_sync(0xFFFFFFFFu, kmax, o));
        }
        __syncthreads();
        if (lane == 0) { s.redk[warp][0] = kmin; s.redk[warp][1] = kmax; }
        __syncthreads();
        kmin = ~0ull; kmax = 0;
        for (int ww = 0; ww < kWarps; ++ww) { kmin = min(kmin, s.redk[ww][0]); kmax = max(kmax, s.redk[ww][1]); }
      }
      if (a[0] > 0.0) {  // NumPy: every statistic of a sample with a NaN is NaN
        fs.has_nan = 1;
        if (threadIdx.x == 0) *dst = fs;
        continue;
      }
      fs.sum = a[1];
      fs.vmin = value_of(kmin); fs.vmax = value_of(kmax);
      const double mean = a[1] / (double)n;
      const double m00 = a[1], xm = a[2] / m00, ym = a[3] / m00;
      const bool any_nonzero = a[4] > 0.0;

      // ---- sweep 2: centred sums (np.std is two-pass), sum of squares, central moments ----
      double b[4] = {0, 0, 0, 0};
      sweep(w, [&](double x, u32 r, u32 c) {
        const double d = x - mean;
        b[0] += d * d;
        b[1] += x * x;
        const double dc = (double)(c + 1u) - xm, dr = (double)(r + 1u) - ym;
        b[2] += x * (dc * dc);
        b[3] += x * (dr * dr);
      });
      block_sum<4>(s, b);
      fs.css = b[0]; fs.sumsq = b[1];
      if (any_nonzero) {  // cell.py:241-262: Eta20 + Eta02 with Mu00 ** 2.0
        const double p2 = m00 * m00;
        fs.moi = b[2] / p2 + b[3] / p2;
      }

      if (feats & (ABX_F_MEDIAN | ABX_F_TOP2P5 | ABX_F_TOP5)) {
        // ---- radix select of four ranks on d = key - kmin, 8 bits per sweep from the highest differing bit ----
        const u32 k2p5 = (u32)ceil((double)n * 0.025);
        const u32 k5 = min(n, 5u);
        const u32 ranks[4] = {(n - 1) / 2, n / 2, n - k2p5, n - k5};
        const u64 span = kmax - kmin;
        int bits = span ? 64 - __clzll((long long)span) : 0;  // number of significant bits of d
        bits = (bits + 7) & ~7;
        __syncthreads();
        if (threadIdx.x < 4) { s.prefix[threadIdx.x] = 0; s.rank[threadIdx.x] = ranks[threadIdx.x]; }
        __syncthreads();
        for (int shift = bits - 8; shift >= 0; shift -= 8) {
          for (u32 i = threadIdx.x; i < 4u * 256u; i += kThreads) (&s.hist[0][0])[i] = 0;
          __syncthreads();
          const u64 p0 = s.prefix[0], p1 = s.prefix[1], p2 = s.prefix[2], p3 = s.prefix[3];
          const int hs = shift + 8;
          sweep(w, [&](double x, u32, u32) {
            const u64 d = key_of(x) - kmin;
            const u64 hi = hs >= 64 ? 0ull : (d >> hs);
            const u32 dig = (u32)(d >> shift) & 255u;
            if (hi == p0) atomicAdd(&s.hist[0][dig], 1u);
            if (hi == p1) atomicAdd(&s.hist[1][dig], 1u);
            if (hi == p2) atomicAdd(&s.hist[2][dig], 1u);
            if (hi == p3) atomicAdd(&s.hist[3][dig], 1u);
          });
          __syncthreads();
          {  // warp j locates target j in its 256-bin histogram: 8 bins per lane
            const u32 j = warp;
            const u32 want = s.rank[j];
            u32 c8[8], cnt = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { c8[k] = s.hist[j][lane * 8 + k]; cnt += c8[k]; }
            u32 inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
              if (lane >= (u32)o) inc += t;
            }
            const u32 exc = inc - cnt;
            if (want >= exc && want < exc + cnt) {
              u32 acc = exc;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                if (want >= acc && want < acc + c8[k]) {
                  s.prefix[j] = (s.prefix[j] << 8) | (u64)(lane * 8 + k);
                  s.rank[j] = want - acc;
                }
                acc += c8[k];
              }
            }
          }
          __syncthreads();
        }
        const double v0 = value_of(kmin + s.prefix[0]), v1 = value_of(kmin + s.prefix[1]);
        const u64 key2 = kmin + s.prefix[2], key3 = kmin + s.prefix[3];
        const double v2 = value_of(key2), v3 = value_of(key3);
        fs.med_lo = v0; fs.med_hi = v1;
        // ---- sums of the k largest values: everything above the threshold + the tied share ----
        double t[4] = {0, 0, 0, 0};
        sweep(w, [&](double x, u32, u32) {
          const u64 k = key_of(x);
          if (k > key2) { t[0] += x; t[1] += 1.0; }
          if (k > key3) { t[2] += x; t[3] += 1.0; }
        });
        block_sum<4>(s, t);
        fs.top2p5_sum = t[0] + ((double)k2p5 - t[1]) * v2;
        fs.top5_sum = t[2] + ((double)k5 - t[3]) * v3;
      }
      if (threadIdx.x == 0) *dst = fs;
      __syncthreads();
    }
  }
}

}  // namespace

int launch_object_float(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  if (a->n_requests == 0) return ABX_OK;
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  if (n_total == 0) return ABX_OK;
  const bool float_px = a->pixel_dtype == ABX_F32 || a->pixel_dtype == ABX_F64;
  if (!float_px && !((u32)a->request_feature_union & ABX_F_HAS_DIV)) return ABX_OK;  // no `div` request in the plan
  const int grid = n_total < 148 * 8 ? n_total : 148 * 8;
#define ABX_LAUNCH_OF(PX)                                                                                          \
  object_float_kernel<PX><<<grid, kThreads, 0, st>>>(                                                             \
      static_cast<const uint16_t*>(a->labels), a->label_plane_stride, a->label_row_stride, a->plane_tile,         \
      a->plane_base, a->n_planes, a->n_objects, n_total, static_cast<const PX*>(a->pixels),                       \
      reinterpret_cast<const i64*>(a->tile_offset), a->chan_stride, a->z_stride, a->row_stride, a->Z,             \
      a->pixel_dtype, a->requests, a->n_requests, ws.recs, ws.chan)
  if (a->pixel_dtype == ABX_F64) ABX_LAUNCH_OF(double);
  else if (a->pixel_dtype == ABX_F32) ABX_LAUNCH_OF(float);
  else if (a->pixel_dtype == ABX_U16) ABX_LAUNCH_OF(uint16_t);
  else if (a->pixel_dtype == ABX_U8) ABX_LAUNCH_OF(uint8_t);
  else return abx_set_error(ABX_ERR_UNSUPPORTED, "object_float: pixel dtype %d has no kernel", a->pixel_dtype);
#undef ABX_LAUNCH_OF
  return abx_check_cuda(cudaGetLastError(), "object_float");
}
