// Finalisation: raw per-object records -> dense fp64 [objects x columns] table.
//
// One thread per table cell.  The arithmetic follows the reference functions to the
// operation (src/extraction/core/functions/cell.py, trap.py) so that every value that
// NumPy derives from exact integer sums is reproduced bit-for-bit:
//   mean = f64(sum) / f64(n)                     cell.py:43-53
//   median = (lo + hi) / 2                       cell.py:86-96 (np.median on integers)
//   max2p5pc = f64(top sum) / f64(k)             cell.py:99-116
//   max5px_median = (top5 / 5) / median          cell.py:119-144 (NaN for n <= 5 or median 0)
//   centroid = f64(sum(col+1)) / f64(n)          cell.py:282-303
//   std from exact 128-bit integer variance      cell.py:147-157 (<= 1e-12 from NumPy's two-pass)
// The dense table replaces the Python long->wide pivot of extract.py:574-598.
#include "common.cuh"

namespace {

__device__ __forceinline__ double u128_to_double(unsigned __int128 v) {
  return (double)(u64)(v >> 64) * 18446744073709551616.0 + (double)(u64)v;
}

__device__ __forceinline__ double moment_of_inertia(const ChanStats& c) {
  // cell.py:232-265 with coordinates relative to the bbox origin (central moments are
  // translation invariant): mu20 = m20 - m10^2 / m00, eta = mu / m00^2
  if (c.sum == 0) return nan("");
  const unsigned __int128 m00 = c.sum;
  // eta20 + eta02 from one exact integer numerator: m00 * (m20 + m02) - m10^2 - m01^2 (each central moment is
  // non-negative, so the total is; producers may store the two second moments separately or as one sum in m20)
  const unsigned __int128 num = m00 * ((unsigned __int128)c.m20 + c.m02) - (unsigned __int128)c.m10 * c.m10 -
                                (unsigned __int128)c.m01 * c.m01;
  const double d = (double)c.sum;
  return u128_to_double(num) / (d * d * d);
}

// intensity metric of a floating-point request (object_float.cu), NumPy float semantics
__device__ double float_metric(const FloatStats& f, int metric, u32 n_px, int pixel_dtype) {
  const double kNaN = nan("");
  if (n_px == 0)  // absent label: np.sum of an empty selection is 0.0, every other statistic is NaN
    return (metric == ABX_M_TOTAL || metric == ABX_M_TOTAL_SQUARED) ? 0.0 : kNaN;
  if (f.has_nan) return kNaN;  // a NaN poisons every statistic
  const double n = (double)n_px;
  // np.median averages the two middle values in the array's own dtype: float32 pixels round in float32
  const double med = pixel_dtype == ABX_F32 ? (double)(((float)f.med_lo + (float)f.med_hi) / 2.0f)
                                            : (f.med_lo + f.med_hi) / 2.0;
  switch (metric) {
    case ABX_M_MEAN: return f.sum / n;
    case ABX_M_TOTAL: return f.sum;
    case ABX_M_TOTAL_SQUARED: return f.sumsq;
    case ABX_M_STD: return sqrt(f.css / n);
    case ABX_M_MEDIAN:
    case ABX_M_IMBACKGROUND: return med;
    case ABX_M_MAX2P5PC: return f.top2p5_sum / (double)(u32)ceil(n * 0.025);
    case ABX_M_MAX5PX_MEDIAN: {
      if (n_px <= 5u) return kNaN;
      return med == 0.0 ? kNaN : (f.top5_sum / 5.0) / med;
    }
    case ABX_M_BACKGROUND_MAX5: return f.top5_sum / (double)(n_px < 5u ? n_px : 5u);
    case ABX_M_MOMENT_OF_INERTIA: return f.moi;
    case ABX_M_MAX: return f.vmax;
    case ABX_M_MIN: return f.vmin;
    default: return kNaN;  // ratio
  }
}

// CellProfiler's order-statistic rule (MeasureObjectIntensity): q = n f, i = floor(q); v[i] (1 - frac) + v[i + 1] frac when
// i < n - 1, else v[i].  k = 4 f in {1, 2, 3}; (lo, hi) = v[i], v[min(i + 1, n - 1)]; half: the values are lo + 1/2, hi + 1/2.
__device__ __forceinline__ double cp_quantile(u32 n, u32 k, double lo, double hi) {
  const u32 i = (n * k) >> 2;
  const double frac = (double)((n * k) & 3u) * 0.25;
  return i < n - 1u ? lo * (1.0 - frac) + hi * frac : lo;
}

// Two-image features from the exact sums of object_pair.cu (CellProfiler MeasureColocalization, one object at a time)
__device__ double pair_metric(const PairStats& p, const ChanStats& ca, const ChanStats& cb, u32 n_px, int metric) {
  const double kNaN = nan("");
  if (n_px == 0 || (p.flags & 1u)) return kNaN;
  if (metric == ABX_M_CO_PEARSON) {
    // n sum(xy) - sum(x) sum(y) over sqrt of the two variances' numerators, each an exact 128-bit integer
    const unsigned __int128 n = n_px;
    const unsigned __int128 vx = n * ca.sumsq - (unsigned __int128)ca.sum * ca.sum;
    const unsigned __int128 vy = n * cb.sumsq - (unsigned __int128)cb.sum * cb.sum;
    if (vx == 0 || vy == 0) return kNaN;  // a constant image: 0 / 0
    const unsigned __int128 pos = n * p.sxy, neg = (unsigned __int128)ca.sum * cb.sum;
    const double cov = pos >= neg ? u128_to_double(pos - neg) : -u128_to_double(neg - pos);
    return cov / (sqrt(u128_to_double(vx)) * sqrt(u128_to_double(vy)));
  }
  if (p.n_both == 0) return 0.0;  // no pixel above both thresholds
  switch (metric) {
    case ABX_M_CO_MANDERS_1: return (double)p.cx / (double)p.tot_x;
    case ABX_M_CO_MANDERS_2: return (double)p.cy / (double)p.tot_y;
    case ABX_M_CO_RWC_1: return (double)p.wx / (double)p.big_r / (double)p.tot_x;
    case ABX_M_CO_RWC_2: return (double)p.wy / (double)p.big_r / (double)p.tot_y;
    case ABX_M_CO_OVERLAP: return (double)p.cxy / sqrt((double)p.cxx * (double)p.cyy);
    case ABX_M_CO_K_1: return (double)p.cxy / (double)p.cxx;
    case ABX_M_CO_K_2: return (double)p.cxy / (double)p.cyy;
    default: return kNaN;
  }
}

__device__ double finalize_cell(const abx_object_rec& r, int obj, int col, const abx_object_rec* __restrict__ recs,
                                const ChanStats* __restrict__ chan, const ShapeStats* __restrict__ shape,
                                const MaskMoments* __restrict__ mom, const PairStats* __restrict__ pair_stats,
                                const abx_pair* __restrict__ pairs, int n_pairs,
                                const int32_t* __restrict__ plane_base, int n_planes, int n_objects,
                                const abx_request* __restrict__ requests, int n_requests,
                                const abx_column* __restrict__ columns, int pixel_dtype,
                                const ChanStats* __restrict__ own /* the records of this object, staged (or chan + obj * n_requests) */) {
  const abx_column cd = columns[col];
  const double n = (double)r.n;
  const double kNaN = nan("");
  double v = kNaN;
  if (cd.metric >= ABX_M_CO_PEARSON) {
    const abx_pair pr = pairs[cd.request];
    v = pair_metric(pair_stats[(i64)obj * n_pairs + cd.request], own[pr.request_a], own[pr.request_b], r.n, cd.metric);
  } else if (cd.metric >= ABX_M_CP_BBOX_AREA) {
    // ---- cp_measure `sizeshape` subset (label plane only) ----
    if (r.n) {
      const double hh = (double)(r.rmax - r.rmin + 1u), ww = (double)(r.cmax - r.cmin + 1u);
      switch (cd.metric) {
        case ABX_M_CP_BBOX_AREA: v = hh * ww; break;
        case ABX_M_CP_BBOX_MAX_X: v = (double)r.cmax + 1.0; break;  // exclusive, like skimage's bbox
        case ABX_M_CP_BBOX_MAX_Y: v = (double)r.rmax + 1.0; break;
        case ABX_M_CP_CENTER_X: v = (double)(r.sum_col - r.n) / n; break;  // 0-based centroid
        case ABX_M_CP_CENTER_Y: v = (double)(r.sum_row - r.n) / n; break;
        case ABX_M_CP_EQUIVALENT_DIAMETER: v = sqrt(4.0 * n / 3.141592653589793); break;
        case ABX_M_CP_EXTENT: v = n / (hh * ww); break;
        case ABX_M_CP_MAXIMUM_RADIUS: v = sqrt((double)shape[obj].max_nn2); break;
        case ABX_M_CP_MEAN_RADIUS: v = shape[obj].sum_nn / n; break;
        default: {
          // second central moments of the coordinates from the raw sums relative to the bbox origin
          const MaskMoments m = mom[obj];
          const double sr = (double)(r.sum_row - (u64)r.n * (r.rmin + 1u)), sc = (double)(r.sum_col - (u64)r.n * (r.cmin + 1u));
          const double mu_rr = ((double)m.s_rr - sr * sr / n) / n, mu_cc = ((double)m.s_cc - sc * sc / n) / n;
          const double mu_rc = ((double)m.s_rc - sr * sc / n) / n;
          const double half_tr = (mu_rr + mu_cc) / 2.0, dd = (mu_rr - mu_cc) / 2.0;
          const double root = sqrt(dd * dd + mu_rc * mu_rc);
          const double l1 = half_tr + root, l2 = fmax(half_tr - root, 0.0);
          if (cd.metric == ABX_M_CP_MAJOR_AXIS_LENGTH) v = 4.0 * sqrt(l1);
          else if (cd.metric == ABX_M_CP_MINOR_AXIS_LENGTH) v = 4.0 * sqrt(l2);
          else v = l1 > 0.0 ? sqrt(1.0 - l2 / l1) : 0.0;  // ABX_M_CP_ECCENTRICITY
        } break;
      }
    }
  } else if (cd.metric < 16) {
    double minor = 0, major = 0;
    if (cd.metric == ABX_M_ECCENTRICITY || cd.metric == ABX_M_VOLUME || cd.metric == ABX_M_MINOR_AXIS ||
        cd.metric == ABX_M_MAJOR_AXIS) {
      const ShapeStats& s = shape[obj];
      minor = rint(sqrt((double)s.max_nn2));                       // np.round: half to even
      major = rint(sqrt((double)s.max_dn2) + s.sum_top / 2.0);
    }
    switch (cd.metric) {
      case ABX_M_AREA: v = n; break;
      // (an absent label is 0 / 0 = NaN in NumPy: said outright, the special-value subroutine of the fp64 division is long)
      case ABX_M_CENTROID_X: if (r.n) v = (double)r.sum_col / n; break;
      case ABX_M_CENTROID_Y: if (r.n) v = (double)r.sum_row / n; break;
      case ABX_M_SPHERICAL_VOLUME: {
        const double rad = sqrt(n / 3.141592653589793);
        v = (4.0 * 3.141592653589793 * (rad * rad * rad)) / 3.0;
      } break;
      // (a round cell, major == minor, is sqrt(0) / major = +0: said outright for the same reason; 0 / 0 stays NaN)
      case ABX_M_ECCENTRICITY:
        if (major > 0.0) v = major == minor ? 0.0 : sqrt(major * major - minor * minor) / major;
        break;
      case ABX_M_VOLUME: v = (4.0 * 3.141592653589793 * (minor * minor) * major) / 3.0; break;
      case ABX_M_CONICAL_VOLUME: v = 4.0 * shape[obj].sum_nn; break;
      case ABX_M_MINOR_AXIS: v = minor; break;
      case ABX_M_MAJOR_AXIS: v = major; break;
      case ABX_M_BBOX_RMIN: v = r.n ? (double)r.rmin : kNaN; break;
      case ABX_M_BBOX_RMAX: v = r.n ? (double)r.rmax : kNaN; break;
      case ABX_M_BBOX_CMIN: v = r.n ? (double)r.cmin : kNaN; break;
      case ABX_M_BBOX_CMAX: v = r.n ? (double)r.cmax : kNaN; break;
      default: break;
    }
  } else if (request_is_float(pixel_dtype, requests[cd.request].reduction)) {
    const bool bg = cd.metric == ABX_M_IMBACKGROUND || cd.metric == ABX_M_BACKGROUND_MAX5;
    const int row = bg ? n_objects + find_plane(plane_base, n_planes, obj) : obj;
    v = float_metric(*reinterpret_cast<const FloatStats*>(bg ? &chan[(i64)row * n_requests + cd.request] : &own[cd.request]), cd.metric,
                     recs[row].n, pixel_dtype);
  } else if (cd.metric == ABX_M_IMBACKGROUND || cd.metric == ABX_M_BACKGROUND_MAX5) {
    const int p = find_plane(plane_base, n_planes, obj);
    const u32 nb = recs[n_objects + p].n;
    const ChanStats& c = chan[(i64)(n_objects + p) * n_requests + cd.request];  // only the fields used are loaded
    if (nb) {
      if (cd.metric == ABX_M_IMBACKGROUND) v = ((double)c.med_lo + (double)c.med_hi) / 2.0;
      else v = (double)c.top5_sum / (double)(nb < 5u ? nb : 5u);  // np.mean(np.sort(bg)[-5:])
    }
  } else {
    const ChanStats& c = own[cd.request];
    const bool add = requests[cd.request].reduction == ABX_RED_ADD;
    switch (cd.metric) {
      case ABX_M_MEAN: v = r.n ? (double)c.sum / n : kNaN; break;
      case ABX_M_TOTAL: v = (double)c.sum; break;
      case ABX_M_TOTAL_SQUARED: v = (double)(add ? c.sumsq : c.wrapsq); break;
      case ABX_M_STD:
        if (r.n) {
          const unsigned __int128 num = (unsigned __int128)r.n * c.sumsq - (unsigned __int128)c.sum * c.sum;
          v = num == 0 ? 0.0 : sqrt(u128_to_double(num) / (n * n));  // (a constant cell: +0)
        }
        break;
      case ABX_M_MEDIAN: if (r.n) v = ((double)c.med_lo + (double)c.med_hi) / 2.0; break;
      case ABX_M_MAX2P5PC:
        if (r.n) v = (double)c.top2p5_sum / (double)(u32)ceil(n * 0.025);
        break;
      case ABX_M_MAX5PX_MEDIAN:
        if (r.n > 5u) {
          const double med = ((double)c.med_lo + (double)c.med_hi) / 2.0;
          if (med != 0.0) v = ((double)c.top5_sum / 5.0) / med;
        }
        break;
      case ABX_M_MOMENT_OF_INERTIA: if (r.n) v = moment_of_inertia(c); break;
      case ABX_M_MAX: if (r.n) v = (double)c.vmax; break;
      case ABX_M_MIN: if (r.n) v = (double)c.vmin; break;
      // ---- cp_measure `intensity` (object_sweep.cu fills q / mad / maxpos) ----
      case ABX_M_CP_LOWER_QUARTILE: if (r.n) v = cp_quantile(r.n, 1u, (double)c.q[0], (double)c.q[1]); break;
      case ABX_M_CP_MEDIAN: if (r.n) v = cp_quantile(r.n, 2u, (double)c.q[2], (double)c.q[3]); break;
      case ABX_M_CP_UPPER_QUARTILE: if (r.n) v = cp_quantile(r.n, 3u, (double)c.q[4], (double)c.q[5]); break;
      case ABX_M_CP_MAD:
        if (r.n) {
          const double half = (c.mad_hi >> 31) ? 0.5 : 0.0;
          v = cp_quantile(r.n, 2u, (double)c.mad_lo + half, (double)(c.mad_hi & 0x7FFFFFFFu) + half);
        }
        break;
      case ABX_M_CP_CENTER_MASS_X: if (r.n && c.sum) v = (double)r.cmin + (double)c.m10 / (double)c.sum; break;
      case ABX_M_CP_CENTER_MASS_Y: if (r.n && c.sum) v = (double)r.rmin + (double)c.m01 / (double)c.sum; break;
      case ABX_M_CP_MASS_DISPLACEMENT:
        if (r.n && c.sum) {
          const double dx = (double)r.cmin + (double)c.m10 / (double)c.sum - (double)(r.sum_col - r.n) / n;
          const double dy = (double)r.rmin + (double)c.m01 / (double)c.sum - (double)(r.sum_row - r.n) / n;
          v = sqrt(dx * dx + dy * dy);
        }
        break;
      case ABX_M_CP_MAX_POS_X: if (r.n) v = (double)(r.cmin + (c.maxpos & 0xFFFFu)); break;
      case ABX_M_CP_MAX_POS_Y: if (r.n) v = (double)(r.rmin + (c.maxpos >> 16)); break;
      case ABX_M_CP_ZERO: if (r.n) v = 0.0; break;
      case ABX_M_CP_CENTER_MASS_Z: if (r.n && c.sum) v = 0.0; break;
      case ABX_M_RATIO: default: break;  // cell.py:268-279: NaN for any 2-D image
    }
  }
  return v;
}

// Thread mapping: a CTA of 256 threads takes 32 objects; lane <-> object, warp w walks the columns w, w + 8, ... — so a
// warp evaluates ONE metric for 32 objects (no divergence in the switch below; one column per lane cost every warp the
// union of all metric bodies).  Results go through a shared-memory tile so that the table rows are written coalesced.
constexpr int kFinObjects = 32, kFinWarps = 8, kFinColChunk = 64;
constexpr int kFinStageRequests = 7;  // up to 7 requests (28 KB next to the 16.6 KB tile: under the 48 KB a launch gets without opting in)

__global__ void __launch_bounds__(kFinWarps * 32, 3)
finalize_kernel(const abx_object_rec* __restrict__ recs, const ChanStats* __restrict__ chan,
                                const ShapeStats* __restrict__ shape, const MaskMoments* __restrict__ mom,
                                const PairStats* __restrict__ pair_stats, const abx_pair* __restrict__ pairs, int n_pairs,
                                const int32_t* __restrict__ plane_base,
                                int n_planes, int n_objects, const abx_request* __restrict__ requests,
                                int n_requests, const abx_column* __restrict__ columns, int n_columns,
                                int pixel_dtype, double* __restrict__ table, const u32* __restrict__ err,
                                u32* __restrict__ status) {
  __shared__ double tile[kFinObjects][kFinColChunk + 1];
  // the ChanStats of the CTA's 32 objects are contiguous in memory (object-major): staged with coalesced 16-byte loads,
  // all independent — one memory latency instead of one per column (lane <-> object reads them 640 bytes apart)
  extern __shared__ __align__(16) unsigned char fin_dyn[];
  const int obj_lo = blockIdx.x * kFinObjects;
  const bool staged = n_requests > 0 && n_requests <= kFinStageRequests;
  if (staged) {
    const int n_own = min(kFinObjects, n_objects - obj_lo);
    const uint4* src = reinterpret_cast<const uint4*>(chan + (i64)obj_lo * n_requests);
    uint4* dst = reinterpret_cast<uint4*>(fin_dyn);
    const int n16 = n_own * n_requests * (int)(sizeof(ChanStats) / 16);
    for (int i = threadIdx.x; i < n16; i += kFinWarps * 32) dst[i] = src[i];
    __syncthreads();
  }
  if (status != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *status = *err;  // every other kernel of the call has finished
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int obj0 = blockIdx.x * kFinObjects;
  const int obj = obj0 + lane;
  const bool live = obj < n_objects;
  const abx_object_rec r = recs[live ? obj : 0];
  for (int cbase = 0; cbase < n_columns; cbase += kFinColChunk) {
    const int ncol = min(kFinColChunk, n_columns - cbase);
    for (int cl = warp; cl < ncol; cl += kFinWarps) {
      const int col = cbase + cl;
      const ChanStats* own = staged ? reinterpret_cast<const ChanStats*>(fin_dyn) + lane * n_requests
                                    : chan + (i64)(live ? obj : 0) * n_requests;
      tile[lane][cl] = live ? finalize_cell(r, obj, col, recs, chan, shape, mom, pair_stats, pairs, n_pairs, plane_base, n_planes, n_objects, requests,
                                            n_requests, columns, pixel_dtype, own)
                            : 0.0;
    }
    __syncthreads();
    // coalesced write-out: consecutive threads -> consecutive columns of one object
    for (int e = threadIdx.x; e < kFinObjects * ncol; e += kFinWarps * 32) {
      const int o = e / ncol, c = e - o * ncol;
      if (obj0 + o < n_objects) table[(i64)(obj0 + o) * n_columns + cbase + c] = tile[o][c];
    }
    __syncthreads();
  }
}

}  // namespace

int launch_finalize(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  const i64 cells = (i64)a->n_objects * a->n_columns;
  if (cells == 0) {
    if (a->status) return abx_check_cuda(cudaMemcpyAsync(a->status, ws.err, sizeof(u32), cudaMemcpyDeviceToDevice, st), "status copy");
    return ABX_OK;
  }
  const unsigned blocks = (unsigned)((a->n_objects + kFinObjects - 1) / kFinObjects);
  const size_t dyn = (a->n_requests > 0 && a->n_requests <= kFinStageRequests)
                         ? (size_t)kFinObjects * a->n_requests * sizeof(ChanStats) : 0;  // (+ 16.6 KB static: under 48 KB)
  finalize_kernel<<<blocks, kFinWarps * 32, dyn, st>>>(ws.recs, ws.chan, ws.shape, ws.mom, ws.pairs, a->pairs, a->n_pairs, a->plane_base, a->n_planes,
                                                   a->n_objects, a->requests, a->n_requests, a->columns,
                                                   a->n_columns, a->pixel_dtype, a->table, ws.err, a->status);
  return abx_check_cuda(cudaGetLastError(), "finalize");
}
