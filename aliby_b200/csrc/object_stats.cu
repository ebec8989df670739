// Per-object intensity statistics: one CTA per object, one pass per (channel, Z-reduction)
// request over the object's bounding-box window.
//
// Replaces the |objects| x |instructions| loop of src/extraction/extract.py:346-359: where the
// reference gathers `reduce_z(pixels[tile, ch])[mask]` from the whole plane for every call
// (extract.py:105-107, cell.py:43-157), this kernel touches only the bbox window found by the
// label scan, fuses the Z reduction (distributors.py:19-21) into the load, fuses the tile crop
// (tiler.py:309-366) through the per-tile element offset, and derives every order statistic
// (median, top-2.5 %, top-5; background median / top-5 for label 0, trap.py:6-43) from one
// range-adaptive histogram instead of a sort:
//
//   sweep 1  window -> sum, sum of squares, min, max, first/second moments; the object's values
//            are compacted into shared memory (if they fit) so later sweeps never touch labels
//   sweep 2  histogram of (x - min) >> s with s chosen so that the range fits 1024 bins
//            (s = 0, i.e. exact, for every object whose value range is < 1024)
//   refine   only if s > 0: 8 more bits per sweep inside the (up to four) target bins
//   top sums exact, from the prefix sums of the histogram (s = 0) or one more sweep (s > 0)
//
// All integer work is exact (u64 sums, integer order statistics): results are bit-identical
// to NumPy's on integer pixels.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kCap = 4096;   // values of one object kept in shared memory
constexpr int kBins = 1024;  // level-0 histogram bins (4 x 256 during refinement)
constexpr int kRed = 9;      // u64 quantities reduced per request

struct Smem {
  u32 vals[kCap];
  u32 hist[kBins];
  u64 red[kWarps][kRed];
  u32 redmin[kWarps], redmax[kWarps];
  u32 n_pushed;
  int plane, tile;
  u32 label;
  abx_object_rec rec;
  u32 t_key[4], t_rank[4], t_below_cnt[4];
  u64 t_below_sum[4];
  u32 r_key[4], r_rank[4], r_cnt[4];
  u64 r_sum[4];
  u64 redpos[kWarps];  // (max value << 32) | ~position: the first maximum in row-major order
};
static_assert(kWarps == 4, "one warp per selection target during refinement");

template <typename PX>
struct Window {
  const uint16_t* lab;  // plane base
  const PX* px;         // (tile, channel, z = 0, row 0, col 0)
  i64 lab_row_stride, px_row_stride, z_stride;
  u32 rmin, rmax, cmin, cmax, label;
  int Z, red;
};

template <typename PX>
__device__ __forceinline__ u32 load_reduced(const PX* __restrict__ p, int Z, i64 z_stride, int red) {
  u32 x = (u32)__ldg(p);
  if (red == ABX_RED_MAX) {
    for (int z = 1; z < Z; ++z) x = max(x, (u32)__ldg(p + (i64)z * z_stride));
  } else {
    for (int z = 1; z < Z; ++z) x += (u32)__ldg(p + (i64)z * z_stride);
  }
  return x;
}

// Warp-uniform walk over the window: f(hit, x, r, c) is called by all 32 lanes.
template <typename PX, class F>
__device__ __forceinline__ void sweep_window(const Window<PX>& w, F&& f) {
  const u32 warp = threadIdx.x >> 5, lane = lane_id();
  for (u32 r = w.rmin + warp; r <= w.rmax; r += kWarps) {
    const uint16_t* lrow = w.lab + (i64)r * w.lab_row_stride;
    const PX* prow = w.px + (i64)r * w.px_row_stride;
    for (u32 c0 = w.cmin; c0 <= w.cmax; c0 += 32) {
      const u32 c = c0 + lane;
      const bool hit = (c <= w.cmax) && ((u32)__ldg(lrow + c) == w.label);
      u32 x = 0;
      if (hit) x = load_reduced(prow + c, w.Z, w.z_stride, w.red);
      f(hit, x, r, c);
    }
  }
}

// f(x) for every value of the object: from shared memory when compacted, else from the window.
template <typename PX, class F>
__device__ __forceinline__ void for_each_value(const Window<PX>& w, const Smem& s, bool compact, u32 n, F&& f) {
  if (compact) {
    for (u32 i = threadIdx.x; i < n; i += kThreads) f(s.vals[i]);
  } else {
    sweep_window(w, [&](bool hit, u32 x, u32, u32) { if (hit) f(x); });
  }
}

__device__ __forceinline__ u64 warp_sum(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// Locate rank `t` inside hist[0, nb): one warp, `per` consecutive bins per lane.
// value_of(b) gives the pixel value of bin b for the weighted prefix (only used when exact).
__device__ __forceinline__ void warp_find_rank(const u32* hist, u32 nb, u32 vbase, const u32* ranks, int n_ranks,
                                                u32* out_key, u32* out_rank, u32* out_below_cnt, u64* out_below_sum) {
  const u32 lane = lane_id();
  const u32 per = (nb + 31) / 32;
  const u32 b0 = lane * per;
  const u32 b1 = min(b0 + per, nb);
  u32 cnt = 0;
  u64 wsum = 0;
  for (u32 b = b0; b < b1; ++b) {
    const u32 c = hist[b];
    cnt += c;
    wsum += (u64)c * (u64)(vbase + b);
  }
  u32 icnt = cnt;
  u64 iw = wsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 c = __shfl_up_sync(0xFFFFFFFFu, icnt, o);
    const u64 q = __shfl_up_sync(0xFFFFFFFFu, iw, o);
    if (lane >= (u32)o) { icnt += c; iw += q; }
  }
  const u32 ecnt = icnt - cnt;
  const u64 ew = iw - wsum;
  for (int j = 0; j < n_ranks; ++j) {
    const u32 t = ranks[j];
    if (t >= ecnt && t < ecnt + cnt) {
      u32 acc = ecnt;
      u64 ws = ew;
      for (u32 b = b0; b < b1; ++b) {
        const u32 c = hist[b];
        if (t < acc + c) {
          out_key[j] = b; out_rank[j] = t - acc; out_below_cnt[j] = acc; out_below_sum[j] = ws;
          break;
        }
        acc += c;
        ws += (u64)c * (u64)(vbase + b);
      }
    }
  }
}

template <typename PX>
__global__ void __launch_bounds__(kThreads)
object_stats_kernel(const uint16_t* __restrict__ labels, i64 lab_plane_stride, i64 lab_row_stride,
                    const int32_t* __restrict__ plane_tile, const int32_t* __restrict__ plane_base, int n_planes,
                    int n_objects, int n_total, const PX* __restrict__ pixels,
                    const i64* __restrict__ tile_offset, i64 chan_stride, i64 z_stride, i64 px_row_stride, int Z,
                    const abx_request* __restrict__ requests, int n_requests,
                    const abx_object_rec* __restrict__ recs, ChanStats* __restrict__ out,
                    const int* __restrict__ work_list, const u32* __restrict__ work_count, int big_background,
                    int all_objects, u32* __restrict__ err) {
  __shared__ Smem s;
  const u32 lane = lane_id(), warp = threadIdx.x >> 5;
  constexpr u32 kWrapMask = (sizeof(PX) == 1) ? 0xFFu : 0xFFFFu;

  // objects handed over by the plan / warp-per-object kernel; or EVERY object: cp_measure rank statistics in a layout
  // the sweep kernel cannot address (this kernel reads with plain loads and recomputes the whole record)
  const u32 n_work = all_objects ? (u32)n_total : *work_count;
  if (all_objects && blockIdx.x == 0 && threadIdx.x == 0)
    for (int q = 0; q < n_requests; ++q)  // the float kernel has no rank statistics: status bit 1
      if ((requests[q].features & (ABX_F_CPQ | ABX_F_CPMAD)) && requests[q].reduction == ABX_RED_DIV) atomicOr(err, 2u);
  for (u32 wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const int obj = all_objects ? (int)wi : work_list[wi];
    __syncthreads();  // previous object's smem is dead
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
    const bool compact = n <= (u32)kCap;

    for (int q = 0; q < n_requests; ++q) {
      const abx_request rq = requests[q];
      if (rq.reduction == ABX_RED_DIV) continue;  // floating-point request: object_float.cu (block-uniform)
      if (is_bg && big_background && !(rq.reduction == ABX_RED_ADD && Z > 1)) continue;  // background.cu
      const u32 feats = is_bg ? rq.bg_features : rq.features;
      ChanStats* dst = out + (i64)obj * n_requests + q;
      if (n == 0 || (is_bg && feats == 0)) {  // block-uniform
        if (threadIdx.x == 0) {
          ChanStats z;
          z.sum = z.sumsq = z.wrapsq = z.m10 = z.m01 = z.m20 = z.m02 = z.top2p5_sum = z.top5_sum = 0;
          z.vmin = z.vmax = z.med_lo = z.med_hi = 0;
          *dst = z;
        }
        continue;
      }
      Window<PX> w;
      w.lab = labels + (i64)s.plane * lab_plane_stride;
      w.px = pixels + tile_offset[s.tile] + (i64)rq.channel * chan_stride;
      w.lab_row_stride = lab_row_stride; w.px_row_stride = px_row_stride; w.z_stride = z_stride;
      w.rmin = s.rec.rmin; w.rmax = s.rec.rmax; w.cmin = s.rec.cmin; w.cmax = s.rec.cmax;
      w.label = s.label; w.Z = Z; w.red = rq.reduction;
      const bool wrap16 = (rq.reduction == ABX_RED_MAX);  // add -> NumPy promotes to uint64: no wrap
      const bool want_moi = (feats & ABX_F_MOI) != 0;

      if (threadIdx.x == 0) s.n_pushed = 0;
      __syncthreads();

      // ---- sweep 1: moments, extrema, compaction ----
      u64 a_sum = 0, a_sq = 0, a_wrap = 0, a_m10 = 0, a_m01 = 0, a_m20 = 0, a_m02 = 0;
      u32 a_min = 0xFFFFFFFFu, a_max = 0;
      u64 a_pos = 0;  // (value << 32) | ~((row << 16) | col): its maximum is the first maximum in row-major order
      sweep_window(w, [&](bool hit, u32 x, u32 r, u32 c) {
        if (hit) {
          a_pos = max(a_pos, ((u64)x << 32) | (u64)(0xFFFFFFFFu - (((r - w.rmin) << 16) | (c - w.cmin))));
          a_sum += x;
          const u64 xx = (u64)x * (u64)x;
          a_sq += xx;
          a_wrap += wrap16 ? (u64)((u32)xx & kWrapMask) : xx;
          a_min = min(a_min, x);
          a_max = max(a_max, x);
          if (want_moi) {
            const u64 rc = c - w.cmin, rr = r - w.rmin;
            a_m10 += (u64)x * rc; a_m01 += (u64)x * rr;
            a_m20 += (u64)x * rc * rc; a_m02 += (u64)x * rr * rr;
          }
        }
        if (compact) {
          const u32 m = __ballot_sync(0xFFFFFFFFu, hit);
          if (m) {
            u32 base = 0;
            if (lane == 0) base = atomicAdd(&s.n_pushed, (u32)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (hit) s.vals[base + __popc(m & ((1u << lane) - 1u))] = x;
          }
        }
      });
      {
        u64 v[kRed] = {a_sum, a_sq, a_wrap, a_m10, a_m01, a_m20, a_m02, 0, 0};
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] = warp_sum(v[k]);
        a_min = __reduce_min_sync(0xFFFFFFFFu, a_min);
        a_max = __reduce_max_sync(0xFFFFFFFFu, a_max);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a_pos = max(a_pos, __shfl_xor_sync(0xFFFFFFFFu, a_pos, o));
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 7; ++k) s.red[warp][k] = v[k];
          s.redmin[warp] = a_min; s.redmax[warp] = a_max;
          s.redpos[warp] = a_pos;
        }
      }
      __syncthreads();
      u64 pos = 0;
      for (int ww = 0; ww < kWarps; ++ww) pos = max(pos, s.redpos[ww]);
      const u32 first_max_pos = 0xFFFFFFFFu - (u32)pos;
      u64 tot[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) { tot[k] = 0; for (int ww = 0; ww < kWarps; ++ww) tot[k] += s.red[ww][k]; }
      u32 vmin = 0xFFFFFFFFu, vmax = 0;
      for (int ww = 0; ww < kWarps; ++ww) { vmin = min(vmin, s.redmin[ww]); vmax = max(vmax, s.redmax[ww]); }

      ChanStats cs;
      cs.sum = tot[0]; cs.sumsq = tot[1]; cs.wrapsq = tot[2];
      cs.m10 = tot[3]; cs.m01 = tot[4]; cs.m20 = tot[5]; cs.m02 = tot[6];
      cs.vmin = vmin; cs.vmax = vmax; cs.med_lo = cs.med_hi = 0; cs.top2p5_sum = cs.top5_sum = 0;
      cs.q[0] = cs.q[1] = cs.q[2] = cs.q[3] = cs.q[4] = cs.q[5] = 0; cs.mad_lo = cs.mad_hi = 0; cs.maxpos = 0; cs.pad_ = 0;

      // ---- order statistics: range-adaptive histogram of xf(x) in [lo, hi], four ranks at a time ----
      //   histogram of (xf(x) - lo) >> s with s chosen so that the range fits 1024 bins (s = 0: exact), then 8 more bits
      //   per sweep inside the (up to four) target bins.  Leaves the exact-level prefix data in s.t_*; returns s.
      auto select4 = [&](const u32 (&rk)[4], auto xf, u32 lo, u32 hi, u32 (&value)[4]) -> int {
        const u32 range = hi - lo;
        int s0 = 0;
        while ((range >> s0) >= (u32)kBins) ++s0;
        const u32 nb = (range >> s0) + 1;
        __syncthreads();  // the histogram and the targets of an earlier selection have been consumed
        for (u32 b = threadIdx.x; b < nb; b += kThreads) s.hist[b] = 0;
        __syncthreads();
        for_each_value(w, s, compact, n, [&](u32 x) { atomicAdd(&s.hist[(xf(x) - lo) >> s0], 1u); });
        __syncthreads();
        if (warp == 0) warp_find_rank(s.hist, nb, lo, rk, 4, s.t_key, s.t_rank, s.t_below_cnt, s.t_below_sum);
        __syncthreads();
        int cur = s0;
        while (cur > 0) {
          const int nxt = cur > 8 ? cur - 8 : 0;
          const u32 nsub = 1u << (cur - nxt);
          const u32 k0 = s.t_key[0], k1 = s.t_key[1], k2 = s.t_key[2], k3 = s.t_key[3];
          __syncthreads();  // everyone has read the keys and finished with hist
          for (u32 b = threadIdx.x; b < 4u * 256u; b += kThreads) s.hist[b] = 0;
          __syncthreads();
          for_each_value(w, s, compact, n, [&](u32 x) {
            const u32 d = xf(x) - lo;
            const u32 hi_bits = d >> cur;
            const u32 sb = (d >> nxt) & (nsub - 1u);
            if (hi_bits == k0) atomicAdd(&s.hist[sb], 1u);
            if (hi_bits == k1) atomicAdd(&s.hist[256 + sb], 1u);
            if (hi_bits == k2) atomicAdd(&s.hist[512 + sb], 1u);
            if (hi_bits == k3) atomicAdd(&s.hist[768 + sb], 1u);
          });
          __syncthreads();
          {  // warp j refines target j (kWarps == 4)
            const u32 j = warp;
            const u32 want = s.t_rank[j];
            const u32 oldkey = s.t_key[j];
            warp_find_rank(s.hist + 256 * j, nsub, 0u, &want, 1, &s.r_key[j], &s.r_rank[j], &s.r_cnt[j], &s.r_sum[j]);
            __syncwarp();
            if (lane == 0) {
              s.t_key[j] = (oldkey << (cur - nxt)) | s.r_key[j];
              s.t_rank[j] = s.r_rank[j];
            }
          }
          __syncthreads();
          cur = nxt;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) value[j] = lo + s.t_key[j];
        return s0;
      };
      const auto identity = [](u32 x) { return x; };

      if (feats & (ABX_F_MEDIAN | ABX_F_TOP2P5 | ABX_F_TOP5)) {
        const u32 k2p5 = (u32)ceil((double)n * 0.025);  // int(np.ceil(n * 0.025)), cell.py:110-111
        const u32 k5 = min(n, 5u);
        const u32 ranks[4] = {(n - 1) / 2, n / 2, n - k2p5, n - k5};
        u32 value[4];
        u64 below_sum[2];  // sum of the values ranked below targets 2 and 3
        const int s0 = select4(ranks, identity, vmin, vmax, value);
        if (s0 == 0) {
          below_sum[0] = s.t_below_sum[2] + (u64)s.t_rank[2] * value[2];
          below_sum[1] = s.t_below_sum[3] + (u64)s.t_rank[3] * value[3];
        } else {
          // ---- exact sums below the two top-k thresholds ----
          u64 sb2 = 0, sb3 = 0;
          u32 cb2 = 0, cb3 = 0;
          if (feats & (ABX_F_TOP2P5 | ABX_F_TOP5)) {
            const u32 v2 = value[2], v3 = value[3];
            for_each_value(w, s, compact, n, [&](u32 x) {
              if (x < v2) { sb2 += x; ++cb2; }
              if (x < v3) { sb3 += x; ++cb3; }
            });
            sb2 = warp_sum(sb2); sb3 = warp_sum(sb3);
            cb2 = __reduce_add_sync(0xFFFFFFFFu, cb2); cb3 = __reduce_add_sync(0xFFFFFFFFu, cb3);
            __syncthreads();  // red[] from sweep 1 has been consumed by everyone
            if (lane == 0) { s.red[warp][0] = sb2; s.red[warp][1] = sb3; s.red[warp][2] = cb2; s.red[warp][3] = cb3; }
            __syncthreads();
            sb2 = sb3 = 0; u64 c2 = 0, c3 = 0;
            for (int ww = 0; ww < kWarps; ++ww) { sb2 += s.red[ww][0]; sb3 += s.red[ww][1]; c2 += s.red[ww][2]; c3 += s.red[ww][3]; }
            below_sum[0] = sb2 + ((u64)ranks[2] - c2) * (u64)v2;
            below_sum[1] = sb3 + ((u64)ranks[3] - c3) * (u64)v3;
          } else {
            below_sum[0] = below_sum[1] = 0;
          }
        }
        cs.med_lo = value[0]; cs.med_hi = value[1];
        cs.top2p5_sum = cs.sum - below_sum[0];
        cs.top5_sum = cs.sum - below_sum[1];
      }
      if (feats & (ABX_F_CPQ | ABX_F_CPMAD)) {
        // ---- cp_measure `intensity` (the same records object_sweep.cu writes for window-sized objects): the order
        // statistics i = floor(n f) and i + 1 for f = 1/4, 1/2, 3/4; the same pair of floor(|2 v - 2 median| / 2); the
        // position of the first maximum ----
        const u32 last = n - 1u;
        const u32 i1 = n >> 2, i2 = n >> 1, i3 = (u32)((3ull * n) >> 2);
        const u32 rb1[4] = {i1, min(i1 + 1u, last), i2, min(i2 + 1u, last)};
        const u32 rb2[4] = {i3, min(i3 + 1u, last), i3, min(i3 + 1u, last)};
        u32 v1[4], v2[4];
        select4(rb1, identity, vmin, vmax, v1);
        select4(rb2, identity, vmin, vmax, v2);
        cs.q[0] = v1[0]; cs.q[1] = v1[1]; cs.q[2] = v1[2]; cs.q[3] = v1[3]; cs.q[4] = v2[0]; cs.q[5] = v2[1];
        cs.mad_lo = cs.mad_hi = 0;
        if (feats & ABX_F_CPMAD) {
          // twice the median (an integer): f = 1/2 exactly when n is odd
          const u64 med2 = ((n & 1u) && i2 < last) ? (u64)v1[2] + v1[3] : 2ull * v1[2];
          const auto absdev = [med2](u32 x) {
            const i64 d = 2ll * (i64)x - (i64)med2;
            return (u32)((d < 0 ? -d : d) >> 1);
          };
          const u32 rmad[4] = {i2, min(i2 + 1u, last), i2, min(i2 + 1u, last)};
          u32 vm[4];
          select4(rmad, absdev, 0u, max(absdev(vmin), absdev(vmax)), vm);
          cs.mad_lo = vm[0];
          cs.mad_hi = vm[1] | ((u32)(med2 & 1ull) << 31);
        }
        cs.maxpos = first_max_pos;
      }
      if (threadIdx.x == 0) *dst = cs;
      __syncthreads();  // smem reused by the next request
    }
  }
}

}  // namespace

int launch_object_stats(const abx_extract_args* a, const Workspace& ws, cudaStream_t st, bool all_objects) {
  if (a->n_requests == 0) return ABX_OK;
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  if (n_total == 0) return ABX_OK;
  const int grid = n_total < 148 * 8 ? n_total : 148 * 8;
#define ABX_LAUNCH_OS(PX)                                                                                          \
  object_stats_kernel<PX><<<grid, kThreads, 0, st>>>(                                                             \
      static_cast<const uint16_t*>(a->labels), a->label_plane_stride, a->label_row_stride, a->plane_tile,         \
      a->plane_base, a->n_planes, a->n_objects, n_total, static_cast<const PX*>(a->pixels),                       \
      reinterpret_cast<const i64*>(a->tile_offset), a->chan_stride, a->z_stride, a->row_stride, a->Z, a->requests, \
      a->n_requests, ws.recs, ws.chan, ws.stats_list, ws.list_counts, (int)abx_big_background(a),           \
      all_objects ? 1 : 0, ws.err)
  if (a->pixel_dtype == ABX_U16) ABX_LAUNCH_OS(uint16_t);
  else if (a->pixel_dtype == ABX_U8) ABX_LAUNCH_OS(uint8_t);
  else return ABX_OK;  // float pixels: every request belongs to object_float.cu
#undef ABX_LAUNCH_OS
  return abx_check_cuda(cudaGetLastError(), "object_stats");
}
