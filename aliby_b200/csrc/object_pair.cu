// Two-image features of `extractmulti_*` steps: for every object and every pair of requests (x = the object's values in
// request a, y = in request b) the exact integer sums behind CellProfiler's MeasureColocalization for objects — what
// src/extraction/extract.py:200-237 (measure_multi, `red_ch == "None"`) obtains from cp_measure through
// loaders.py:75-77,153-168, one object at a time:
//
//   sum x y                                             -> Pearson correlation (with n, sum x, sum x^2 ... of ChanStats)
//   tx = fraction * max x, ty = fraction * max y        (the maxima come from the statistics kernels: one pass suffices)
//   sum x [x >= tx], sum y [y >= ty]                    -> denominators of Manders / RWC
//   over "both" (x >= tx and y >= ty): sum x, y, xy, x^2, y^2        -> Manders, overlap, K1, K2
//   over both: sum x (R - |rank x - rank y|), the same for y          -> rank-weighted colocalisation
//
// Dense ranks (equal values share a rank) without sorting: a presence bit per value of [vmin, vmax] in shared memory
// (at most 65 536 bits per image), an exclusive prefix of the word popcounts, rank(v) = prefix[word] + popc(bits below v).
//
// Two kernels.  object_pair_object: ONE CTA of 8 warps per OBJECT — the window of every request is staged in shared
// memory once (a warp per (request, row) line, independent loads; Z stacks reduced per pixel on the fly), the row masks
// come from the label window by ballot, the rank table of a request is built once per object (not once per pair), and
// then a warp per pair sums from shared memory (the values are staged compacted — pixel k of the object — so every
// lane of the sum loop is busy).  Five channels and ten pairs read every pixel once instead of forty times.
// Objects it cannot hold (window above 64 x 64, more than 24 576 staged values or 8 requests, a value range above
// 16 384) put their pairs on a list for object_pair_cta: one CTA of 4 warps per (object, pair) with 65 536-bit tables,
// reading the label plane and the pixels from global memory (objects of any size).
// This is a "next" row of the hot path: see DESIGN.md for its measured cost.
#include "common.cuh"

namespace {

#include "warp_common.cuh"

constexpr int kPairWarps = 4;
constexpr int kPairThreads = kPairWarps * 32;
constexpr u32 kRankWords = 2048;     // CTA kernel: 65 536 presence bits per image
constexpr u32 kWarpRankWords = 512;  // warp kernel: 16 384
constexpr u32 kWarpMaxPixels = 16384;

struct PairArgs {
  const abx_object_rec* recs;
  const ChanStats* chan;
  PairStats* out;
  const abx_pair* pairs;
  const abx_request* requests;
  const uint16_t* labels;
  const void* pixels;
  const int32_t* plane_tile;
  const int32_t* plane_base;
  const int64_t* tile_offset;
  u32* err;
  u32* counts;     // Workspace::list_counts
  int* wide_list;  // items left to the CTA kernel
  i64 label_plane_stride, label_row_stride, chan_stride, z_stride, row_stride;
  int n_planes, n_objects, n_pairs, n_requests, Z;
};

template <typename PX>
__device__ __forceinline__ u32 value_at(const PX* __restrict__ p, int Z, i64 z_stride, int reduction) {
  u32 v = p[0];
  for (int z = 1; z < Z; ++z) {
    const u32 t = p[(i64)z * z_stride];
    v = reduction == ABX_RED_MAX ? max(v, t) : v + t;
  }
  return v;
}

// What both kernels know about an item before they touch a pixel.
template <typename PX>
struct Item {
  abx_object_rec rec;
  abx_pair pr;
  abx_request qa, qb;
  const uint16_t* lab;  // the object's label plane
  const PX *pa, *pb;    // the two channels of its pixel tile
  u32 id, amin, amax, bmin, bmax;
  double tx, ty;
  bool ranks;
};

template <typename PX>
__device__ __forceinline__ Item<PX> load_item(const PairArgs& a, i64 item) {
  Item<PX> it;
  const int obj = (int)(item / a.n_pairs), pi = (int)(item - (i64)obj * a.n_pairs);
  it.rec = a.recs[obj];
  it.pr = a.pairs[pi];
  const ChanStats* ca = a.chan + (i64)obj * a.n_requests + it.pr.request_a;
  const ChanStats* cb = a.chan + (i64)obj * a.n_requests + it.pr.request_b;
  it.amin = ca->vmin; it.amax = ca->vmax; it.bmin = cb->vmin; it.bmax = cb->vmax;
  const int plane = find_plane(a.plane_base, a.n_planes, obj);
  it.id = (u32)(obj - a.plane_base[plane]) + 1u;
  it.lab = a.labels + (i64)plane * a.label_plane_stride;
  it.qa = a.requests[it.pr.request_a];
  it.qb = a.requests[it.pr.request_b];
  const PX* base = static_cast<const PX*>(a.pixels) + a.tile_offset[a.plane_tile[plane]];
  it.pa = base + (i64)it.qa.channel * a.chan_stride;
  it.pb = base + (i64)it.qb.channel * a.chan_stride;
  it.tx = it.pr.threshold_fraction * (double)it.amax;
  it.ty = it.pr.threshold_fraction * (double)it.bmax;
  it.ranks = (it.pr.features & ABX_PF_RWC) != 0;
  return it;
}

// Walk the object's pixels: rows r0, r0 + step, ... of the bounding box, this lane's columns; f(x, y) for every pixel
// of the object.  The label and the two values of a pixel are loaded independently (the window lies inside the plane),
// four rows at a time, so that twelve loads are in flight instead of a chain of two.
template <typename PX, typename F>
__device__ __forceinline__ void for_each_pixel(const Item<PX>& it, const PairArgs& a, u32 r0, u32 step, u32 lane, F f) {
  for (u32 c = it.rec.cmin + lane; c <= it.rec.cmax; c += 32u) {
    for (u32 r = it.rec.rmin + r0; r <= it.rec.rmax; r += 4u * step) {
      u32 id[4], x[4], y[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const u32 rr = min(r + (u32)k * step, it.rec.rmax);  // (a repeated last row is masked out below)
        const i64 off = (i64)rr * a.row_stride + c;
        id[k] = it.lab[(i64)rr * a.label_row_stride + c];
        x[k] = value_at(it.pa + off, a.Z, a.z_stride, it.qa.reduction);
        y[k] = value_at(it.pb + off, a.Z, a.z_stride, it.qb.reduction);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (id[k] == it.id && r + (u32)k * step <= it.rec.rmax) f(x[k], y[k]);
    }
  }
}

// The ten sums of one thread, and one pixel's contribution to them.
struct Sums {
  u64 v[10];  // sxy, tot_x, tot_y, cx, cy, cxy, cxx, cyy, wx, wy
  u32 n_both;
};

template <typename Word, typename Prefix>
__device__ __forceinline__ void add_pixel(Sums& s, u32 x, u32 y, double tx, double ty, bool ranks, u32 amin, u32 bmin,
                                          u32 big_r, const Word* bits_a, const Word* bits_b, const Prefix* pre_a,
                                          const Prefix* pre_b) {
  const u64 xy = (u64)x * y;
  s.v[0] += xy;
  const bool ox = (double)x >= tx, oy = (double)y >= ty;
  if (ox) s.v[1] += x;
  if (oy) s.v[2] += y;
  if (ox && oy) {
    ++s.n_both;
    s.v[3] += x;
    s.v[4] += y;
    s.v[5] += xy;
    s.v[6] += (u64)x * x;
    s.v[7] += (u64)y * y;
    if (ranks) {
      const u32 xr = x - amin, yr = y - bmin;
      const u32 ra = pre_a[xr >> 5] + __popc(bits_a[xr >> 5] & ((1u << (xr & 31u)) - 1u));
      const u32 rb = pre_b[yr >> 5] + __popc(bits_b[yr >> 5] & ((1u << (yr & 31u)) - 1u));
      const u32 wgt = big_r - (ra > rb ? ra - rb : rb - ra);
      s.v[8] += (u64)x * wgt;
      s.v[9] += (u64)y * wgt;
    }
  }
}

// ---------------------------------------------------------------------------------------- one CTA per object
// Shared memory of object_pair_object (dynamic): the object's values in every request, COMPACTED (pixel k of the object
// in row-major order at q n + k) u16 [kValBudget] | row masks u64 [64] | first pixel index of every row u16 [64] |
// presence bits u32 [kMaxStaged][512] | rank prefixes u16 [kMaxStaged][512].
constexpr int kObjWarps = 8;
constexpr int kObjThreads = kObjWarps * 32;
constexpr u32 kValBudget = 24576;  // staged values: six requests of a full 64 x 64 object
constexpr int kMaxStaged = 8;      // requests with rank tables
constexpr u32 kObjValOff = 0, kObjMaskOff = kValBudget * 2, kObjRowOff = kObjMaskOff + 512, kObjBitsOff = kObjRowOff + 128,
              kObjPreOff = kObjBitsOff + kMaxStaged * kWarpRankWords * 4,
              kObjSmem = kObjPreOff + kMaxStaged * kWarpRankWords * 2;

// One pixel of a pair, thresholds as integers (x >= tx <=> x >= ceil(tx) for an integer x).  Per-lane partial sums: a
// lane sees at most 128 of an object's 4096 pixels, so the plain sums of 16-bit values fit 32 bits.
struct LaneSums {
  u64 sxy, cxy, cxx, cyy, wx, wy;
  u32 tot_x, tot_y, cx, cy, n_both;
};
__device__ __forceinline__ void add_staged(LaneSums& s, u32 x, u32 y, u32 itx, u32 ity, bool ranks, u32 amin, u32 bmin,
                                           u32 big_r, const u32* ba, const u32* bb, const uint16_t* pa, const uint16_t* pb) {
  const u64 xy = (u64)x * y;
  s.sxy += xy;
  const bool ox = x >= itx, oy = y >= ity;
  s.tot_x += ox ? x : 0u;
  s.tot_y += oy ? y : 0u;
  if (ox && oy) {
    ++s.n_both;
    s.cx += x;
    s.cy += y;
    s.cxy += xy;
    s.cxx += (u64)(x * x);  // (32-bit products of 16-bit values)
    s.cyy += (u64)(y * y);
    if (ranks) {
      const u32 xr = x - amin, yr = y - bmin;
      const u32 ra = pa[xr >> 5] + __popc(ba[xr >> 5] & ((1u << (xr & 31u)) - 1u));
      const u32 rb = pb[yr >> 5] + __popc(bb[yr >> 5] & ((1u << (yr & 31u)) - 1u));
      const u32 wgt = big_r - (ra > rb ? ra - rb : rb - ra);
      s.wx += (u64)(x * wgt);  // x < 2^16, wgt <= 4096 distinct values
      s.wy += (u64)(y * wgt);
    }
  }
}

template <typename PX>
__global__ void __launch_bounds__(kObjThreads)
object_pair_object(const PairArgs a) {
  uint16_t* val = reinterpret_cast<uint16_t*>(dyn + kObjValOff);
  u64* mask = reinterpret_cast<u64*>(dyn + kObjMaskOff);
  uint16_t* rowbase = reinterpret_cast<uint16_t*>(dyn + kObjRowOff);
  u32* bits = reinterpret_cast<u32*>(dyn + kObjBitsOff);
  uint16_t* pre = reinterpret_cast<uint16_t*>(dyn + kObjPreOff);
  __shared__ int s_obj;
  __shared__ u32 s_list_base;
  __shared__ u32 s_distinct[kMaxStaged];
  const u32 lane = lane_id(), warp = threadIdx.x >> 5;
  const int R = a.n_requests;
  u32 feats = 0;
  for (int p = threadIdx.x; p < a.n_pairs; p += kObjThreads) feats |= a.pairs[p].features;
  const bool any_ranks = __syncthreads_or((feats & ABX_PF_RWC) != 0) != 0;  // some pair wants the rank tables
  if (threadIdx.x == 0) s_obj = (int)atomicAdd(a.counts + kCntPairWork, 1u);
  __syncthreads();
  for (int obj = s_obj; obj < a.n_objects; obj = s_obj) {
    __syncthreads();  // everybody has read s_obj, and the previous object's shared memory is no longer read
    if (threadIdx.x == 0) s_obj = (int)atomicAdd(a.counts + kCntPairWork, 1u);  // the NEXT object: the atomic's latency hides behind this one
    do {
    const abx_object_rec rec = a.recs[obj];
    PairStats* out = a.out + (i64)obj * a.n_pairs;
    if (rec.n == 0) {  // absent label: zero records -> NaN in finalize
      PairStats z;
      memset(&z, 0, sizeof(z));
      for (int p = threadIdx.x; p < a.n_pairs; p += kObjThreads) out[p] = z;
      break;
    }
    const u32 h = rec.rmax - rec.rmin + 1u, w = rec.cmax - rec.cmin + 1u, n = rec.n;
    // ---- can this CTA hold the object?  (block-uniform: every thread evaluates the same records) ----
    bool fits = h <= (u32)kSide && w <= (u32)kSide && (u32)R * n <= kValBudget && R <= kMaxStaged;
    const ChanStats* cs = a.chan + (i64)obj * R;
    for (int q = 0; q < R; ++q) {
      const u32 lo = cs[q].vmin, hi = cs[q].vmax;
      if (hi >= 65536u) fits = false;  // values of 65 536 or more (the `add` of a stack): the CTA kernel flags the pairs that use them
      else if (any_ranks && (hi - lo) / 32u + 1u > kWarpRankWords) fits = false;
    }
    if (!fits) {
      if (threadIdx.x == 0) s_list_base = atomicAdd(a.counts + kCntPairWide, (u32)a.n_pairs);
      __syncthreads();
      for (int p = threadIdx.x; p < a.n_pairs; p += kObjThreads) a.wide_list[s_list_base + p] = obj * a.n_pairs + p;
      break;
    }
    const int plane = find_plane(a.plane_base, a.n_planes, obj);
    const u32 id = (u32)(obj - a.plane_base[plane]) + 1u;
    const uint16_t* lab = a.labels + (i64)plane * a.label_plane_stride + (i64)rec.rmin * a.label_row_stride + rec.cmin;
    const PX* base = static_cast<const PX*>(a.pixels) + a.tile_offset[a.plane_tile[plane]] + (i64)rec.rmin * a.row_stride + rec.cmin;
    // ---- row masks from the label window (a warp per row), index of every row's first pixel, empty rank tables ----
    for (u32 r0 = warp; r0 < h; r0 += 4u * kObjWarps) {  // four rows' loads in flight
      u32 l0[4], l1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const u32 r = r0 + (u32)k * kObjWarps;
        l0[k] = (r < h && lane < w) ? lab[(i64)r * a.label_row_stride + lane] : 0u;
        l1[k] = (r < h && lane + 32u < w) ? lab[(i64)r * a.label_row_stride + lane + 32u] : 0u;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const u32 r = r0 + (u32)k * kObjWarps;
        const u32 m0 = __ballot_sync(kFull, l0[k] == id), m1 = __ballot_sync(kFull, l1[k] == id);  // (id >= 1)
        if (lane == 0 && r < h) mask[r] = (u64)m0 | ((u64)m1 << 32);
      }
    }
    if (any_ranks)
      for (u32 i = threadIdx.x; i < (u32)R * kWarpRankWords; i += kObjThreads) bits[i] = 0;  // (whole tables: 16 stores per thread)
    __syncthreads();
    if (warp == 0) {
      const u32 c0 = lane < h ? (u32)__popcll(mask[lane]) : 0u, c1 = lane + 32u < h ? (u32)__popcll(mask[lane + 32u]) : 0u;
      u32 i0 = c0, i1 = c1;
#pragma unroll
      for (int k = 1; k < 32; k <<= 1) {
        const u32 t0 = __shfl_up_sync(kFull, i0, k), t1 = __shfl_up_sync(kFull, i1, k);
        if ((int)lane >= k) { i0 += t0; i1 += t1; }
      }
      const u32 first_half = __shfl_sync(kFull, i0, 31);
      rowbase[lane] = (uint16_t)(i0 - c0);
      rowbase[lane + 32u] = (uint16_t)(first_half + i1 - c1);
    }
    __syncthreads();
    // ---- stage: a warp per (request, row) line, only the object's pixels, independent loads; presence bits on the way ----
    for (u32 q = 0; q < (u32)R; ++q) {
      const abx_request rq = a.requests[q];
      const PX* src_q = base + (i64)rq.channel * a.chan_stride;
      const u32 lo = cs[q].vmin;
      uint16_t* dst_q = val + q * n;
      u32* bits_q = bits + q * kWarpRankWords;
      for (u32 r0 = warp; r0 < h; r0 += 4u * kObjWarps) {  // four rows' loads in flight
        u32 x[4][2];
        u64 mk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const u32 r = r0 + (u32)k * kObjWarps;
          mk[k] = r < h ? mask[r] : 0ull;
          const PX* src = src_q + (i64)r * a.row_stride;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const u32 c = lane + 32u * half;
            x[k][half] = ((mk[k] >> c) & 1ull) ? value_at(src + c, a.Z, a.z_stride, rq.reduction) : 0u;
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const u32 r = r0 + (u32)k * kObjWarps;
          if (r >= h) break;
          uint16_t* dst = dst_q + rowbase[r];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const u32 c = lane + 32u * half;
            if ((mk[k] >> c) & 1ull) {
              dst[__popcll(mk[k] & ((1ull << c) - 1ull))] = (uint16_t)x[k][half];
              if (any_ranks) atomicOr(&bits_q[(x[k][half] - lo) >> 5], 1u << ((x[k][half] - lo) & 31u));
            }
          }
        }
      }
    }
    __syncthreads();
    if (any_ranks) {
      // ---- a warp per request: the exclusive prefix of the word popcounts ----
      for (u32 q = warp; q < (u32)R; q += kObjWarps) {
        const u32 nw = (cs[q].vmax - cs[q].vmin) / 32u + 1u;
        const u32* bq = bits + q * kWarpRankWords;
        uint16_t* pq = pre + q * kWarpRankWords;
        const u32 per = (nw + 31u) >> 5;  // lane L owns the words [per L, per L + per)
        u32 mine = 0;
        for (u32 k = 0; k < per; ++k) {
          const u32 i = lane * per + k;
          if (i < nw) mine += __popc(bq[i]);
        }
        u32 incl = mine;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
          const u32 t = __shfl_up_sync(kFull, incl, k);
          if ((int)lane >= k) incl += t;
        }
        u32 before = incl - mine;
        for (u32 k = 0; k < per; ++k) {
          const u32 i = lane * per + k;
          if (i < nw) {
            pq[i] = (uint16_t)before;
            before += __popc(bq[i]);
          }
        }
        if (lane == 31u) s_distinct[q] = incl;
      }
      __syncthreads();
    }
    // ---- the sums: a warp per pair over the compacted values, every lane busy ----
    for (int p = warp; p < a.n_pairs; p += kObjWarps) {
      const abx_pair pr = a.pairs[p];
      const u32 qa = (u32)pr.request_a, qb = (u32)pr.request_b;
      const u32 amin = cs[qa].vmin, amax = cs[qa].vmax, bmin = cs[qb].vmin, bmax = cs[qb].vmax;
      const u32 itx = (u32)ceil(pr.threshold_fraction * (double)amax), ity = (u32)ceil(pr.threshold_fraction * (double)bmax);
      const bool ranks = (pr.features & ABX_PF_RWC) != 0;
      const u32 big_r = ranks ? max(s_distinct[qa], s_distinct[qb]) : 0u;
      const u32 *ba = bits + qa * kWarpRankWords, *bb = bits + qb * kWarpRankWords;
      const uint16_t *pa = pre + qa * kWarpRankWords, *pb = pre + qb * kWarpRankWords;
      const uint16_t *xa = val + qa * n, *xb = val + qb * n;
      LaneSums s;
      memset(&s, 0, sizeof(s));
#pragma unroll 4
      for (u32 k = lane; k < n; k += 32u) add_staged(s, xa[k], xb[k], itx, ity, ranks, amin, bmin, big_r, ba, bb, pa, pb);
      PairStats ps;
      ps.sxy = warp_sum64(s.sxy);
      ps.tot_x = warp_sum64(s.tot_x);
      ps.tot_y = warp_sum64(s.tot_y);
      ps.cx = warp_sum64(s.cx);
      ps.cy = warp_sum64(s.cy);
      ps.cxy = warp_sum64(s.cxy);
      ps.cxx = warp_sum64(s.cxx);
      ps.cyy = warp_sum64(s.cyy);
      ps.wx = warp_sum64(s.wx);
      ps.wy = warp_sum64(s.wy);
      ps.n_both = __reduce_add_sync(kFull, s.n_both);
      ps.big_r = big_r;
      ps.flags = 0;
      ps.pad_ = 0;
      if (lane == 0) out[p] = ps;
    }
    } while (false);
    __syncthreads();  // thread 0's s_obj (the next object) is visible
  }
}

// ---------------------------------------------------------------------------------------- one CTA per listed item
template <typename PX>
__global__ void __launch_bounds__(kPairThreads)
object_pair_cta(const PairArgs a) {
  __shared__ u32 bits[2][kRankWords];
  __shared__ uint16_t prefix[2][kRankWords];
  __shared__ u64 partial[kPairWarps][10];
  __shared__ u32 scan_tot[2][kPairWarps];
  __shared__ u32 n_both_w[kPairWarps];
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const u32 n_listed = a.counts[kCntPairWide];
  for (u32 k_item = blockIdx.x; k_item < n_listed; k_item += gridDim.x) {
    const i64 item = a.wide_list[k_item];
    const Item<PX> it = load_item<PX>(a, item);
    if (it.amax >= 65536u || it.bmax >= 65536u) {  // (block-uniform) values the sums are not sized for
      if (threadIdx.x == 0) {
        PairStats ps;
        memset(&ps, 0, sizeof(ps));
        ps.flags = 1u;
        a.out[item] = ps;
        atomicOr(a.err, 4u);
      }
      continue;
    }
    u32 big_r = 0;
    if (it.ranks) {
      // ---- presence bits of both images over [vmin, vmax], then the exclusive prefix of the word popcounts ----
      const u32 wa = (it.amax - it.amin) / 32u + 1u, wb = (it.bmax - it.bmin) / 32u + 1u;
      for (u32 i = threadIdx.x; i < wa; i += kPairThreads) bits[0][i] = 0;
      for (u32 i = threadIdx.x; i < wb; i += kPairThreads) bits[1][i] = 0;
      __syncthreads();
      for_each_pixel(it, a, warp, (u32)kPairWarps, lane, [&](u32 x, u32 y) {
        x -= it.amin;
        y -= it.bmin;
        atomicOr(&bits[0][x >> 5], 1u << (x & 31u));
        atomicOr(&bits[1][y >> 5], 1u << (y & 31u));
      });
      __syncthreads();
#pragma unroll
      for (int im = 0; im < 2; ++im) {
        const u32 nw = im ? wb : wa;
        const u32 per = (nw + kPairThreads - 1u) / kPairThreads;  // thread t owns the words [per t, per t + per)
        u32 mine = 0;
        for (u32 k = 0; k < per; ++k) {
          const u32 i = threadIdx.x * per + k;
          if (i < nw) mine += __popc(bits[im][i]);
        }
        u32 incl = mine;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
          const u32 t = __shfl_up_sync(kFull, incl, k);
          if ((int)lane >= k) incl += t;
        }
        if (lane == 31u) scan_tot[im][warp] = incl;
        __syncthreads();
        u32 before = incl - mine;
        for (u32 w = 0; w < warp; ++w) before += scan_tot[im][w];
        for (u32 k = 0; k < per; ++k) {
          const u32 i = threadIdx.x * per + k;
          if (i < nw) {
            prefix[im][i] = (uint16_t)before;  // < 65 536: at most 65 535 distinct values lie below any word's first bit
            before += __popc(bits[im][i]);
          }
        }
      }
      __syncthreads();
      u32 da = 0, db = 0;  // numbers of distinct values
      for (int w = 0; w < kPairWarps; ++w) { da += scan_tot[0][w]; db += scan_tot[1][w]; }
      big_r = max(da, db);
    }
    Sums s;
    memset(&s, 0, sizeof(s));
    for_each_pixel(it, a, warp, (u32)kPairWarps, lane, [&](u32 x, u32 y) {
      add_pixel(s, x, y, it.tx, it.ty, it.ranks, it.amin, it.bmin, big_r, bits[0], bits[1], prefix[0], prefix[1]);
    });
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const u64 t = warp_sum64(s.v[k]);
      if (lane == 0) partial[warp][k] = t;
    }
    const u32 nb = __reduce_add_sync(kFull, s.n_both);
    if (lane == 0) n_both_w[warp] = nb;
    __syncthreads();
    if (threadIdx.x == 0) {
      u64 t[10];
      for (int k = 0; k < 10; ++k) {
        t[k] = 0;
        for (int w = 0; w < kPairWarps; ++w) t[k] += partial[w][k];
      }
      PairStats ps;
      memset(&ps, 0, sizeof(ps));
      ps.sxy = t[0]; ps.tot_x = t[1]; ps.tot_y = t[2]; ps.cx = t[3]; ps.cy = t[4];
      ps.cxy = t[5]; ps.cxx = t[6]; ps.cyy = t[7]; ps.wx = t[8]; ps.wy = t[9];
      ps.big_r = big_r;
      for (int w = 0; w < kPairWarps; ++w) ps.n_both += n_both_w[w];
      a.out[item] = ps;
    }
    __syncthreads();  // shared memory is reused by the next item
  }
}

}  // namespace

int launch_object_pair(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  const i64 n_items = (i64)a->n_objects * a->n_pairs;
  if (n_items == 0) return ABX_OK;
  if (n_items > 2147483647LL) return abx_set_error(ABX_ERR_INVALID, "too many (object, pair) items");
  PairArgs p;
  p.recs = ws.recs;
  p.chan = ws.chan;
  p.out = ws.pairs;
  p.pairs = a->pairs;
  p.requests = a->requests;
  p.labels = static_cast<const uint16_t*>(a->labels);
  p.pixels = a->pixels;
  p.plane_tile = a->plane_tile;
  p.plane_base = a->plane_base;
  p.tile_offset = a->tile_offset;
  p.err = ws.err;
  p.counts = ws.list_counts;
  p.wide_list = ws.pair_wide;
  p.label_plane_stride = a->label_plane_stride;
  p.label_row_stride = a->label_row_stride;
  p.chan_stride = a->chan_stride;
  p.z_stride = a->z_stride;
  p.row_stride = a->row_stride;
  p.n_planes = a->n_planes;
  p.n_objects = a->n_objects;
  p.n_pairs = a->n_pairs;
  p.n_requests = a->n_requests;
  p.Z = a->Z;
  static thread_local bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(object_pair_object<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kObjSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(object_pair_object<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kObjSmem);
    if (e != cudaSuccess) return abx_check_cuda(e, "object_pair_object smem attribute");
    done[dev] = true;
  }
  const unsigned grid = (unsigned)(a->n_objects < 148 * 3 ? a->n_objects : 148 * 3);  // persistent, 3 CTAs of 8 warps per SM
  const unsigned grid_cta = (unsigned)(n_items < 148 * 8 ? n_items : 148 * 8);
  if (a->pixel_dtype == ABX_U8) {
    object_pair_object<uint8_t><<<grid, kObjThreads, kObjSmem, st>>>(p);
    object_pair_cta<uint8_t><<<grid_cta, kPairThreads, 0, st>>>(p);
  } else {
    object_pair_object<uint16_t><<<grid, kObjThreads, kObjSmem, st>>>(p);
    object_pair_cta<uint16_t><<<grid_cta, kPairThreads, 0, st>>>(p);
  }
  return abx_check_cuda(cudaGetLastError(), "object_pair");
}
