// Per-object intensity statistics, ONE SWEEP per (object, request): one warp per object in its own single-warp CTA,
// the object's pixel window staged by TMA (cp.async.bulk.tensor boxes of 8 rows, mbarrier completion), the object's
// pixels addressed through a list built from the 64 x 64 torus bitmap the label scan wrote (label_scan.cu) — the label
// planes are not read here at all.  Same outputs and reference semantics as before
// (src/extraction/extract.py:346-359 loop; cell.py:43-157,232-265; tile crop of tiler.py:309-366 fused through the
// tile offset).
//
//   plan kernel   one thread per object: window geometry and TMA coordinates (ObjPlan), the work order (big objects
//                 first), zero records for absent labels, hand-over lists for what this kernel does not take
//                 (windows above 64 x 64 and backgrounds -> object_stats.cu; windows that start in front of the buffer ->
//                 object_stats_warp), the same routing for the shape kernel (object_edt.cu), and the second moments of
//                 the pixel coordinates for cp_measure `sizeshape`
//   sweep kernel  four warps per CTA, four CTAs per SM, warp per object.  Shared memory per warp:
//                   window [rows][pitch] PX (one TMA box, at most 64 rows x 144 bytes)
//                   hist u32[1024], 4 KB aligned in the shared window (bin address = base | (v << 2) & 0xFFC)
//                   scratch u32[16], mbarrier, row table
//                 The pixel list lives in TENSOR MEMORY (tcgen05.alloc / st / ld, 64 columns per CTA, one lane
//                 quarter per warp): 16-bit entries, each the SHARED ADDRESS of a pixel.
//                 Per request ONE pass over the list: moments, extrema and a histogram of the LOW 10 BITS of every
//                 value.  When max - min < 1021 (checked afterwards) that circular histogram is exact — bin
//                 (v & 1023) holds one value only — and the four ranks (two medians, top 2.5 %, top 5) come out of
//                 it by warp scans; wider ranges take a coarse histogram + 7-bit refinement sweeps on the window
//                 that is still resident.  The TMA copy of the NEXT window (next request, or first request of the
//                 warp's next object) is issued right after the pass, so that it runs under the rank search, and
//                 the window after that one is prefetched into L2 by cp.async.bulk.prefetch.tensor.
//                 cp_measure `intensity` adds six CellProfiler-rule ranks and a second pass (MAD, first maximum).
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <cstring>

#include "common.cuh"

namespace {

#include "warp_common.cuh"
#include "tma.cuh"

constexpr u32 kBigFirst = 1536;  // objects above this many pixels are processed first

// ---- shared-memory accessors on 32-bit shared addresses (no generic-pointer arithmetic in the hot loops) ----
__device__ __forceinline__ u32 lds_u16(u32 a) { u32 v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
template <int kOff>
__device__ __forceinline__ u32 lds_u16_off(u32 a) {
  u32 v;
  asm volatile("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(kOff));
  return v;
}
__device__ __forceinline__ u32 lds_u8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u16(u32 a, u32 v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
template <typename PX>
__device__ __forceinline__ u32 lds_px(u32 a) { return sizeof(PX) == 1 ? lds_u8(a) : lds_u16(a); }

// ------------------------------------------------------------------------------------------------
// plan kernel
// ------------------------------------------------------------------------------------------------
// Appends `item` to a list that is filled from the front (front = true) or from the back, one atomic per warp.
__device__ __forceinline__ void append(bool want, bool front, int item, int* __restrict__ list, int cap, u32* cnt_front,
                                       u32* cnt_back) {
  const u32 lane = lane_id();
  const u32 mf = __ballot_sync(kFull, want && front), mb = __ballot_sync(kFull, want && !front);
  u32 bf = 0, bb = 0;
  if (lane == 0) {
    if (mf) bf = atomicAdd(cnt_front, (u32)__popc(mf));
    if (mb) bb = atomicAdd(cnt_back, (u32)__popc(mb));
  }
  bf = __shfl_sync(kFull, bf, 0);
  bb = __shfl_sync(kFull, bb, 0);
  const u32 lt = (1u << lane) - 1u;
  if (want && front) list[bf + (u32)__popc(mf & lt)] = item;
  if (want && !front) list[cap - 1 - (int)(bb + (u32)__popc(mb & lt))] = item;
}

struct PlanArgs {
  const abx_object_rec* recs;
  const int32_t* plane_base;
  const int32_t* plane_tile;
  const i64* tile_offset;
  int n_planes, n_objects, n_total;
  i64 row_stride;
  int align;       // elements per 16 bytes of the pixel dtype
  int n_requests;
  int sweep;       // the sweep kernel takes the statistics (otherwise object_stats_warp does its own bookkeeping)
  int need_edt;
  ChanStats* chan;
  ShapeStats* shape;
  ObjPlan* plan;
  int* order_stats;
  int* order_edt;
  int* stats_list;  // hand-over to the CTA-per-object statistics kernel
  int* pair_list;   // (object, request) pairs for object_stats_warp
  int* edt_list;    // hand-over to the CTA-per-object shape kernel
  u32* counts;      // Workspace::list_counts
  // cp_measure features
  int cp_requests;        // some request wants ABX_F_CPQ / ABX_F_CPMAD: only the sweep kernel computes those
  const abx_request* requests;  // as the statistics kernels see them (zreduce.cu marks what it leaves to the gather pass as DIV)
  int want_moments;       // need_edt bit 2: second coordinate moments of every object, from its bitmap
  const u64* bitmaps;
  MaskMoments* mom;
  u32* err;
};

// sum of k and of k^2 over [a, b]
__device__ __forceinline__ u64 sum1(u64 a, u64 b) { return (a + b) * (b - a + 1) / 2; }
__device__ __forceinline__ u64 sum2(u64 a, u64 b) { return (b * (b + 1) * (2 * b + 1) - (a ? (a - 1) * a * (2 * a - 1) : 0)) / 6; }

__global__ void __launch_bounds__(256) plan_kernel(const PlanArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < a.n_total;
  abx_object_rec rec;
  rec.n = 0; rec.rmin = rec.rmax = rec.cmin = rec.cmax = 0;
  if (live) rec = a.recs[i];
  const bool is_bg = i >= a.n_objects;
  const int h = (int)(rec.rmax - rec.rmin) + 1, w = (int)(rec.cmax - rec.cmin) + 1;
  const bool windowed = live && rec.n > 0 && !is_bg && h <= kSide && w <= kSide;
  // cp_measure rank statistics exist in the sweep kernel and in the CTA-per-object kernel only: a request that goes to
  // the float kernel (`div`) or to the gather pass on the sum planes of a Z stack (`add`) has none -> status bit 1
  if (i == 0 && a.cp_requests)
    for (int q = 0; q < a.n_requests; ++q)
      if ((a.requests[q].features & (ABX_F_CPQ | ABX_F_CPMAD)) && a.requests[q].reduction == ABX_RED_DIV) atomicOr(a.err, 2u);
  // ---- statistics ----
  if (a.sweep) {
    bool take = false, too_wide = false;
    if (live && rec.n == 0) {  // absent label (or empty background): zero records -> NaN in finalize
      ChanStats z;
      z.sum = z.sumsq = z.wrapsq = z.m10 = z.m01 = z.m20 = z.m02 = z.top2p5_sum = z.top5_sum = 0;
      z.vmin = z.vmax = z.med_lo = z.med_hi = 0;
      for (int q = 0; q < a.n_requests; ++q) a.chan[(i64)i * a.n_requests + q] = z;
    }
    if (windowed) {
      const int p = find_plane(a.plane_base, a.n_planes, i);
      const i64 org = a.tile_offset[a.plane_tile[p]] + (i64)rec.rmin * a.row_stride + rec.cmin;
      const i64 row0 = org / a.row_stride;
      const int col0 = (int)(org - row0 * a.row_stride);
      const u32 s_px = (u32)col0 & (u32)(a.align - 1);
      if (org < 0) {
        too_wide = true;  // a window that starts in front of the buffer: the gather kernel addresses anything
      } else {
        ObjPlan pl;
        pl.tma_x = col0 - (int)s_px;
        pl.tma_y = (int)row0;
        pl.n = rec.n;
        pl.geom = (u32)(h - 1) | ((u32)(w - 1) << 6) | ((rec.rmin & 63u) << 12) | ((rec.cmin & 63u) << 18) | (s_px << 24);
        a.plan[i] = pl;
        take = true;
      }
    }
    append(take, rec.n > kBigFirst, i, a.order_stats, a.n_total, a.counts + kCntOrderBig, a.counts + kCntOrderSmall);
    // windows above 64 x 64 and backgrounds: the CTA-per-object kernel (object_stats.cu).  It also takes the (rare)
    // window that starts in front of the buffer when cp_measure statistics are wanted: the gather kernel has none.
    const bool hand = live && rec.n > 0 && (!windowed || (too_wide && a.cp_requests));
    {
      const u32 m = __ballot_sync(kFull, hand);
      u32 b = 0;
      if (lane_id() == 0 && m) b = atomicAdd(a.counts + kCntStatsList, (u32)__popc(m));
      b = __shfl_sync(kFull, b, 0);
      if (hand) a.stats_list[b + (u32)__popc(m & ((1u << lane_id()) - 1u))] = i;
    }
    if (too_wide && !a.cp_requests) {
      const u32 b = atomicAdd(a.counts + kCntLeftover, (u32)a.n_requests);  // rare: no aggregation
      for (int q = 0; q < a.n_requests; ++q) a.pair_list[b + q] = i * a.n_requests + q;
    }
  }
  // ---- second moments of the pixel coordinates (cp_measure sizeshape), relative to the bounding box origin ----
  if (a.want_moments && live && !is_bg) {
    MaskMoments mm;
    mm.s_rr = mm.s_cc = mm.s_rc = mm.pad_ = 0;
    if (rec.n > 0 && h <= kSide && w <= kSide) {
      const u64* bm = a.bitmaps + (size_t)i * 64u;
      const u32 rot = rec.cmin & 63u;
      for (int r = 0; r < h; ++r) {
        u64 m = bm[(rec.rmin + (u32)r) & 63u];
        m = (m >> rot) | (rot ? (m << (64u - rot)) : 0ull);
        while (m) {  // run by run: closed forms for sum c and sum c^2
          const u32 c0 = (u32)__ffsll((long long)m) - 1u;
          const u64 run = m >> c0;
          const u32 len = (~run) ? (u32)__ffsll((long long)~run) - 1u : 64u - c0;
          const u64 s1 = sum1(c0, c0 + len - 1u);
          mm.s_cc += sum2(c0, c0 + len - 1u);
          mm.s_rc += (u64)r * s1;
          mm.s_rr += (u64)r * (u64)r * len;
          m = (len + c0 >= 64u) ? 0ull : (m & (~0ull << (c0 + len)));
        }
      }
      a.mom[i] = mm;
    } else if (rec.n == 0) {
      a.mom[i] = mm;
    }  // (a cell above 64 x 64 has no bitmap: large_moments_kernel reads its label window)
  }
  // ---- shape ----
  if (a.need_edt) {
    const bool obj = live && !is_bg;
    if (obj && rec.n == 0) {
      ShapeStats z;
      z.sum_nn = 0; z.sum_top = 0; z.max_nn2 = 0; z.max_dn2 = 0;
      a.shape[i] = z;
    }
    const bool take = obj && rec.n > 0 && h <= kSide && w <= kSide;
    append(take, rec.n > kBigFirst, i, a.order_edt, a.n_objects, a.counts + kCntEdtBig, a.counts + kCntEdtSmall);
    const bool hand = obj && rec.n > 0 && !take;
    if (hand) a.edt_list[atomicAdd(a.counts + kCntEdtList, 1u)] = i;  // rare
  }
}

// Second coordinate moments of the cells above 64 x 64 (no bitmap): one CTA per such cell walks its label window.
// Every CTA looks at the records of its share of the objects and skips the window-sized ones (a record load each).
__global__ void __launch_bounds__(256) large_moments_kernel(const abx_object_rec* __restrict__ recs, int n_objects,
                                                            const uint16_t* __restrict__ labels, i64 plane_stride,
                                                            i64 row_stride, const int32_t* __restrict__ plane_base,
                                                            int n_planes, MaskMoments* __restrict__ mom) {
  __shared__ u64 acc[3];
  for (int i = blockIdx.x; i < n_objects; i += gridDim.x) {
    const abx_object_rec rec = recs[i];
    const u32 h = rec.rmax - rec.rmin + 1u, w = rec.cmax - rec.cmin + 1u;
    if (rec.n == 0 || (h <= (u32)kSide && w <= (u32)kSide)) continue;  // (block-uniform)
    if (threadIdx.x < 3) acc[threadIdx.x] = 0;
    __syncthreads();
    const int p = find_plane(plane_base, n_planes, i);
    const u32 id = (u32)(i - plane_base[p]) + 1u;
    const uint16_t* lab = labels + (i64)p * plane_stride + (i64)rec.rmin * row_stride + rec.cmin;
    u64 s_rr = 0, s_cc = 0, s_rc = 0;
    for (u32 r = threadIdx.x >> 5; r < h; r += 8u)
      for (u32 c = threadIdx.x & 31u; c < w; c += 32u)
        if (lab[(i64)r * row_stride + c] == id) {
          s_rr += (u64)r * r;
          s_cc += (u64)c * c;
          s_rc += (u64)r * c;
        }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
      s_rr += __shfl_xor_sync(kFull, s_rr, k);
      s_cc += __shfl_xor_sync(kFull, s_cc, k);
      s_rc += __shfl_xor_sync(kFull, s_rc, k);
    }
    if ((threadIdx.x & 31u) == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&acc[0]), s_rr);
      atomicAdd(reinterpret_cast<unsigned long long*>(&acc[1]), s_cc);
      atomicAdd(reinterpret_cast<unsigned long long*>(&acc[2]), s_rc);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      MaskMoments mm;
      mm.s_rr = acc[0]; mm.s_cc = acc[1]; mm.s_rc = acc[2]; mm.pad_ = 0;
      mom[i] = mm;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// sweep kernel
// ------------------------------------------------------------------------------------------------
// ---- tensor memory: the pixel list of the warp's object lives in TMEM (tcgen05.ld / tcgen05.st) ----
// A CTA of four warps allocates 64 columns; warp w owns TMEM lanes [32 w, 32 w + 32), i.e. 32 lanes x 64 columns x
// 32 bits = 4096 list entries of 16 bits.  Word (lane L, column c) = entries (2 c) * 32 + L (low half) and
// (2 c + 1) * 32 + L (high half): for a fixed column and half the 32 lanes hold 32 CONSECUTIVE list positions —
// consecutive pixels of a row, i.e. conflict-free shared-memory reads of the window.  The list never touches shared
// memory: the 9 KB a window can take plus the 4 KB histogram are all an object needs there, so 16 objects are in
// flight per SM instead of 12.
constexpr u32 kTmemCols = 64;
__device__ __forceinline__ void tmem_ld4(u32 taddr, u32 (&w)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ u32 tmem_ld1(u32 taddr) {
  u32 w;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(w) : "r"(taddr) : "memory");
  tmem_wait_ld();
  return w;
}
__device__ __forceinline__ void tmem_st4(u32 taddr, const u32 (&w)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// f(shared address of the pixel, list position) for every list entry of this lane (positions lane, lane + 32, ...);
// warp-collective: every lane calls it with the same n.  For the rare passes; the hot sweep has its own loop.
template <typename F>
__device__ __forceinline__ void for_each_entry(u32 tlist, u32 n, F f) {
  const u32 lane = lane_id();
  const u32 cols = (n + 63u) >> 6;
#pragma unroll 1
  for (u32 c = 0; c < cols; ++c) {
    const u32 w = tmem_ld1(tlist + c);
    const u32 e0 = 64u * c + lane, e1 = e0 + 32u;
    if (e0 < n) f(w & 0xFFFFu, e0);
    if (e1 < n) f(w >> 16, e1);
  }
}

struct Geo {  // one object's window as the kernel sees it (warp-uniform)
  int obj;
  int tma_x, tma_y;  // box coordinates of the window in channel 0
  u32 n, h, w, s_px;
  u32 rot, row0;     // bitmap rotation (columns) and first bitmap row
  u32 pitchB;        // window row pitch in bytes (a multiple of 16, at most 144)
  u32 h8;            // rows rounded up to whole boxes of 8
};

template <typename PX>
__device__ __forceinline__ Geo make_geo(const ObjPlan& pl, int obj) {
  Geo g;
  g.obj = obj;
  g.tma_x = pl.tma_x; g.tma_y = pl.tma_y;
  g.n = pl.n;
  g.h = (pl.geom & 63u) + 1u;
  g.w = ((pl.geom >> 6) & 63u) + 1u;
  g.row0 = (pl.geom >> 12) & 63u;
  g.s_px = pl.geom >> 24;
  g.rot = (pl.geom >> 18) & 63u;  // the masks stay in bbox coordinates: the window is up to 15 columns wider than 64
  g.pitchB = ((g.w + g.s_px) * (u32)sizeof(PX) + 15u) & ~15u;
  g.h8 = (g.h + 7u) & ~7u;
  return g;
}

struct Acc {  // per-lane partial sums of one request
  u32 sum, wh, vmin, vmax, m10, m01;
  u64 sq, q;
};

template <typename PX>
__device__ __forceinline__ void accumulate(Acc& a, u32 v, u32 hbase) {
  constexpr int kShift = (sizeof(PX) == 1) ? 12 : 8;  // (x << kShift)^2 >> 32 == x^2 >> bits(PX)
  a.sum += v;
  a.sq += (u64)v * (u64)v;
  const u32 s = v << kShift;
  a.wh += __umulhi(s, s);
  a.vmin = min(a.vmin, v);
  a.vmax = max(a.vmax, v);
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hbase | ((v << 2) & 0xFFCu)) : "memory");
}

// moment-of-inertia terms of one pixel: (r, c) are window coordinates recovered from the entry's offset inside the window
template <typename PX>
__device__ __forceinline__ void accumulate_moi(Acc& a, u32 v, u32 off, u32 inv_pitch, u32 pitchB) {
  // off = r * pitchB + c * sizeof(PX), off < 64 * pitchB: r = off / pitchB by a reciprocal multiply, inv = 2^32 / pitchB + 1
  // (exact for pitchB = 16, 32 ... 144: checked exhaustively in tests/test_host_logic.py)
  const u32 r = __umulhi(off, inv_pitch);
  const u32 c = (off - r * pitchB) >> (sizeof(PX) == 1 ? 0 : 1);
  a.m10 += v * c;
  a.m01 += v * r;
  a.q += (u64)v * (u64)(c * c + r * r);
}

// One pass over the object's n list entries (TMEM, shared addresses of its pixels inside the resident window): blocks
// of 256 entries — four TMEM columns, eight entries per lane — unpredicated, the last partial block predicated.  The
// TMEM load of block i + 1 is issued before block i is added up.
template <typename PX, bool kMoi>
__device__ __forceinline__ void sweep_list(Acc& a, u32 tlist, u32 n, u32 hbase, u32 win_base, u32 inv_pitch, u32 pitchB) {
  const u32 lane = lane_id();
  const u32 full = n >> 8;
  u32 w[4];
  if (full) {
    tmem_ld4(tlist, w);
    tmem_wait_ld();
  }
#pragma unroll 1
  for (u32 blk = 0; blk < full; ++blk) {
    u32 k[8], v[8];
#pragma unroll
    for (int u = 0; u < 4; ++u) { k[2 * u] = w[u] & 0xFFFFu; k[2 * u + 1] = w[u] >> 16; }
    if (blk + 1 < full) tmem_ld4(tlist + 4u * (blk + 1u), w);  // (in flight under the arithmetic below)
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = lds_px<PX>(k[u]);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      accumulate<PX>(a, v[u], hbase);
      if (kMoi) accumulate_moi<PX>(a, v[u], k[u] - win_base, inv_pitch, pitchB);
    }
    tmem_wait_ld();
  }
  const u32 rem = n & 255u;
  if (rem) {
    tmem_ld4(tlist + 4u * full, w);
    tmem_wait_ld();
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (lane + 32u * (u32)u < rem) {
        const u32 k = (u & 1) ? (w[u >> 1] >> 16) : (w[u >> 1] & 0xFFFFu);
        const u32 v = lds_px<PX>(k);
        accumulate<PX>(a, v, hbase);
        if (kMoi) accumulate_moi<PX>(a, v, k - win_base, inv_pitch, pitchB);
      }
    }
  }
}

// Zero the relative bins [lane * per, (lane + 1) * per) of the circular histogram: together the 32 lanes clear every
// bin a request with this `per` can have touched.
__device__ __forceinline__ void zero_touched(u32* h, u32 rot, u32 per) {
  const u32 b0 = lane_id() * per;
#pragma unroll
  for (u32 k = 0; k < 32u; k += 4u)
    if (k < per) *reinterpret_cast<uint4*>(h + ((b0 + k + rot) & 1023u)) = make_uint4(0, 0, 0, 0);
}

// Locate four ranks in the circular histogram: relative bin j lives at h[(j + rot) & 1023], rot a multiple of 4, bins
// [nb, 1024) relative are zero.  Same scheme and outputs as find_ranks32 (warp_common.cuh); kZero: the touched bins are
// zeroed on the way out.
template <bool kZero>
__device__ __forceinline__ void find_ranks_rot(u32* h, u32 rot, u32 nb, const u32 (&ranks)[4], u32* t) {
  const u32 lane = lane_id();
  const u32 per = bins_per_lane(nb);
  const u32 b0 = lane * per;
  u32 cnt = 0, cb = 0;
#pragma unroll 1
  for (u32 k = 0; k < per; k += 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(h + ((b0 + k + rot) & 1023u));
    const uint4 w = *reinterpret_cast<const uint4*>(h + ((b0 + k + 4u + rot) & 1023u));
    const u32 sv = v.x + v.y + v.z + v.w, sw = w.x + w.y + w.z + w.w;
    cnt += sv + sw;
    cb += (b0 + k) * (sv + sw) + v.y + 2u * v.z + 3u * v.w + 4u * sw + w.y + 2u * w.z + 3u * w.w;
  }
  u32 icnt = cnt, icb = cb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 c = __shfl_up_sync(kFull, icnt, o);
    const u32 q = __shfl_up_sync(kFull, icb, o);
    if (lane >= (u32)o) { icnt += c; icb += q; }
  }
  const u32 ecnt = icnt - cnt, ecb = icb - cb;
  const u32 grp = lane >> 3, sub = lane & 7u;
  u32 own = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 o = (u32)__ffs(__ballot_sync(kFull, ranks[j] >= ecnt && ranks[j] < ecnt + cnt)) - 1u;
    if (grp == (u32)j) own = o;
  }
  const u32 tr = grp == 0 ? ranks[0] : (grp == 1 ? ranks[1] : (grp == 2 ? ranks[2] : ranks[3]));
  const u32 e = __shfl_sync(kFull, ecnt, own), eb = __shfl_sync(kFull, ecb, own);
  const u32 nper = per >> 3;              // bins per lane of the group: 1..4
  const u32 lb = own * per + sub * nper;  // first (relative) bin of this lane
  u32 c4[4], lc = 0, lq = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c4[k] = ((u32)k < nper) ? h[(lb + k + rot) & 1023u] : 0u;
    lc += c4[k];
    lq += c4[k] * (lb + k);
  }
  __syncwarp();
  if (kZero) zero_touched(h, rot, per);
  u32 ic = lc, iq = lq;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const u32 a = __shfl_up_sync(kFull, ic, o, 8);
    const u32 b = __shfl_up_sync(kFull, iq, o, 8);
    if (sub >= (u32)o) { ic += a; iq += b; }
  }
  u32 acc = e + ic - lc, accq = eb + iq - lq;
  if (tr >= acc && tr < acc + lc) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (tr >= acc && tr < acc + c4[k]) {
        t[grp] = lb + k;        // key (relative bin)
        t[4 + grp] = tr - acc;  // rank inside the bin
        t[8 + grp] = acc;       // count below
        t[12 + grp] = accq;     // sum(count * relative bin) below
      }
      acc += c4[k];
      accq += c4[k] * (lb + k);
    }
  }
  __syncwarp();
}

struct Ranked {
  u32 med_lo, med_hi;
  u64 top2p5_sum, top5_sum;
};

// Exact order statistics by radix selection on the resident window, for value ranges the 1024-bin histogram does not
// resolve: coarse histogram of (x - lo) >> s0, then 7 more bits per sweep inside the four target bins.  x = xf(v): the
// pixel value itself, or floor(|2 v - med2| / 2) for the MAD.  key[j] = the order statistic of rank ranks[j], minus lo.
// Leaves the histogram clean.
struct XfIdentity {
  __device__ __forceinline__ u32 operator()(u32 v) const { return v; }
};
struct XfAbsDev {
  u32 med2;
  __device__ __forceinline__ u32 operator()(u32 v) const {
    const int d = (int)(2u * v) - (int)med2;
    return (u32)(d < 0 ? -d : d) >> 1;
  }
};

template <typename PX, typename Xf>
__device__ __forceinline__ void wide_select(u32 tlist, u32 n, u32 lo, u32 hi, const Xf& xf, const u32 (&ranks)[4], u32* hist,
                                            u32* t, u32 (&key)[4]) {
  const u32 range = hi - lo;
  int s0 = 0;
  while ((range >> s0) >= 1024u) ++s0;
  const u32 nb = (range >> s0) + 1;
  __syncwarp();
  hist_zero(hist, 1024u);
  __syncwarp();
  for_each_entry(tlist, n, [&](u32 k, u32) { hist_add(hist, (xf(lds_px<PX>(k)) - lo) >> s0); });
  __syncwarp();
  find_ranks32(hist, nb, ranks, t);
  int cur = s0;
#pragma unroll
  for (int j = 0; j < 4; ++j) key[j] = t[j];
#pragma unroll 1
  while (cur > 0) {
    const int nxt = cur > 7 ? cur - 7 : 0;
    const u32 nsub = 1u << (cur - nxt);
    __syncwarp();
    hist_zero(hist, 512u);
    __syncwarp();
    for_each_entry(tlist, n, [&](u32 k, u32) {
      const u32 d = xf(lds_px<PX>(k)) - lo;
      const u32 hi_bits = d >> cur;
      const u32 sb = (d >> nxt) & (nsub - 1u);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (hi_bits == key[j]) hist_add(hist, 128u * j + sb);
    });
    __syncwarp();
    find_ranks32_x4(hist, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) key[j] = (key[j] << (cur - nxt)) | t[j];
    cur = nxt;
  }
  __syncwarp();
  hist_zero(hist, 1024u);
  __syncwarp();
}

// The four ranks of the cell functions (two medians, top 2.5 %, top 5) of a wide-range request.
template <typename PX>
__device__ __forceinline__ Ranked wide_ranks(u32 tlist, u32 n, u32 vmin, u32 vmax, u64 sum, u32 feats, u32* hist, u32* t,
                                             const u32 (&ranks)[4]) {
  u32 key[4];
  wide_select<PX>(tlist, n, vmin, vmax, XfIdentity(), ranks, hist, t, key);
  Ranked r;
  r.med_lo = vmin + key[0]; r.med_hi = vmin + key[1];
  const u32 v2 = vmin + key[2], v3 = vmin + key[3];
  u64 below2 = 0, below3 = 0;
  if (feats & (ABX_F_TOP2P5 | ABX_F_TOP5)) {
    u64 sb2 = 0, sb3 = 0;
    u32 cb2 = 0, cb3 = 0;
    for_each_entry(tlist, n, [&](u32 k, u32) {
      const u32 x = lds_px<PX>(k);
      if (x < v2) { sb2 += x; ++cb2; }
      if (x < v3) { sb3 += x; ++cb3; }
    });
    sb2 = warp_sum64(sb2); sb3 = warp_sum64(sb3);
    cb2 = __reduce_add_sync(kFull, cb2); cb3 = __reduce_add_sync(kFull, cb3);
    below2 = sb2 + (u64)(ranks[2] - cb2) * (u64)v2;
    below3 = sb3 + (u64)(ranks[3] - cb3) * (u64)v3;
  }
  r.top2p5_sum = sum - below2;
  r.top5_sum = sum - below3;
  __syncwarp();
  return r;
}


// Four warps per CTA (one TMEM lane quarter each), four CTAs per SM: sixteen objects in flight per SM, and every
// shared address of a CTA stays below 2^16 (1 KB reserved + 56 KB), so that a list entry can be the 16-bit shared
// address of its pixel.
constexpr int kSwWarps = 4, kSwCtasPerSm = 4;
constexpr int kMaxRequests = 64;   // request table in shared memory
constexpr u32 kSwSmem = 57344;     // 4 x (56 KB + 1 KB reserved) = the 228 KB of an SM
// head: request table 1 KB | per warp 512 B: scratch u32[16], mbarrier (@64), row table u32[65] (@128)
constexpr u32 kWarpHead = 512, kSwHead = 1024 + kSwWarps * kWarpHead;
constexpr u32 kMaxWindow = 64 * 144;  // 64 rows of at most 144 bytes

struct SweepMaps {
  // the pixel buffer as rows of row_stride elements; map [i][j]: box of 16 (i + 1) bytes x 8 (j + 1) rows, so that one
  // copy brings a whole window.  Up to 144 bytes wide: 64 columns of the bounding box plus the columns between the
  // 16-byte aligned start of the box and the bounding box.
  CUtensorMap px[9][8];
};

struct ReqEntry {   // one request this kernel computes
  int row_off;      // channel * rows per channel: added to the window's box row
  u32 features;
  int q;            // index into the caller's request list
  int pad_;
};

template <typename PX>
__global__ void __launch_bounds__(kSwWarps * 32, kSwCtasPerSm)
object_sweep(const __grid_constant__ SweepMaps maps, const ObjPlan* __restrict__ plan, const int* __restrict__ order,
             const u32* __restrict__ order_counts /* [0] big, [1] small */, int order_cap, u32* __restrict__ work_counter,
             const u64* __restrict__ bitmaps, int chan_rows, const abx_request* __restrict__ requests, int n_requests,
             ChanStats* __restrict__ chan, int split_log2) {
  const u32 lane = lane_id();
  const u32 warp = threadIdx.x >> 5;
  const u32 sbase = smem_addr_of(dyn);
  // ---- shared memory: head | histograms (4 KB each, 4 KB aligned in the shared window) | one window per warp ----
  ReqEntry* rtab = reinterpret_cast<ReqEntry*>(dyn);
  int* n_valid_p = reinterpret_cast<int*>(dyn + kMaxRequests * sizeof(ReqEntry) - 16);  // (the table holds < 63 entries)
  u32* tmem_base_p = reinterpret_cast<u32*>(dyn + kMaxRequests * sizeof(ReqEntry) - 12);
  const u32 hist0 = (sbase + kSwHead + 4095u) & ~4095u;
  const u32 flex0 = hist0 + kSwWarps * 4096u;
  const u32 flex_bytes = ((sbase + kSwSmem - flex0) / kSwWarps) & ~127u;
  const u32 hbase = hist0 + warp * 4096u;
  const u32 win_base = flex0 + warp * flex_bytes;
  u32* hist = reinterpret_cast<u32*>(dyn + (hbase - sbase));
  u32* t = reinterpret_cast<u32*>(dyn + 1024u + warp * kWarpHead);
  const u32 bar = sbase + 1024u + warp * kWarpHead + 64u;
  u32* rowinfo = reinterpret_cast<u32*>(dyn + 1024u + warp * kWarpHead + 128u);  // [65]: list start | row address << 16
  if (threadIdx.x == 0) {  // the requests this kernel computes (div requests belong to object_float.cu)
    int nv = 0;
    for (int q = 0; q < n_requests; ++q) {
      const abx_request rq = requests[q];
      if (rq.reduction == ABX_RED_DIV) continue;
      ReqEntry e;
      e.row_off = rq.channel * chan_rows; e.features = rq.features; e.q = q; e.pad_ = 0;
      rtab[nv++] = e;
    }
    *n_valid_p = nv;
  }
  if (warp == 0) {  // tensor memory for the four pixel lists of this CTA
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_of(tmem_base_p)), "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  hist_zero(hist, 1024u);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int n_valid = *n_valid_p;
  const u32 tmem_base = *tmem_base_p;
  const u32 tlist = tmem_base + ((warp * 32u) << 16);  // this warp's lane quarter, column 0
  u32 parity = 0;
  if (flex_bytes < kMaxWindow) __trap();  // (layout constants out of step)

  const u32 n_big = order_counts[0], n_small = order_counts[1];
  const int n_items = (int)(n_big + n_small) << split_log2;
  const int split = 1 << split_log2;
  auto fetch = [&]() {
    int v = 0;
    if (lane == 0) v = (int)atomicAdd(work_counter, 1u);
    return __shfl_sync(kFull, v, 0);
  };
  // work item -> object: the big objects from the front of the order array, then the others from its back
  auto object_of = [&](int item) {
    const u32 idx = (u32)(item >> split_log2);
    return order[idx < n_big ? (int)idx : order_cap - 1 - (int)(idx - n_big)];
  };
  auto item_requests = [&](int item, int& lo, int& hi) {  // this item's share of the request table
    const int part = item & (split - 1);
    lo = (part * n_valid) >> split_log2;
    hi = ((part + 1) * n_valid) >> split_log2;
  };
  // One copy = the whole window of one request: a single box of h8 rows.
  auto issue = [&](const Geo& g, int row_off) {  // every lane is done with the window (syncwarp by the caller)
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, g.h8 * g.pitchB);
      tma_box_2d(win_base, &maps.px[(g.pitchB >> 4) - 1u][(g.h8 >> 3) - 1u], g.tma_x, g.tma_y + row_off, bar);
    }
  };
  auto prefetch = [&](const Geo& g, int row_off) {  // the same box, into L2
#ifndef ABX_NO_PREFETCH
    if (lane == 0) tma_prefetch_2d(&maps.px[(g.pitchB >> 4) - 1u][(g.h8 >> 3) - 1u], g.tma_x, g.tma_y + row_off);
#endif
  };

  int item = fetch();
  int nxt_item = item < n_items ? fetch() : n_items;
  Geo g, gn;
  int i_lo = 0, i_hi = 0, ni_lo = 0, ni_hi = 0;
  bool have = item < n_items;
  if (have) {
    const int obj = object_of(item);
    g = make_geo<PX>(plan[obj], obj);
    item_requests(item, i_lo, i_hi);
    if (i_lo < i_hi) issue(g, rtab[i_lo].row_off);
  }
  while (have) {
    // the warp's next object (its first window is requested by this object's last sweep)
    const bool have_next = nxt_item < n_items;
    if (have_next) {
      const int nobj = object_of(nxt_item);
      gn = make_geo<PX>(plan[nobj], nobj);
      item_requests(nxt_item, ni_lo, ni_hi);
    }
    const bool next_has = have_next && ni_lo < ni_hi;
    const int after_item = have_next ? fetch() : n_items;

    if (i_lo < i_hi) {
      // ---- pixel list from the torus bitmap, into TMEM (while the first window is in flight) ----
      {
        const u64* bm = bitmaps + (size_t)g.obj * 64u;
        u64 m0 = bm[(g.row0 + lane) & 63u], m1 = bm[(g.row0 + lane + 32u) & 63u];
        m0 = (m0 >> g.rot) | (g.rot ? (m0 << (64u - g.rot)) : 0ull);
        m1 = (m1 >> g.rot) | (g.rot ? (m1 << (64u - g.rot)) : 0ull);
        const u32 c0 = (u32)__popcll(m0), c1 = (u32)__popcll(m1);
        u32 i0 = c0, i1 = c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 a = __shfl_up_sync(kFull, i0, o), b = __shfl_up_sync(kFull, i1, o);
          if (lane >= (u32)o) { i0 += a; i1 += b; }
        }
        const u32 tot0 = __shfl_sync(kFull, i0, 31);
        const u32 base0 = i0 - c0, base1 = tot0 + i1 - c1;  // list position of the first pixel of rows lane, lane + 32
        const u32 a0 = m0 ? (u32)__ffsll((long long)m0) - 1u : 0u, a1 = m1 ? (u32)__ffsll((long long)m1) - 1u : 0u;
        const u64 run0 = m0 >> a0, run1 = m1 >> a1;
        const bool single = __all_sync(kFull, ((run0 & (run0 + 1ull)) == 0ull) && ((run1 & (run1 + 1ull)) == 0ull));
        constexpr u32 es = (u32)sizeof(PX);
        const u32 cols = (g.n + 63u) >> 6;
        __syncwarp();
        if (single) {
          // One run per row (convex cells).  The lanes own ROWS (two each) but a TMEM lane holds the list positions
          // congruent to it, so the entries take one hop through shared memory: the row owners scatter them — position p
          // of the list at staging[p] — into the (clean) histogram area, 2048 positions per round, and every lane
          // then collects its own positions into TMEM words.
          const u32 f0 = win_base + lane * g.pitchB + (a0 + g.s_px) * es;          // first pixel of row `lane`
          const u32 f1 = win_base + (lane + 32u) * g.pitchB + (a1 + g.s_px) * es;  // and of row `lane + 32`
          const u32 it0 = __reduce_max_sync(kFull, c0), it1 = g.h > 32u ? __reduce_max_sync(kFull, c1) : 0u;
#pragma unroll 1
          for (u32 lo = 0; lo < g.n; lo += 2048u) {
            const u32 hi = min(g.n, lo + 2048u);
            // staging address of the row's first entry (may lie below or above the staged range: unsigned compare)
            u32 p = 2u * (base0 - lo), v = f0;
#pragma unroll 4
            for (u32 it = 0; it < it0; ++it) {
              if (it < c0 && p < 4096u) sts_u16(hbase + p, v);
              p += 2u; v += es;
            }
            p = 2u * (base1 - lo); v = f1;
#pragma unroll 4
            for (u32 it = 0; it < it1; ++it) {
              if (it < c1 && p < 4096u) sts_u16(hbase + p, v);
              p += 2u; v += es;
            }
            __syncwarp();
            const u32 c_end = (hi + 63u) >> 6;
#pragma unroll 1
            for (u32 c = lo >> 6; c < c_end; c += 4) {
              u32 w[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const u32 q = hbase + 2u * (64u * (c + (u32)u) + lane - lo);  // (columns past the list read zeros or stale
                w[u] = (q < hbase + 4032u) ? (lds_u16(q) | (lds_u16(q + 64u) << 16)) : 0u;  //  entries: never used)
              }
              tmem_st4(tlist + c, w);
            }
            __syncwarp();
            // the staging area becomes a clean histogram again
            for (u32 k = lane; k < ((2u * (hi - lo) + 15u) >> 4); k += 32u)
              *reinterpret_cast<uint4*>(hist + 4u * k) = make_uint4(0, 0, 0, 0);
            __syncwarp();
          }
        } else {
          // Any shape: the row masks go through the (clean) histogram area, the column of a list position is the
          // (e - start)-th set bit of its row's mask.
          hist[2u * lane] = (u32)m0; hist[2u * lane + 1u] = (u32)(m0 >> 32);
          hist[2u * (lane + 32u)] = (u32)m1; hist[2u * (lane + 32u) + 1u] = (u32)(m1 >> 32);
          rowinfo[lane] = base0; rowinfo[lane + 32] = base1;
          if (lane == 0) rowinfo[64] = 0xFFFFu;
          __syncwarp();
          u32 r = 0, start = rowinfo[0], nx = rowinfo[1];
#pragma unroll 1
          for (u32 c = 0; c < cols; c += 4) {
            u32 w[4];
#pragma unroll 1
            for (int u = 0; u < 4; ++u) {
              u32 e = 64u * (c + (u32)u) + lane, pair = 0;
#pragma unroll 1
              for (int hf = 0; hf < 2; ++hf, e += 32u) {
                while (e >= nx) { start = nx; ++r; nx = rowinfo[r + 1u]; }
                u32 addr = 0;
                if (e < g.n) {
                  const u32 lo = hist[2u * r], hi = hist[2u * r + 1u], nth = e - start, pl = (u32)__popc(lo);
                  const u32 col = nth < pl ? __fns(lo, 0u, (int)nth + 1) : 32u + __fns(hi, 0u, (int)(nth - pl) + 1);
                  addr = win_base + r * g.pitchB + (col + g.s_px) * es;
                }
                pair |= (addr & 0xFFFFu) << (16 * hf);
              }
              w[u] = pair;
            }
            tmem_st4(tlist + c, w);
          }
          __syncwarp();
          *reinterpret_cast<uint4*>(hist + 4u * lane) = make_uint4(0, 0, 0, 0);  // the 128 words of the masks
        }
        tmem_wait_st();
        __syncwarp();
      }
      const u32 inv_pitch = 0xFFFFFFFFu / g.pitchB + 1u;
      const u32 k2p5 = (u32)ceil((double)g.n * 0.025);  // int(np.ceil(n * 0.025)), cell.py:110-111
      const u32 ranks[4] = {(g.n - 1) / 2, g.n / 2, g.n - k2p5, g.n - min(g.n, 5u)};

#pragma unroll 1
      for (int i = i_lo; i < i_hi; ++i) {
        const ReqEntry rq = rtab[i];
        Acc a;
        a.sum = a.wh = a.vmax = a.m10 = a.m01 = 0; a.vmin = kFull; a.sq = a.q = 0;
        const bool want_moi = (rq.features & ABX_F_MOI) != 0;
        const bool want_ranks = (rq.features & (ABX_F_MEDIAN | ABX_F_TOP2P5 | ABX_F_TOP5)) != 0;
        const bool want_cp = (rq.features & (ABX_F_CPQ | ABX_F_CPMAD)) != 0;  // cp_measure `intensity` order statistics
        Ranked rk;
        rk.med_lo = rk.med_hi = 0; rk.top2p5_sum = rk.top5_sum = 0;
        u32 cpq[6] = {0, 0, 0, 0, 0, 0}, mad_lo = 0, mad_hi = 0, maxpos = 0;
        bool ranks_done = false;  // order statistics found and histogram cleaned before the next copy was issued
        u64 sum64 = 0;
        mbar_wait(bar, parity);
        parity ^= 1u;
        if (want_moi) sweep_list<PX, true>(a, tlist, g.n, hbase, win_base, inv_pitch, g.pitchB);
        else sweep_list<PX, false>(a, tlist, g.n, hbase, win_base, inv_pitch, g.pitchB);
        __syncwarp();
        const u32 vmin = __reduce_min_sync(kFull, a.vmin), vmax = __reduce_max_sync(kFull, a.vmax);
        const bool wide = (want_ranks || want_cp) && vmax - (vmin & ~3u) > 1023u;
        if (want_cp) {
          // ---- cp_measure `intensity`: every order statistic now, while the window is resident (the MAD needs
          // a second pass over it) ----
          sum64 = (u64)__reduce_add_sync(kFull, a.sum);
          const u32 n = g.n, last = n - 1u;
          const u32 i1 = n >> 2, i2 = n >> 1, i3 = (3u * n) >> 2;  // floor(n f): CellProfiler's rank rule
          const u32 rb1[4] = {i1, min(i1 + 1u, last), i2, min(i2 + 1u, last)};
          const u32 rb2[4] = {i3, min(i3 + 1u, last), i3, min(i3 + 1u, last)};
          const u32 vbase = vmin & ~3u, rot = vbase & 1023u, nb = vmax - vbase + 1u;
          if (!wide) {
            if (want_ranks) {
              find_ranks_rot<false>(hist, rot, nb, ranks, t);
              rk.med_lo = vbase + t[0]; rk.med_hi = vbase + t[1];
              const u32 v2 = vbase + t[2], v3 = vbase + t[3];
              rk.top2p5_sum = sum64 - ((u64)vbase * t[10] + t[14] + (u64)t[6] * v2);
              rk.top5_sum = sum64 - ((u64)vbase * t[11] + t[15] + (u64)t[7] * v3);
            }
            find_ranks_rot<false>(hist, rot, nb, rb1, t);
#pragma unroll
            for (int j = 0; j < 4; ++j) cpq[j] = vbase + t[j];
            find_ranks_rot<true>(hist, rot, nb, rb2, t);
            cpq[4] = vbase + t[0]; cpq[5] = vbase + t[1];
          } else {
            if (want_ranks) rk = wide_ranks<PX>(tlist, n, vmin, vmax, sum64, rq.features, hist, t, ranks);
            u32 key[4];
            wide_select<PX>(tlist, n, vmin, vmax, XfIdentity(), rb1, hist, t, key);
#pragma unroll
            for (int j = 0; j < 4; ++j) cpq[j] = vmin + key[j];
            wide_select<PX>(tlist, n, vmin, vmax, XfIdentity(), rb2, hist, t, key);
            cpq[4] = vmin + key[0]; cpq[5] = vmin + key[1];
          }
          if (rq.features & ABX_F_CPMAD) {
            // twice the median (an integer): f = 1/2 exactly when n is odd
            const u32 med2 = ((n & 1u) && i2 < last) ? cpq[2] + cpq[3] : 2u * cpq[2];
            const XfAbsDev xf{med2};
            const u32 rmad[4] = {i2, min(i2 + 1u, last), i2, min(i2 + 1u, last)};
            u32 first = kFull;  // list position of the first maximum (row-major order)
            __syncwarp();
            for_each_entry(tlist, n, [&](u32 k, u32 e) {
              const u32 v = lds_px<PX>(k);
              if (v == vmax) first = min(first, e);
              if (!wide) hist_add(hist, xf(v));  // deviations fit the histogram: <= max - min <= 1023
            });
            __syncwarp();
            u32 key[4];
            if (!wide) {
              find_ranks_rot<true>(hist, 0u, vmax - vmin + 1u, rmad, t);
              key[0] = t[0]; key[1] = t[1];
            } else {
              wide_select<PX>(tlist, n, 0u, vmax - vmin, xf, rmad, hist, t, key);
            }
            mad_lo = key[0];
            mad_hi = key[1] | ((med2 & 1u) << 31);
            first = __reduce_min_sync(kFull, first);
            // the entry at list position `first`: lane first & 31, column first >> 6, half (first >> 5) & 1
            const u32 word = __shfl_sync(kFull, tmem_ld1(tlist + (first >> 6)), first & 31u);
            const u32 off = (((first >> 5) & 1u) ? word >> 16 : word & 0xFFFFu) - win_base;
            const u32 r = __umulhi(off, inv_pitch);
            maxpos = (r << 16) | (((off - r * g.pitchB) >> (sizeof(PX) == 1 ? 0 : 1)) - g.s_px);
          }
          ranks_done = true;
        } else if (wide) {  // the window is still here: refine now, before it is overwritten
          sum64 = (u64)__reduce_add_sync(kFull, a.sum);
          rk = wide_ranks<PX>(tlist, g.n, vmin, vmax, sum64, rq.features, hist, t, ranks);
          ranks_done = true;
        }
        __syncwarp();
        // ---- the next window: next request, or the first request of the warp's next object; the window after that
        // one goes to L2 ----
        if (i + 1 < i_hi) {
          issue(g, rtab[i + 1].row_off);
          if (i + 2 < i_hi) prefetch(g, rtab[i + 2].row_off);
          else if (next_has) prefetch(gn, rtab[ni_lo].row_off);
        } else if (next_has) {
          issue(gn, rtab[ni_lo].row_off);
          if (ni_lo + 1 < ni_hi) prefetch(gn, rtab[ni_lo + 1].row_off);
        }
        // ---- reductions (under the copy that was just issued) ----
        ChanStats cs;
        cs.sum = (want_cp || wide) ? sum64 : (u64)__reduce_add_sync(kFull, a.sum);  // n * 65535 < 2^32
        cs.sumsq = warp_sum64(a.sq);
        constexpr int kShift = (sizeof(PX) == 1) ? 12 : 8;
        cs.wrapsq = cs.sumsq - ((u64)__reduce_add_sync(kFull, a.wh) << (32 - 2 * kShift));
        cs.m10 = cs.m01 = cs.m20 = cs.m02 = 0;
        if (want_moi) {
          // the sweep counted columns from the TMA box (bbox column 0 = box column s_px): back to the bbox origin
          const u64 m10w = warp_sum64((u64)a.m10), s = (u64)g.s_px;
          cs.m10 = m10w - s * cs.sum;
          cs.m01 = warp_sum64((u64)a.m01);
          cs.m20 = warp_sum64(a.q) - 2ull * s * m10w + s * s * cs.sum;  // m20 + m02 as one sum (finalize.cu)
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) cs.q[j] = cpq[j];
        cs.mad_lo = mad_lo; cs.mad_hi = mad_hi; cs.maxpos = maxpos; cs.pad_ = 0;
        cs.vmin = vmin; cs.vmax = vmax;
        cs.med_lo = rk.med_lo; cs.med_hi = rk.med_hi;
        cs.top2p5_sum = rk.top2p5_sum; cs.top5_sum = rk.top5_sum;
        if (!ranks_done) {
          const u32 vbase = vmin & ~3u;
          if (want_ranks) {
            find_ranks_rot<true>(hist, vbase & 1023u, vmax - vbase + 1u, ranks, t);
            cs.med_lo = vbase + t[0]; cs.med_hi = vbase + t[1];
            const u32 v2 = vbase + t[2], v3 = vbase + t[3];
            // sum of the smallest values up to the rank = vbase * cnt + sum(count * bin) below + rank * value
            const u64 below2 = (u64)vbase * t[10] + t[14] + (u64)t[6] * v2;
            const u64 below3 = (u64)vbase * t[11] + t[15] + (u64)t[7] * v3;
            cs.top2p5_sum = cs.sum - below2;
            cs.top5_sum = cs.sum - below3;
          } else {  // no order statistics wanted: just clean up
            const u32 span = vmax - vbase;
            zero_touched(hist, vbase & 1023u, span > 1023u ? 32u : bins_per_lane(span + 1u));
          }
        }
        if (lane == 0) chan[(i64)g.obj * n_requests + rq.q] = cs;
        __syncwarp();
      }
    } else if (next_has) {
      issue(gn, rtab[ni_lo].row_off);  // this item had no request for this kernel: start the next object's first window
    }
    g = gn; i_lo = ni_lo; i_hi = ni_hi;
    have = have_next;
    item = nxt_item;
    nxt_item = after_item;
  }
  // ---- give the tensor memory back ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}

// The tensor maps, or false when the pixel layout does not qualify for TMA (the caller then takes object_stats_warp).
bool make_maps(const abx_extract_args* a, SweepMaps* m) {
  EncodeFn encode = tensor_map_encoder();
  if (!encode || a->Z != 1 || a->pixel_elems <= 0 || a->n_requests > kMaxRequests - 2) return false;
  const size_t es = a->pixel_dtype == ABX_U8 ? 1 : 2;
  if ((reinterpret_cast<uintptr_t>(a->pixels) & 15u) || (a->row_stride * (i64)es) % 16 || a->row_stride * (i64)es < 144 ||
      a->chan_stride % a->row_stride || a->chan_stride / a->row_stride > 0x7FFFFFFF / (a->C > 0 ? a->C : 1))
    return false;
  const i64 rows = a->pixel_elems / a->row_stride;  // whole rows inside the caller's buffer
  if (rows < 64 || rows > 0x7FFFFFFF) return false;
  // the 64 encodings are cached per host thread: a pipeline calls with the same buffer at every time point
  struct Key { const void* px; i64 rs, rows; int dtype; };
  static thread_local Key ckey = {nullptr, 0, 0, -1};
  static thread_local SweepMaps cmaps;
  if (ckey.px == a->pixels && ckey.rs == a->row_stride && ckey.rows == rows && ckey.dtype == a->pixel_dtype) {
    memcpy(m, &cmaps, sizeof(SweepMaps));
    return true;
  }
  const cuuint32_t estr[2] = {1u, 1u};
  const cuuint64_t pdim[2] = {(cuuint64_t)a->row_stride, (cuuint64_t)rows};
  const cuuint64_t pstr[1] = {(cuuint64_t)a->row_stride * es};
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 8; ++j) {
      const cuuint32_t pbox[2] = {(cuuint32_t)(16u * (i + 1) / es), (cuuint32_t)(8 * (j + 1))};
      if (encode(&m->px[i][j], es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
                 const_cast<void*>(a->pixels), pdim, pstr, pbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    }
  memcpy(&cmaps, m, sizeof(SweepMaps));
  ckey = Key{a->pixels, a->row_stride, rows, a->pixel_dtype};
  return true;
}

template <typename PX>
int launch_sweep(const abx_extract_args* a, const Workspace& ws, const SweepMaps& maps, cudaStream_t st) {
  static thread_local bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(object_sweep<PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSwSmem);
    if (e != cudaSuccess) return abx_check_cuda(e, "object_sweep smem attribute");
    done[dev] = true;
  }
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  const int resident = 148 * kSwCtasPerSm * kSwWarps;
  // fewer objects than two rounds of resident warps: several work items per object, each with its share of the requests
  int split_log2 = 0;
  while (split_log2 < 3 && (n_total << split_log2) < 2 * resident && (2 << split_log2) <= a->n_requests) ++split_log2;
  int grid = ((n_total << split_log2) + kSwWarps - 1) / kSwWarps;
  if (grid > 148 * kSwCtasPerSm) grid = 148 * kSwCtasPerSm;  // persistent: warps pull objects from a counter
  object_sweep<PX><<<grid, kSwWarps * 32, kSwSmem, st>>>(maps, ws.plan, ws.order_stats, ws.list_counts + kCntOrderBig,
                                                        n_total /* the plan kernel's capacity of the order array */,
                                                        ws.list_counts + kCntSweepWork, ws.bitmaps,
                                                        (int)(a->chan_stride / a->row_stride), a->requests, a->n_requests,
                                                        ws.chan, split_log2);
  return abx_check_cuda(cudaGetLastError(), "object_sweep");
}

}  // namespace

// Whether abx_extract takes the sweep kernel for this call (layout and dtype qualify).
bool abx_sweep_ok(const abx_extract_args* a) {
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  if (n_total == 0 || a->n_requests == 0) return false;
  if (a->pixel_dtype != ABX_U16 && a->pixel_dtype != ABX_U8) return false;
  SweepMaps maps;
  memset(&maps, 0, sizeof(maps));
  return make_maps(a, &maps);
}

// Routing of every object (statistics when `sweep`, shape always): runs after the label scan.
int launch_plan(const abx_extract_args* a, const Workspace& ws, cudaStream_t st, bool sweep) {
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  if (n_total == 0 || (!sweep && !a->need_edt)) return ABX_OK;
  PlanArgs p;
  p.recs = ws.recs;
  p.plane_base = a->plane_base;
  p.plane_tile = a->plane_tile;
  p.tile_offset = reinterpret_cast<const i64*>(a->tile_offset);
  p.n_planes = a->n_planes;
  p.n_objects = a->n_objects;
  p.n_total = n_total;
  p.row_stride = a->row_stride > 0 ? a->row_stride : 1;
  p.align = a->pixel_dtype == ABX_U8 ? 16 : 8;
  p.n_requests = a->n_requests;
  p.sweep = sweep ? 1 : 0;
  p.need_edt = (a->need_edt & 3) ? 1 : 0;
  p.chan = ws.chan;
  p.shape = ws.shape;
  p.plan = ws.plan;
  p.order_stats = ws.order_stats;
  p.order_edt = ws.order_edt;
  p.stats_list = ws.stats_list;
  p.pair_list = ws.pair_list;
  p.edt_list = ws.edt_list;
  p.counts = ws.list_counts;
  p.cp_requests = (a->request_feature_union & (int)(ABX_F_CPQ | ABX_F_CPMAD)) != 0;
  p.requests = a->requests;
  p.want_moments = (a->need_edt & 4) != 0;
  p.bitmaps = ws.bitmaps;
  p.mom = ws.mom;
  p.err = ws.err;
  plan_kernel<<<(n_total + 255) / 256, 256, 0, st>>>(p);
  if (p.want_moments && a->n_objects > 0 && (a->H > kSide || a->W > kSide)) {  // cells above the window can exist
    const int grid = a->n_objects < 148 * 8 ? a->n_objects : 148 * 8;
    large_moments_kernel<<<grid, 256, 0, st>>>(ws.recs, a->n_objects, static_cast<const uint16_t*>(a->labels),
                                               a->label_plane_stride, a->label_row_stride, a->plane_base, a->n_planes, ws.mom);
  }
  return abx_check_cuda(cudaGetLastError(), "plan");
}

int launch_object_sweep(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  SweepMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (!make_maps(a, &maps)) return abx_set_error(ABX_ERR_INVALID, "object_sweep: layout does not qualify for TMA");
  if (a->pixel_dtype == ABX_U16) return launch_sweep<uint16_t>(a, ws, maps, st);
  return launch_sweep<uint8_t>(a, ws, maps, st);
}
