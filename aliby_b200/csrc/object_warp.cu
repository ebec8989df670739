// Fused per-object kernel for objects whose bounding box fits 64 x 64: ONE WARP PER OBJECT.
//
// Covers, for one object, everything the reference computes with |instructions| full-plane
// passes (src/extraction/extract.py:346-359): the intensity statistics of every
// (channel, Z-reduction) request (cell.py:43-157,232-265, distributors.py:19-21 fused into the
// load, tile crop of tiler.py:309-366 fused through the tile offset) and the three chained
// exact EDTs of the shape metrics (cell.py:176-229).  The label window is read once:
//
//   phase M  64-bit row bitmasks of the object from warp ballots (512 B), plus the compact list
//            of its pixel offsets — no atomics, deterministic order
//   phase S  per request: gather through the offset list (coalesced along rows), moments in
//            registers, values staged in shared memory, range-adaptive 512-bin histogram,
//            ranks located by one warp scan; refinement sweeps only when the value range > 511
//   phase E  row distances from the bitmasks (clz/ffs), exact column pass with early exit,
//            cone top as a second bitmask, plateau distances with lanes over rows
//
// No __syncthreads: the eight warps of a CTA work on eight different objects.  Objects with a
// larger window (and the per-plane background) are appended to work lists that the CTA-per-object
// kernels (object_stats.cu, shape_edt.cu) consume.
#include "common.cuh"

// Optional per-phase cycle counters (debug builds: -DABX_PHASE_TIMING): lane 0 of every warp adds its
// clock64() deltas; read back through abx_debug_phase_cycles().
#ifdef ABX_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8];
#define PHASE_T0() long long _pt = clock64()
#define PHASE_ADD(i)                                                                      \
  do {                                                                                    \
    const long long _n = clock64();                                                       \
    if (lane_id() == 0) atomicAdd(&g_phase_cycles[i], (unsigned long long)(_n - _pt));    \
    _pt = _n;                                                                             \
  } while (0)
#else
#define PHASE_T0() do {} while (0)
#define PHASE_ADD(i) do {} while (0)
#endif

namespace {

constexpr int kSide = 64;    // maximum window side
// Two size classes share the code: <= 2048 pixels (8 warps per CTA) and 2049..4096 pixels (4 warps
// per CTA, twice the shared memory per warp), so that every object with a window <= 64 x 64 is
// handled through the compact pixel list.
constexpr int kCapSmall = 2048, kWarpsSmall = 8;
constexpr int kCapLarge = 4096, kWarpsLarge = 4;
constexpr int kBins = 1024;  // level-0 histogram bins, 16-bit counters packed in pairs (4 x 128 during refinement)

template <int CAP>
struct __align__(16) WSmemT {
  u64 rowmask[kSide];            // bit c of rowmask[r]: window pixel (r, c) belongs to the object
  unsigned short rowbase[kSide]; // number of object pixels in rows < r
  unsigned short offs[CAP];      // compact list: (r << 6) | c
  unsigned short vals[CAP];      // staged values of the current request; u8 g[66][64] in phase E
  u32 hist[kBins / 2];           // 1024 packed 16-bit counters; g overflow + u64 topmask[64] in phase E
  u32 t_key[4], t_rank[4], t_cnt[4], t_cb[4];
};

struct Obj {
  const uint16_t* lab;  // label window origin
  i64 lab_rs;
  u32 label, n;
  int h, w;
  bool listed;          // compact offset list valid (n <= CAP of the size class)
};

__device__ __forceinline__ u64 lanemask_lt64(u32 c) { return (c == 0) ? 0ull : (~0ull >> (64 - c)); }

// f(r, c, i): every object pixel once; i = compact index.  Warp-uniform control flow around f is
// NOT guaranteed (lanes without a pixel skip f).
template <class S, class F>
__device__ __forceinline__ void for_each_px(const Obj& o, const S& s, F&& f) {
  const u32 lane = lane_id();
  if (o.listed) {
    for (u32 i = lane; i < o.n; i += 32) {
      const u32 k = s.offs[i];
      f(k >> 6, k & 63u, i);
    }
  } else {
    for (int r = 0; r < o.h; ++r) {
      const u64 m = s.rowmask[r];
      if (m == 0) continue;
      const u32 base = s.rowbase[r];
      for (u32 c = lane; c < (u32)o.w; c += 32)
        if ((m >> c) & 1ull) f((u32)r, c, base + (u32)__popcll(m & lanemask_lt64(c)));
    }
  }
}

__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// ---- histogram with packed 16-bit counters (counts <= 4096 per object) ------------------------
__device__ __forceinline__ void hist_zero(u32* hist) {  // 1024 bins = 128 x uint4
  uint4* h4 = reinterpret_cast<uint4*>(hist);
#pragma unroll
  for (int k = 0; k < kBins / 8 / 32; ++k) h4[lane_id() + 32 * k] = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void hist_add(u32* hist, u32 bin) {
  atomicAdd(&hist[bin >> 1], (bin & 1u) ? 0x10000u : 1u);
}

// Locate up to four ranks in the 16-bit histogram h16[0, nb): each lane owns `per` consecutive
// bins (a multiple of 8, read as uint4).  For rank t: key = its bin, rank = t - (count below the
// bin), cnt = count below the bin, cb = sum over the bins below of count * bin index.
__device__ __forceinline__ void find_ranks16(const unsigned short* h16, u32 nb, const u32* ranks, int n_ranks,
                                             u32* out_key, u32* out_rank, u32* out_cnt, u32* out_cb) {
  const u32 lane = lane_id();
  const u32 per = (((nb + 31u) >> 5) + 7u) & ~7u;  // 8, 16, 24 or 32
  const u32 b0 = lane * per;
  u32 cnt = 0, cb = 0;
  for (u32 k = 0; k < per; k += 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(h16 + b0 + k);
    const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const u32 lo = w[j] & 0xFFFFu, hi = w[j] >> 16;
      cnt += lo + hi;
      cb += lo * (b0 + k + 2 * j) + hi * (b0 + k + 2 * j + 1);
    }
  }
  u32 icnt = cnt, icb = cb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 c = __shfl_up_sync(0xFFFFFFFFu, icnt, o);
    const u32 q = __shfl_up_sync(0xFFFFFFFFu, icb, o);
    if (lane >= (u32)o) { icnt += c; icb += q; }
  }
  const u32 ecnt = icnt - cnt, ecb = icb - cb;
  for (int j = 0; j < n_ranks; ++j) {
    const u32 t = ranks[j];
    if (t >= ecnt && t < ecnt + cnt) {
      u32 acc = ecnt, accb = ecb;
      for (u32 bq = b0; bq < b0 + per; ++bq) {
        const u32 c = h16[bq];
        if (t < acc + c) { out_key[j] = bq; out_rank[j] = t - acc; out_cnt[j] = acc; out_cb[j] = accb; break; }
        acc += c;
        accb += c * bq;
      }
    }
  }
}

// Four independent rank searches at once: eight lanes per 128-bin sub-histogram (refinement).
__device__ __forceinline__ void find_ranks16_x4(const unsigned short* h16, const u32* ranks, u32* out_key, u32* out_rank) {
  const u32 lane = lane_id();
  const u32 grp = lane >> 3, sub = lane & 7u;
  const u32 b0 = grp * 128u + sub * 16u;
  u32 cnt = 0;
#pragma unroll
  for (int k = 0; k < 16; k += 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(h16 + b0 + k);
    cnt += (v.x & 0xFFFFu) + (v.x >> 16) + (v.y & 0xFFFFu) + (v.y >> 16) + (v.z & 0xFFFFu) + (v.z >> 16) +
           (v.w & 0xFFFFu) + (v.w >> 16);
  }
  u32 icnt = cnt;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const u32 c = __shfl_up_sync(0xFFFFFFFFu, icnt, o, 8);
    if (sub >= (u32)o) icnt += c;
  }
  const u32 ecnt = icnt - cnt;
  const u32 t = ranks[grp];
  if (t >= ecnt && t < ecnt + cnt) {
    u32 acc = ecnt;
    for (u32 bq = b0; bq < b0 + 16u; ++bq) {
      const u32 c = h16[bq];
      if (t < acc + c) { out_key[grp] = bq - grp * 128u; out_rank[grp] = t - acc; break; }
      acc += c;
    }
  }
}

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

// ------------------------------------------------------------------------------------------------
// phase S: one (channel, reduction) request
// ------------------------------------------------------------------------------------------------
template <typename PX, class S>
__device__ __forceinline__ void request_stats(const Obj& o, S& s, const PX* __restrict__ px, i64 px_rs,
                                              i64 z_stride, int Z, const abx_request rq, u32 feats,
                                              ChanStats* __restrict__ dst) {
  const u32 lane = lane_id();
  constexpr u32 kWrapMask = (sizeof(PX) == 1) ? 0xFFu : 0xFFFFu;
  const bool add = rq.reduction == ABX_RED_ADD;
  const bool staged = o.listed && !add;  // values fit u16 and the list exists
  const bool want_moi = (feats & ABX_F_MOI) != 0;
  const u32 n = o.n;

  // ---- pass 1: moments and extrema; stage values ----
  PHASE_T0();
  ChanStats cs;
  u32 a_min = 0xFFFFFFFFu, a_max = 0;
  if (staged) {
    // Fast path: at most CAP / 32 <= 128 values per lane, each < 2^16, so 32-bit partial sums of
    // x, x*c, x*r and (x*x mod 2^16) are exact and the squares go through one IMAD.WIDE each.
    // Gathers are issued in batches of kBatch per lane so that their round trips overlap.
    constexpr int kBatch = 8;
    u32 f_sum = 0, f_wrap = 0, f_m10 = 0, f_m01 = 0;
    u64 f_sq = 0, f_m20 = 0, f_m02 = 0;
    const u32 rs = (u32)px_rs;
    for (u32 i0 = lane; i0 < n; i0 += 32 * kBatch) {
      u32 k[kBatch], x[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const u32 i = i0 + 32u * u;
        k[u] = (i < n) ? (u32)s.offs[i] : 0xFFFFu;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        x[u] = 0;
        if (k[u] != 0xFFFFu) {
          const PX* q = px + ((k[u] >> 6) * rs + (k[u] & 63u));
          x[u] = (Z == 1) ? (u32)__ldg(q) : load_reduced(q, Z, z_stride, rq.reduction);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const bool ok = k[u] != 0xFFFFu;
        const u32 v = x[u];
        f_sum += v;
        f_sq += (u64)v * (u64)v;
        f_wrap += (v * v) & kWrapMask;
        a_min = min(a_min, ok ? v : 0xFFFFFFFFu);
        a_max = max(a_max, v);
        if (want_moi) {
          const u32 c = k[u] & 63u, r = (k[u] >> 6) & 63u;
          const u32 xc = v * c, xr = v * r;
          f_m10 += xc; f_m01 += xr;
          f_m20 += (u64)xc * (u64)c; f_m02 += (u64)xr * (u64)r;
        }
        if (ok) s.vals[i0 + 32u * u] = (unsigned short)v;
      }
    }
    cs.sum = (u64)__reduce_add_sync(0xFFFFFFFFu, f_sum);      // n * 65535 < 2^27
    cs.wrapsq = (u64)__reduce_add_sync(0xFFFFFFFFu, f_wrap);
    cs.sumsq = warp_sum64(f_sq);
    if (want_moi) {
      cs.m10 = warp_sum64((u64)f_m10); cs.m01 = warp_sum64((u64)f_m01);
      cs.m20 = warp_sum64(f_m20); cs.m02 = warp_sum64(f_m02);
    } else {
      cs.m10 = cs.m01 = cs.m20 = cs.m02 = 0;
    }
  } else {
    u64 a_sum = 0, a_sq = 0, a_wrap = 0, a_m10 = 0, a_m01 = 0, a_m20 = 0, a_m02 = 0;
    for_each_px(o, s, [&](u32 r, u32 c, u32) {
      const u32 x = load_reduced(px + (i64)r * px_rs + c, Z, z_stride, rq.reduction);
      a_sum += x;
      const u64 xx = (u64)x * (u64)x;
      a_sq += xx;
      a_wrap += add ? xx : (u64)((u32)xx & kWrapMask);
      a_min = min(a_min, x);
      a_max = max(a_max, x);
      if (want_moi) {
        a_m10 += (u64)x * c; a_m01 += (u64)x * r;
        a_m20 += (u64)x * c * c; a_m02 += (u64)x * r * r;
      }
    });
    cs.sum = warp_sum64(a_sum);
    cs.sumsq = warp_sum64(a_sq);
    cs.wrapsq = warp_sum64(a_wrap);
    if (want_moi) {
      cs.m10 = warp_sum64(a_m10); cs.m01 = warp_sum64(a_m01);
      cs.m20 = warp_sum64(a_m20); cs.m02 = warp_sum64(a_m02);
    } else {
      cs.m10 = cs.m01 = cs.m20 = cs.m02 = 0;
    }
  }
  const u32 vmin = __reduce_min_sync(0xFFFFFFFFu, a_min);
  const u32 vmax = __reduce_max_sync(0xFFFFFFFFu, a_max);
  PHASE_ADD(1);
  cs.vmin = vmin; cs.vmax = vmax;
  cs.med_lo = cs.med_hi = 0;
  cs.top2p5_sum = cs.top5_sum = 0;

  if (feats & (ABX_F_MEDIAN | ABX_F_TOP2P5 | ABX_F_TOP5)) {
    // values again: from shared memory when staged, else re-gathered
    auto for_each_value = [&](auto&& f) {
      if (staged) {
        for (u32 i = lane; i < n; i += 32) f((u32)s.vals[i]);
      } else {
        for_each_px(o, s, [&](u32 r, u32 c, u32) { f(load_reduced(px + (i64)r * px_rs + c, Z, z_stride, rq.reduction)); });
      }
    };
    // ---- pass 2: range-adaptive histogram, 1024 bins of packed 16-bit counters ----
    const unsigned short* h16 = reinterpret_cast<const unsigned short*>(s.hist);
    const u32 range = vmax - vmin;
    int s0 = 0;
    while ((range >> s0) >= (u32)kBins) ++s0;
    const u32 nb = (range >> s0) + 1;
    __syncwarp();
    hist_zero(s.hist);
    __syncwarp();
    if (staged) {
      u32 i = lane;
      for (; i + 96 < n; i += 128) {  // four independent atomics in flight
        const u32 x0 = s.vals[i], x1 = s.vals[i + 32], x2 = s.vals[i + 64], x3 = s.vals[i + 96];
        hist_add(s.hist, (x0 - vmin) >> s0); hist_add(s.hist, (x1 - vmin) >> s0);
        hist_add(s.hist, (x2 - vmin) >> s0); hist_add(s.hist, (x3 - vmin) >> s0);
      }
      for (; i < n; i += 32) hist_add(s.hist, ((u32)s.vals[i] - vmin) >> s0);
    } else {
      for_each_value([&](u32 x) { hist_add(s.hist, (x - vmin) >> s0); });
    }
    __syncwarp();
    PHASE_ADD(2);
    const u32 k2p5 = (u32)ceil((double)n * 0.025);  // int(np.ceil(n * 0.025)), cell.py:110-111
    const u32 k5 = min(n, 5u);
    const u32 ranks[4] = {(n - 1) / 2, n / 2, n - k2p5, n - k5};
    find_ranks16(h16, nb, ranks, 4, s.t_key, s.t_rank, s.t_cnt, s.t_cb);
    __syncwarp();
    PHASE_ADD(3);
    u32 value[4];
    u64 below[2];
    if (s0 == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) value[j] = vmin + s.t_key[j];
      // sum of the t smallest values = vmin * cnt + sum(count * bin) over the bins below + rank * value
      below[0] = (u64)vmin * s.t_cnt[2] + s.t_cb[2] + (u64)s.t_rank[2] * value[2];
      below[1] = (u64)vmin * s.t_cnt[3] + s.t_cb[3] + (u64)s.t_rank[3] * value[3];
    } else {
      // ---- refinement: 7 more bits per sweep inside the four target bins, searched in parallel ----
      int cur = s0;
      u32 key[4], rnk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { key[j] = s.t_key[j]; rnk[j] = s.t_rank[j]; }
      while (cur > 0) {
        const int nxt = cur > 7 ? cur - 7 : 0;
        const u32 nsub = 1u << (cur - nxt);
        __syncwarp();
        hist_zero(s.hist);
        __syncwarp();
        for_each_value([&](u32 x) {
          const u32 d = x - vmin;
          const u32 hi = d >> cur;
          const u32 sb = (d >> nxt) & (nsub - 1u);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (hi == key[j]) hist_add(s.hist, 128u * j + sb);
        });
        __syncwarp();
        find_ranks16_x4(h16, rnk, s.t_key, s.t_rank);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) { key[j] = (key[j] << (cur - nxt)) | s.t_key[j]; rnk[j] = s.t_rank[j]; }
        cur = nxt;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) value[j] = vmin + key[j];
      below[0] = below[1] = 0;
      if (feats & (ABX_F_TOP2P5 | ABX_F_TOP5)) {
        u64 sb2 = 0, sb3 = 0;
        u32 cb2 = 0, cb3 = 0;
        const u32 v2 = value[2], v3 = value[3];
        for_each_value([&](u32 x) {
          if (x < v2) { sb2 += x; ++cb2; }
          if (x < v3) { sb3 += x; ++cb3; }
        });
        sb2 = warp_sum64(sb2); sb3 = warp_sum64(sb3);
        cb2 = __reduce_add_sync(0xFFFFFFFFu, cb2); cb3 = __reduce_add_sync(0xFFFFFFFFu, cb3);
        below[0] = sb2 + (u64)(ranks[2] - cb2) * (u64)v2;
        below[1] = sb3 + (u64)(ranks[3] - cb3) * (u64)v3;
      }
    }
    PHASE_ADD(4);
    cs.med_lo = value[0]; cs.med_hi = value[1];
    cs.top2p5_sum = cs.sum - below[0];
    cs.top5_sum = cs.sum - below[1];
  }
  if (lane == 0) *dst = cs;
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// phase E: three chained exact EDTs on the 64-bit row masks
// ------------------------------------------------------------------------------------------------
// squared distance of column c to the nearest set bit of m (0xFFFFFFFF if m == 0)
__device__ __forceinline__ u32 nearest_bit_sq(u64 m, u32 c) {
  if (m == 0) return 0xFFFFFFFFu;
  const u64 le = m & (~0ull >> (63 - c));  // bits <= c
  const u64 ge = m >> c;                   // bits >= c, shifted
  u32 d = 64;
  if (le) d = c - (63u - (u32)__clzll((long long)le));
  if (ge) d = min(d, (u32)__ffsll((long long)ge) - 1u);
  return d * d;
}

template <class S, bool kLaneMask>
__device__ __forceinline__ void shape_edt_warp(const Obj& o, S& s, u32 rmin, u32 cmin, bool want_conical,
                                               ShapeStats* __restrict__ dst) {
  const u32 lane = lane_id();
  // g: row distances with one all-zero frame row above and below the window: [66][64] bytes,
  // i.e. vals[] plus the first 128 bytes of hist[]; the cone-top mask sits further into hist[]
  unsigned char* g = reinterpret_cast<unsigned char*>(s.vals);
  u64* topmask = reinterpret_cast<u64*>(s.hist + 256);         // [64]
  const int h = o.h, w = o.w;
  __syncwarp();
  // zero g (non-object pixels have row distance 0) and the cone-top mask
  {
    uint4* g4 = reinterpret_cast<uint4*>(g);
    for (int k = lane; k < ((kSide + 2) * kSide) / 16; k += 32) g4[k] = make_uint4(0, 0, 0, 0);
    topmask[lane] = 0; topmask[lane + 32] = 0;
  }
  __syncwarp();
  // ---- row distances of object pixels: nearest zero to the left/right (frame counts as zero) ----
  for_each_px(o, s, [&](u32 r, u32 c, u32) {
    const u64 m = s.rowmask[r];
    u64 z = ~m;
    if (w < 64) z |= (~0ull << w);                     // beyond the window: frame / other pixels = zero
    const u64 le = z & (~0ull >> (63 - c));            // zeros at columns <= c (never contains c itself)
    const u32 dl = le ? (c - (63u - (u32)__clzll((long long)le))) : (c + 1u);
    const u64 ge = z >> c;
    const u32 dr = ge ? ((u32)__ffsll((long long)ge) - 1u) : (64u - c);
    g[((r + 1u) << 6) | c] = (unsigned char)min(dl, dr);
  });
  __syncwarp();
  // ---- EDT 1: column pass with early exit; frame rows (-1 and h) have g = 0 ----
  // The frame rows make bounds checks unnecessary: the walk stops at the latest when it reaches
  // a frame row (candidate d^2 with g = 0), i.e. before it could leave the buffer.
  auto col_min = [&](u32 r, u32 c) -> u32 {
    const u32 k = ((r + 1u) << 6) | c;
    const u32 g0 = g[k];
    u32 best = g0 * g0;
    u32 d64 = 64, dd = 1, step = 3;  // d * 64, d * d, 2 d + 1
    while (dd < best) {
      const u32 m2 = min((u32)g[k - d64], (u32)g[k + d64]);
      best = min(best, m2 * m2 + dd);
      dd += step; step += 2; d64 += 64;
    }
    return best;
  };
  PHASE_T0();
  u32 lmax = 0;
  double s_nn = 0.0;
  u64 at_max = 0;  // bit j <-> the lane's j-th pixel (i = lane + 32 j) attains lmax (<= 64 pixels per lane)
  if (kLaneMask && o.listed) {
    u32 j = 0;
    for (u32 i = lane; i < o.n; i += 32, ++j) {
      const u32 k = s.offs[i];
      const u32 d2 = col_min(k >> 6, k & 63u);
      if (d2 > lmax) { lmax = d2; at_max = 1ull << j; }
      else if (d2 == lmax) at_max |= 1ull << j;
      if (want_conical) s_nn += sqrt((double)d2);
    }
  } else {
    for_each_px(o, s, [&](u32 r, u32 c, u32) {
      const u32 d2 = col_min(r, c);
      lmax = max(lmax, d2);
      if (want_conical) s_nn += sqrt((double)d2);
    });
  }
  const u32 max_nn2 = __reduce_max_sync(0xFFFFFFFFu, lmax);
  PHASE_ADD(6);
  if (want_conical) {
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s_nn += __shfl_xor_sync(0xFFFFFFFFu, s_nn, k);
  }
  // ---- cone top: pixels with nn2 == max ----
  if (kLaneMask && o.listed) {
    if (lmax == max_nn2) {
      while (at_max) {
        const u32 j = (u32)__ffsll((long long)at_max) - 1u;
        at_max &= at_max - 1;
        const u32 k = s.offs[lane + 32u * j];
        atomicOr(reinterpret_cast<unsigned long long*>(&topmask[k >> 6]), 1ull << (k & 63u));
      }
    }
  } else {  // only pixels with g^2 >= max can qualify: recompute those
    for_each_px(o, s, [&](u32 r, u32 c, u32) {
      const u32 g0 = g[((r + 1u) << 6) | c];
      if (g0 * g0 >= max_nn2 && col_min(r, c) == max_nn2) atomicOr(reinterpret_cast<unsigned long long*>(&topmask[r]), 1ull << c);
    });
  }
  __syncwarp();
  const u64 tm0 = topmask[lane], tm1 = topmask[lane + 32];
  const u32 n_top = __reduce_add_sync(0xFFFFFFFFu, (u32)(__popcll(tm0) + __popcll(tm1)));
  PHASE_ADD(7);
  // ---- EDT 2: distance of every object pixel to the nearest cone-top pixel ----
  u32 lmax2 = 0;
  if (n_top <= 32) {
    // each lane keeps one top pixel; extraction in row-major order
    u32 my_top = 0;
    {
      u64 a = tm0, b = tm1;
      u32 k = 0;
      for (int pass = 0; pass < 2; ++pass) {
        u64& cur = pass == 0 ? a : b;
        u32 any = __ballot_sync(0xFFFFFFFFu, cur != 0);
        while (any) {
          const int src = __ffs(any) - 1;
          const u64 mm = __shfl_sync(0xFFFFFFFFu, cur, src);
          const u32 c = (u32)__ffsll((long long)mm) - 1u;
          const u32 r = (u32)src + 32u * pass;
          if (lane == k) my_top = (r << 6) | c;
          ++k;
          if ((int)lane == src) cur &= cur - 1;
          any = __ballot_sync(0xFFFFFFFFu, cur != 0);
        }
      }
    }
    const bool lst = o.listed;
    // uniform trip count: iterate the compact list (or the window) with all lanes active in the shuffles
    const u32 iters = lst ? (o.n + 31) / 32 : 0;
    if (lst) {
      for (u32 it = 0; it < iters; ++it) {
        const u32 i = it * 32 + lane;
        const bool ok = i < o.n;
        const u32 k = ok ? (u32)s.offs[i] : 0u;
        const int r = (int)(k >> 6), c = (int)(k & 63u);
        u32 best = 0xFFFFFFFFu;
        for (u32 t = 0; t < n_top; ++t) {
          const u32 tp = __shfl_sync(0xFFFFFFFFu, my_top, t);
          const int dr = r - (int)(tp >> 6), dc = c - (int)(tp & 63u);
          best = min(best, (u32)(dr * dr + dc * dc));
        }
        if (ok) lmax2 = max(lmax2, best);
      }
    } else {
      for (int r = 0; r < h; ++r) {
        const u64 m = s.rowmask[r];
        for (u32 c0 = 0; c0 < (u32)w; c0 += 32) {
          const u32 c = c0 + lane;
          const bool ok = c < (u32)w && ((m >> c) & 1ull);
          u32 best = 0xFFFFFFFFu;
          for (u32 t = 0; t < n_top; ++t) {
            const u32 tp = __shfl_sync(0xFFFFFFFFu, my_top, t);
            const int dr = r - (int)(tp >> 6), dc = (int)c - (int)(tp & 63u);
            best = min(best, (u32)(dr * dr + dc * dc));
          }
          if (ok) lmax2 = max(lmax2, best);
        }
      }
    }
  } else {
    // plateau: rows of the cone-top mask, nearest set bit per row
    for_each_px(o, s, [&](u32 r, u32 c, u32) {
      u32 best = 0xFFFFFFFFu;
      for (int rr = 0; rr < h; ++rr) {
        const u64 tm = topmask[rr];
        if (tm == 0) continue;
        const u32 dc2 = nearest_bit_sq(tm, c);
        const int dr = (int)r - rr;
        best = min(best, dc2 + (u32)(dr * dr));
      }
      lmax2 = max(lmax2, best);
    });
  }
  const u32 max_dn2 = __reduce_max_sync(0xFFFFFFFFu, lmax2);
  // ---- EDT 3: size of the cone top = distance of each top pixel to the rest of the object ----
  double s_top = 0.0;
  if (n_top == o.n) {
    // `dn == 0` has no zero at all: SciPy measures to index (-1, 0) of the padded plane
    for_each_px(o, s, [&](u32 r, u32 c, u32) {
      const double dr = (double)rmin + (double)r + 2.0, dc = (double)cmin + (double)c + 1.0;
      s_top += sqrt(dr * dr + dc * dc);
    });
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s_top += __shfl_xor_sync(0xFFFFFFFFu, s_top, k);
  } else {
    // lanes over rows: q = object pixels that are not cone top
    const u64 q0 = (lane < (u32)h) ? (s.rowmask[lane] & ~tm0) : 0ull;
    const u64 q1 = (lane + 32 < (u32)h) ? (s.rowmask[lane + 32] & ~tm1) : 0ull;
    for (int r = 0; r < h; ++r) {
      u64 tm = topmask[r];  // warp-uniform
      while (tm) {
        const u32 c = (u32)__ffsll((long long)tm) - 1u;
        tm &= tm - 1;
        u32 best = 0xFFFFFFFFu;
        {
          const u32 d0 = nearest_bit_sq(q0, c);
          const int dr0 = r - (int)lane;
          if (d0 != 0xFFFFFFFFu) best = d0 + (u32)(dr0 * dr0);
          const u32 d1 = nearest_bit_sq(q1, c);
          const int dr1 = r - (int)lane - 32;
          if (d1 != 0xFFFFFFFFu) best = min(best, d1 + (u32)(dr1 * dr1));
        }
        best = __reduce_min_sync(0xFFFFFFFFu, best);
        s_top += sqrt((double)best);  // same value in every lane
      }
    }
  }
  if (lane == 0) {
    ShapeStats out;
    out.sum_nn = s_nn; out.sum_top = s_top; out.max_nn2 = max_nn2; out.max_dn2 = max_dn2;
    *dst = out;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
template <typename PX, int CAP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
object_warp_kernel(const uint16_t* __restrict__ labels, i64 lab_plane_stride, i64 lab_row_stride,
                   const int32_t* __restrict__ plane_tile, const int32_t* __restrict__ plane_base, int n_planes,
                   int n_objects, int n_total, const PX* __restrict__ pixels, const i64* __restrict__ tile_offset,
                   i64 chan_stride, i64 z_stride, i64 px_row_stride, int Z,
                   const abx_request* __restrict__ requests, int n_requests, int need_edt, int want_conical,
                   const abx_object_rec* __restrict__ recs, ChanStats* __restrict__ chan, ShapeStats* __restrict__ shape,
                   int* __restrict__ stats_list, int* __restrict__ edt_list, u32* __restrict__ list_counts) {
  extern __shared__ __align__(16) unsigned char dyn[];
  using S = WSmemT<CAP>;
  constexpr bool kPrimary = CAP == kCapSmall;  // the small class also zero-fills empty objects and builds the work lists
  constexpr u32 kLo = kPrimary ? 0u : (u32)kCapSmall;
  S& s = reinterpret_cast<S*>(dyn)[threadIdx.x >> 5];
  const u32 lane = lane_id();

  // dynamic work distribution: one atomic per object (objects differ 100x in cost); the warp always
  // holds the NEXT object too and prefetches its label / pixel windows into L2 while it works.
  auto fetch = [&]() -> int {
    int v = 0;
    if (lane == 0) v = (int)atomicAdd(&list_counts[kPrimary ? 2 : 3], 1u);
    return __shfl_sync(0xFFFFFFFFu, v, 0);
  };
  auto prefetch_object = [&](int nobj) {
    if (nobj >= n_objects) return;  // background objects go to the CTA kernels
    const abx_object_rec nr = recs[nobj];
    const int nh = (int)(nr.rmax - nr.rmin) + 1, nw = (int)(nr.cmax - nr.cmin) + 1;
    if (nr.n <= kLo || nr.n > (u32)CAP || nh > kSide || nw > kSide) return;
    const int np = find_plane(plane_base, n_planes, nobj);
    const i64 tail = (i64)nw - 1;
    for (int r = lane; r < nh; r += 32) {
      const uint16_t* lr = labels + (i64)np * lab_plane_stride + (i64)(nr.rmin + r) * lab_row_stride + nr.cmin;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(lr));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(lr + tail));
    }
    if (n_requests > 0) {
      const PX* base = pixels + tile_offset[plane_tile[np]] + (i64)nr.rmin * px_row_stride + nr.cmin;
      const int zmax = Z < 16 ? Z : 16;
      for (int q = 0; q < n_requests; ++q) {
        const PX* cb = base + (i64)requests[q].channel * chan_stride;
        for (int z = 0; z < zmax; ++z)
          for (int r = lane; r < nh; r += 32) {
            const PX* pr = cb + (i64)z * z_stride + (i64)r * px_row_stride;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pr));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + tail));
          }
      }
    }
  };
  int obj = fetch();
  int nxt = obj < n_total ? fetch() : n_total;
  for (; obj < n_total; obj = nxt, nxt = fetch()) {
    if (nxt < n_total) prefetch_object(nxt);
    const abx_object_rec rec = recs[obj];
    const bool is_bg = obj >= n_objects;
    if (!kPrimary && (rec.n <= kLo || rec.n > (u32)CAP)) continue;
    if (rec.n == 0) {
      for (int q = lane; q < n_requests; q += 32) {
        ChanStats z;
        z.sum = z.sumsq = z.wrapsq = z.m10 = z.m01 = z.m20 = z.m02 = z.top2p5_sum = z.top5_sum = 0;
        z.vmin = z.vmax = z.med_lo = z.med_hi = 0;
        chan[(i64)obj * n_requests + q] = z;
      }
      if (!is_bg && need_edt && lane == 0) {
        ShapeStats z; z.sum_nn = 0; z.sum_top = 0; z.max_nn2 = 0; z.max_dn2 = 0;
        shape[obj] = z;
      }
      continue;
    }
    const int h = (int)(rec.rmax - rec.rmin) + 1, w = (int)(rec.cmax - rec.cmin) + 1;
    if (is_bg || h > kSide || w > kSide) {  // hand over to the CTA-per-object kernels
      if (kPrimary && lane == 0) {
        if (n_requests > 0) stats_list[atomicAdd(&list_counts[0], 1u)] = obj;
        if (!is_bg && need_edt) edt_list[atomicAdd(&list_counts[1], 1u)] = obj;
      }
      continue;
    }
    const int p = find_plane(plane_base, n_planes, obj);
    Obj o;
    o.label = (u32)(obj - plane_base[p] + 1);
    o.n = rec.n; o.h = h; o.w = w;
    o.lab_rs = lab_row_stride;
    o.lab = labels + (i64)p * lab_plane_stride + (i64)rec.rmin * lab_row_stride + rec.cmin;
    if (kPrimary && rec.n > (u32)CAP) continue;  // the large size class takes it
    o.listed = true;

    // ---- phase M: row bitmasks, row bases, compact offset list ----
    PHASE_T0();
    __syncwarp();
    {
      u32 base = 0;
      constexpr int kRows = 4;
      for (int r0 = 0; r0 < h; r0 += kRows) {
        u32 l0[kRows], l1[kRows];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {  // all loads of the row group first
          const uint16_t* lrow = o.lab + (i64)(r0 + u) * o.lab_rs;
          const bool in = r0 + u < h;
          l0[u] = (in && lane < (u32)w) ? (u32)__ldg(lrow + lane) : 0xFFFFFFFFu;
          l1[u] = (in && lane + 32 < (u32)w) ? (u32)__ldg(lrow + lane + 32) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          const int r = r0 + u;
          if (r >= h) break;
          const bool hit0 = l0[u] == o.label, hit1 = l1[u] == o.label;
          const u32 b0 = __ballot_sync(0xFFFFFFFFu, hit0);
          const u32 b1 = __ballot_sync(0xFFFFFFFFu, hit1);
          if (lane == 0) { s.rowmask[r] = (u64)b0 | ((u64)b1 << 32); s.rowbase[r] = (unsigned short)base; }
          if (o.listed) {
            const u32 lt = (1u << lane) - 1u;
            if (hit0) s.offs[base + __popc(b0 & lt)] = (unsigned short)((r << 6) | lane);
            if (hit1) s.offs[base + __popc(b0) + __popc(b1 & lt)] = (unsigned short)((r << 6) | (lane + 32));
          }
          base += __popc(b0) + __popc(b1);
        }
      }
      if (lane >= (u32)h) s.rowmask[lane] = 0;  // rows beyond the window read as empty
      if (lane + 32 >= (u32)h) s.rowmask[lane + 32] = 0;
    }
    __syncwarp();
    PHASE_ADD(0);

    // ---- phase S ----
    if (n_requests > 0) {
      const int tile = plane_tile[p];
      const PX* px0 = pixels + tile_offset[tile] + (i64)rec.rmin * px_row_stride + rec.cmin;
      for (int q = 0; q < n_requests; ++q) {
        const abx_request rq = requests[q];
        request_stats<PX>(o, s, px0 + (i64)rq.channel * chan_stride, px_row_stride, z_stride, Z, rq, rq.features,
                          chan + (i64)obj * n_requests + q);
      }
    }
    // ---- phase E ----
#ifdef ABX_PHASE_TIMING
    _pt = clock64();
#endif
    if (need_edt) shape_edt_warp<S, (CAP <= 2048)>(o, s, rec.rmin, rec.cmin, want_conical != 0, shape + obj);
    PHASE_ADD(5);
  }
}

}  // namespace

#ifdef ABX_PHASE_TIMING
extern "C" int abx_debug_phase_cycles(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, g_phase_cycles, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
  }
  return 0;
}
#endif

template <typename PX, int CAP, int WARPS>
static int launch_class(const abx_extract_args* a, const Workspace& ws, cudaStream_t st, int n_total) {
  const size_t smem = sizeof(WSmemT<CAP>) * WARPS;
  static thread_local bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(object_warp_kernel<PX, CAP, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return abx_check_cuda(e, "object_warp smem attribute");
    done[dev] = true;
  }
  int grid = (n_total + WARPS - 1) / WARPS;
  if (grid > 148 * 2) grid = 148 * 2;  // persistent: 2 CTAs per SM, warps pull objects from a counter
  object_warp_kernel<PX, CAP, WARPS><<<grid, WARPS * 32, smem, st>>>(
      static_cast<const uint16_t*>(a->labels), a->label_plane_stride, a->label_row_stride, a->plane_tile, a->plane_base,
      a->n_planes, a->n_objects, n_total, static_cast<const PX*>(a->pixels), reinterpret_cast<const i64*>(a->tile_offset),
      a->chan_stride, a->z_stride, a->row_stride, a->Z, a->requests, a->n_requests, a->need_edt, (a->need_edt & 2) != 0,
      ws.recs, ws.chan, ws.shape, ws.stats_list, ws.edt_list, ws.list_counts);
  return abx_check_cuda(cudaGetLastError(), "object_warp");
}

template <typename PX>
static int launch_both(const abx_extract_args* a, const Workspace& ws, cudaStream_t st, int n_total) {
  int rc = launch_class<PX, kCapSmall, kWarpsSmall>(a, ws, st, n_total);
  if (rc) return rc;
  return launch_class<PX, kCapLarge, kWarpsLarge>(a, ws, st, n_total);
}

int launch_object_warp(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  if (n_total == 0 || (a->n_requests == 0 && !a->need_edt)) return ABX_OK;
  if (a->n_requests == 0 || a->pixel_dtype == ABX_U16) return launch_both<uint16_t>(a, ws, st, n_total);
  if (a->pixel_dtype == ABX_U8) return launch_both<uint8_t>(a, ws, st, n_total);
  return abx_set_error(ABX_ERR_UNSUPPORTED, "object_warp: pixel dtype %d has no kernel", a->pixel_dtype);
}
