// Per-object intensity statistics with TMA-STAGED WINDOWS: one warp per object, every window of the object —
// its label window and then, request by request, its pixel window — is brought into the warp's shared-memory
// slot by cp.async.bulk.tensor boxes (one elected lane, mbarrier completion) and all the arithmetic reads shared
// memory.  This is the fast path of object_stats_warp (object_warp.cu, which stays as the path for layouts
// TMA cannot address: unaligned bases / strides; Z stacks reach this kernel as reduced planes, zreduce.cu).  Same
// outputs, same reference semantics
// (src/extraction/extract.py:346-359 loop; cell.py:43-157,232-265; tile crop of tiler.py:309-366 fused
// through the tile offset).
//
//   slot (16 KB per warp, 14 warps per CTA, one CTA per SM):
//     t u32[16] | mbarrier | flex 16 256 B = offs u16[n_pad] | window [h8][32 or 64] PX | hist u32[1024 or 512]
//   window pitch 32 elements for objects at most 32 columns wide, 64 otherwise, so that a list entry
//   k = (r << shift) | c IS the element index of its pixel inside the window: no address arithmetic.
//   phase M  label window by TMA (boxes of 8 rows) -> compact list of k, padded to a multiple of 128 entries
//            with copies of entry 0 (taken out of the sums afterwards; cannot move min / max)
//   phase S  per request: pixel window by TMA, pass 1 moments + extrema, pass 2 range-adaptive 1024-bin
//            histogram (ATOMS.POPC.INC) and the four ranks by warp scans; 7-bit refinement when range > 1023
//   While request q is processed, the window of request q + 1 (or the label window and first request of the
//   warp's next object) is prefetched into L2, so that the TMA reads hit L2.
// Windows above 64 x 64 and the per-plane background go to the CTA-per-object kernel (object_stats.cu) through the
// hand-over list; window-sized objects that do not fit a slot or a 64-column box go to object_stats_warp through a
// second list.  Big objects are taken first, and short queues are split into several work items per object.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <cstring>

#include "common.cuh"

namespace {

#include "warp_common.cuh"
#include "tma.cuh"

constexpr int kTmaWarps = 14;
constexpr u32 kTmaSlot = 16384;
constexpr u32 kTOff = 0, kBarOff = 64, kFlexOff = 128, kFlex = kTmaSlot - kFlexOff;  // byte offsets inside a slot
constexpr int kBoxRows = 8;
constexpr u32 kBigFirst = 1536;  // objects above this many pixels are processed first
static_assert(kFlexOff % 128 == 0, "TMA destinations are 128-byte aligned");

// phase M from the staged label window: compact list of k = (r << sh) | c, row-major, no atomics
__device__ __forceinline__ void build_list_smem(u32 lwin_off, u32 label, int h, int sh, u32 offs_off) {
  const unsigned short* lwin = reinterpret_cast<const unsigned short*>(dyn + lwin_off);
  unsigned short* offs = reinterpret_cast<unsigned short*>(dyn + offs_off);
  const u32 lane = lane_id();
  const u32 lt = (1u << lane) - 1u;
  const bool two = sh == 6;
  u32 base = 0;
#pragma unroll 4
  for (int r = 0; r < h; ++r) {
    const u32 row = (u32)r << sh;
    // columns beyond the bbox hold other labels (or the hardware's zero fill): they never match
    const bool hit0 = (u32)lwin[row + lane] == label;
    const bool hit1 = two && (u32)lwin[row + (two ? 32u : 0u) + lane] == label;
    const u32 b0 = __ballot_sync(kFull, hit0);
    const u32 b1 = __ballot_sync(kFull, hit1);
    if (hit0) offs[base + __popc(b0 & lt)] = (unsigned short)(row | lane);
    base += __popc(b0);
    if (hit1) offs[base + __popc(b1 & lt)] = (unsigned short)(row | (lane + 32u));
    base += __popc(b1);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// phase S: one request on the staged window
// ------------------------------------------------------------------------------------------------
template <typename PX>
__device__ __forceinline__ void request_stats_win(u32 n, u32 n_pad, u32 slot_off, u32 win_off, u32 hist_off, u32 bins, int sh,
                                                  u32 feats, ChanStats* __restrict__ dst) {
  constexpr int kShift = (sizeof(PX) == 1) ? 12 : 8;  // (x << kShift)^2 >> 32 == x^2 >> bits(PX)
  const unsigned short* offs = reinterpret_cast<const unsigned short*>(dyn + slot_off + kFlexOff);
  const PX* win = reinterpret_cast<const PX*>(dyn + win_off);
  u32* hist = reinterpret_cast<u32*>(dyn + hist_off);
  u32* t = reinterpret_cast<u32*>(dyn + slot_off + kTOff);
  const u32 lane = lane_id();
  const u32 cmask = (1u << sh) - 1u;
  const bool want_moi = (feats & ABX_F_MOI) != 0;
  ChanStats cs;
  // ---- pass 1: moments and extrema over the padded list ----
  {
    u32 f_sum = 0, f_wh = 0, f_m10 = 0, f_m01 = 0;
    u64 f_sq = 0, f_q = 0;
    u32 a_min = kFull, a_max = 0;
#pragma unroll 2
    for (u32 i0 = lane; i0 < n_pad; i0 += 128) {
      u32 k[4], v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) k[u] = offs[i0 + 32u * u];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (u32)win[k[u]];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        f_sum += v[u];
        f_sq += (u64)v[u] * (u64)v[u];
        const u32 a = v[u] << kShift;
        f_wh += __umulhi(a, a);
        a_min = min(a_min, v[u]);
        a_max = max(a_max, v[u]);
        if (want_moi) {
          const u32 c = k[u] & cmask, r = k[u] >> sh;
          f_m10 += v[u] * c;
          f_m01 += v[u] * r;
          f_q += (u64)v[u] * (u64)(c * c + r * r);
        }
      }
    }
    // take the padding (n_pad - n copies of entry 0) out of the sums
    const u32 k0 = offs[0];
    const u32 v0 = (u32)win[k0];
    const u32 c0 = k0 & cmask, r0 = k0 >> sh;
    const u32 a0 = v0 << kShift;
    const u64 p = (u64)(n_pad - n);
    cs.sum = (u64)__reduce_add_sync(kFull, f_sum) - p * v0;  // n * 65535 < 2^32
    cs.sumsq = warp_sum64(f_sq) - p * ((u64)v0 * v0);
    const u64 wh = (u64)__reduce_add_sync(kFull, f_wh) - p * __umulhi(a0, a0);
    cs.wrapsq = cs.sumsq - (wh << (32 - 2 * kShift));  // sum of (x^2 mod 2^bits): NumPy's v**2 in the image dtype
    if (want_moi) {
      cs.m10 = warp_sum64((u64)f_m10) - p * (u64)(v0 * c0);
      cs.m01 = warp_sum64((u64)f_m01) - p * (u64)(v0 * r0);
      cs.m20 = warp_sum64(f_q) - p * ((u64)v0 * (u64)(c0 * c0 + r0 * r0));  // m20 + m02 as one sum (finalize.cu)
      cs.m02 = 0;
    } else {
      cs.m10 = cs.m01 = cs.m20 = cs.m02 = 0;
    }
    cs.vmin = __reduce_min_sync(kFull, a_min);
    cs.vmax = __reduce_max_sync(kFull, a_max);
  }
  cs.med_lo = cs.med_hi = 0;
  cs.top2p5_sum = cs.top5_sum = 0;

  if (feats & (ABX_F_MEDIAN | ABX_F_TOP2P5 | ABX_F_TOP5)) {
    const u32 vmin = cs.vmin;
    // ---- pass 2: range-adaptive histogram ----
    const u32 range = cs.vmax - vmin;
    int s0 = 0;
    while ((range >> s0) >= bins) ++s0;
    const u32 nb = (range >> s0) + 1;
    __syncwarp();
    hist_zero(hist, 32u * bins_per_lane(nb));
    __syncwarp();
    const u32 hbase = smem_addr(hist);
    if (s0 == 0) {
      const u32 base = hbase - 4u * vmin;  // bin address = 4 * value + base
#pragma unroll 2
      for (u32 i0 = lane; i0 < n_pad; i0 += 128) {
        u32 k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) k[u] = offs[i0 + 32u * u];
#pragma unroll
        for (int u = 0; u < 4; ++u) hist_inc(4u * (u32)win[k[u]] + base);
      }
    } else {
#pragma unroll 2
      for (u32 i0 = lane; i0 < n_pad; i0 += 128) {
        u32 k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) k[u] = offs[i0 + 32u * u];
#pragma unroll
        for (int u = 0; u < 4; ++u) hist_inc(hbase + 4u * (((u32)win[k[u]] - vmin) >> s0));
      }
    }
    __syncwarp();
    if (lane == 0) hist[((u32)win[offs[0]] - vmin) >> s0] -= n_pad - n;  // the padding
    __syncwarp();
    const u32 k2p5 = (u32)ceil((double)n * 0.025);  // int(np.ceil(n * 0.025)), cell.py:110-111
    const u32 k5 = min(n, 5u);
    const u32 ranks[4] = {(n - 1) / 2, n / 2, n - k2p5, n - k5};
    find_ranks32(hist, nb, ranks, t);
    u32 v2, v3;
    u64 below2, below3;
    if (s0 == 0) {
      cs.med_lo = vmin + t[0]; cs.med_hi = vmin + t[1];
      v2 = vmin + t[2]; v3 = vmin + t[3];
      // sum of the smallest values up to the rank = vmin * cnt + sum(count * bin) below + rank * value
      below2 = (u64)vmin * t[10] + t[14] + (u64)t[6] * v2;
      below3 = (u64)vmin * t[11] + t[15] + (u64)t[7] * v3;
    } else {
      // ---- refinement: 7 more bits per sweep inside the four target bins, searched in parallel ----
      int cur = s0;
      u32 key[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) key[j] = t[j];
#pragma unroll 1
      while (cur > 0) {
        const int nxt = cur > 7 ? cur - 7 : 0;
        const u32 nsub = 1u << (cur - nxt);
        __syncwarp();
        hist_zero(hist, 512u);
        __syncwarp();
#pragma unroll 1
        for (u32 i = lane; i < n; i += 32) {
          const u32 d = (u32)win[offs[i]] - vmin;
          const u32 hi = d >> cur;
          const u32 sb = (d >> nxt) & (nsub - 1u);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (hi == key[j]) hist_add(hist, 128u * j + sb);
        }
        __syncwarp();
        find_ranks32_x4(hist, t);
#pragma unroll
        for (int j = 0; j < 4; ++j) key[j] = (key[j] << (cur - nxt)) | t[j];
        cur = nxt;
      }
      cs.med_lo = vmin + key[0]; cs.med_hi = vmin + key[1];
      v2 = vmin + key[2]; v3 = vmin + key[3];
      below2 = below3 = 0;
      if (feats & (ABX_F_TOP2P5 | ABX_F_TOP5)) {
        u64 sb2 = 0, sb3 = 0;
        u32 cb2 = 0, cb3 = 0;
#pragma unroll 1
        for (u32 i = lane; i < n; i += 32) {
          const u32 x = (u32)win[offs[i]];
          if (x < v2) { sb2 += x; ++cb2; }
          if (x < v3) { sb3 += x; ++cb3; }
        }
        sb2 = warp_sum64(sb2); sb3 = warp_sum64(sb3);
        cb2 = __reduce_add_sync(kFull, cb2); cb3 = __reduce_add_sync(kFull, cb3);
        below2 = sb2 + (u64)(ranks[2] - cb2) * (u64)v2;
        below3 = sb3 + (u64)(ranks[3] - cb3) * (u64)v3;
      }
    }
    cs.top2p5_sum = cs.sum - below2;
    cs.top5_sum = cs.sum - below3;
  }
  if (lane == 0) *dst = cs;
  __syncwarp();
}

struct TmaMaps {
  CUtensorMap lab[2];  // label planes (W, H, P), box 32 | 64 columns x 8 rows
  CUtensorMap px[2];   // the pixel buffer as rows of row_stride elements, box 32 | 64 columns x 8 rows
};

template <typename PX, bool kSplit>
__global__ void __launch_bounds__(kTmaWarps * 32, 1)
object_stats_tma(const __grid_constant__ TmaMaps maps, const Common cm, const PX* __restrict__ pixels,
                 const i64* __restrict__ tile_offset, i64 chan_stride, i64 px_row_stride, int chan_rows,
                 const abx_request* __restrict__ requests, int n_requests, ChanStats* __restrict__ chan,
                 int* __restrict__ stats_list, u32* __restrict__ stats_count, u32* __restrict__ warp_count, int list_cap,
                 int split_log2_arg) {
  const int split_log2 = kSplit ? split_log2_arg : 0;  // the unsplit instantiation folds all of this away
  const u32 lane = lane_id();
  const u32 slot_off = (threadIdx.x >> 5) * kTmaSlot;
  const u32 bar = smem_addr(dyn + slot_off + kBarOff);
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  u32 parity = 0;
  // Longest first: the queue runs over the objects twice — items [0, n_total) take only the big objects, items
  // [n_total, 2 n_total) the others — so that the launch does not end on one warp working through a 2 500-pixel cell.
  // Short queues (one small field: fewer objects than resident warps) are split: `split` work items per object, each
  // with its share of the requests (the label window and the list are then rebuilt per item, which a launch that would
  // otherwise last one object's latency can afford).
  const int split = 1 << split_log2;  // a power of two: no divisions below
  const int n_half = cm.n_total << split_log2;
  const int n_items = 2 * n_half;
  Queue qu{cm.counters, n_items, 0};
  int item = qu.fetch();
  int nxt_item = item < n_items ? qu.fetch() : n_items;
  while (item < n_items) {
    const bool big_pass = item < n_half;
    const int unit = big_pass ? item : item - n_half;
    const int obj = unit >> split_log2, part = unit & (split - 1);
    const int q_lo = (part * n_requests) >> split_log2, q_hi = ((part + 1) * n_requests) >> split_log2;  // this item's requests
    const int nxt_unit = nxt_item < n_half ? nxt_item : nxt_item - n_half;
    const int nxt = nxt_item < n_items ? nxt_unit >> split_log2 : cm.n_total;
    const abx_object_rec rec = cm.recs[obj];
    if ((rec.n > kBigFirst) != big_pass) {  // not this pass
      item = nxt_item;
      nxt_item = item < n_items ? qu.fetch() : n_items;
      continue;
    }
    const bool is_bg = obj >= cm.n_objects;
    const int h = (int)(rec.rmax - rec.rmin) + 1, w = (int)(rec.cmax - rec.cmin) + 1;
    constexpr u32 kAlign = 16u / (u32)sizeof(PX);  // a TMA box starts at a 16-byte multiple of its innermost coordinate
    const bool cand = rec.n > 0 && !is_bg && h <= kSide && w <= kSide;
    // the window's origin in the pixel buffer seen as rows of row_stride elements
    int p = 0, col0 = 0;
    i64 org = 0, row0 = 0;
    if (cand) {
      p = find_plane(cm.plane_base, cm.n_planes, obj);
      org = tile_offset[cm.plane_tile[p]] + (i64)rec.rmin * px_row_stride + rec.cmin;
      if ((u64)org < 0x100000000ull && (u64)px_row_stride < 0x100000000ull) row0 = (i64)((u32)org / (u32)px_row_stride);
      else row0 = org / px_row_stride;
      col0 = (int)(org - row0 * px_row_stride);
    }
    // both boxes start left of the bbox, at aligned columns: the list indexes the label box, the pixel box is read
    // through a base shifted by the difference of the two margins
    const u32 s_lab = rec.cmin & (kAlign - 1u), s_px = (u32)col0 & (kAlign - 1u);
    const u32 need = (u32)w + max(s_lab, s_px);
    const int sh = need <= 32u ? 5 : 6;
    const u32 h8 = ((u32)h + kBoxRows - 1u) & ~(u32)(kBoxRows - 1);
    const u32 n_pad = (rec.n + (u32)kPad - 1u) & ~((u32)kPad - 1u);
    if (rec.n == 0) {  // absent label (or empty background): zero records -> NaN in finalize
      for (int q = q_lo + (int)lane; q < q_hi; q += 32) {
        ChanStats z;
        z.sum = z.sumsq = z.wrapsq = z.m10 = z.m01 = z.m20 = z.m02 = z.top2p5_sum = z.top5_sum = 0;
        z.vmin = z.vmax = z.med_lo = z.med_hi = 0;
        chan[(i64)obj * n_requests + q] = z;
      }
    } else if (!cand) {
      if (lane == 0 && part == 0) stats_list[atomicAdd(stats_count, 1u)] = obj;  // hand over to the CTA-per-object kernel
    } else if (need > 64u || 2u * n_pad + ((h8 << sh) << 1) + 2048u > kFlex) {
      // window-sized but too wide for a 64-column box at this alignment, or list + window + a 512-bin histogram above
      // the flex area (a few cells per thousand): second list, filled from the back of the same buffer, for
      // object_stats_warp
      if (lane == 0 && part == 0) stats_list[list_cap - 1 - (int)atomicAdd(warp_count, 1u)] = obj;
    } else {
      const u32 win_off = slot_off + kFlexOff + 2u * n_pad;  // n_pad is a multiple of 128: 128-byte aligned
      // the histogram takes what is left behind the window: 1024 bins, or 512 for the largest objects
      const u32 hist_off = win_off + ((h8 << sh) << 1);
      const u32 bins = (slot_off + kTmaSlot - hist_off >= (u32)kBins * 4u) ? (u32)kBins : 512u;
      const u32 win_addr = smem_addr(dyn + win_off);
      const int cls = sh - 5;
      // ---- label window -> list ----
      __syncwarp();
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses to the flex area
        const u32 box_bytes = (u32)kBoxRows << (sh + 1);
        mbar_expect_tx(bar, h8 << (sh + 1));
        for (u32 j = 0; j < h8; j += kBoxRows)
          tma_box_3d(win_addr + (j >> 3) * box_bytes, cls ? &maps.lab[1] : &maps.lab[0], (int)(rec.cmin - s_lab), (int)(rec.rmin + j), p, bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1u;
      build_list_smem(win_off, (u32)(obj - cm.plane_base[p] + 1), h, sh, slot_off + kFlexOff);
      {
        unsigned short* offs = reinterpret_cast<unsigned short*>(dyn + slot_off + kFlexOff);
        const unsigned short first = offs[0];
        for (u32 i = rec.n + lane; i < n_pad; i += 32) offs[i] = first;
        __syncwarp();
      }
      const PX* px0 = pixels + org;
#pragma unroll 1
      for (int q = q_lo; q < q_hi; ++q) {
        const abx_request rq = requests[q];
        // L2 prefetch one request ahead: the next request of this object, or the label window and the first
        // request of the warp's next object (a longer distance does not survive in L2 at these rates)
        if (q + 1 < q_hi) {
          prefetch_request<PX>(px0 + (i64)requests[q + 1].channel * chan_stride, px_row_stride, 0, 1, h, w);
        } else if (nxt < cm.n_objects) {
          const abx_object_rec nr = cm.recs[nxt];
          const int nh = (int)(nr.rmax - nr.rmin) + 1, nw = (int)(nr.cmax - nr.cmin) + 1;
          if (nr.n > 0 && nh <= kSide && nw <= kSide && (nr.n > kBigFirst) == (nxt_item < n_half)) {
            const int np = find_plane(cm.plane_base, cm.n_planes, nxt);
            prefetch_rows(cm.labels + (i64)np * cm.lab_plane_stride + (i64)nr.rmin * cm.lab_row_stride + nr.cmin,
                          cm.lab_row_stride * 2, nh, (u32)nw * 2u);
            prefetch_request<PX>(pixels + tile_offset[cm.plane_tile[np]] + (i64)nr.rmin * px_row_stride + nr.cmin +
                                     (i64)requests[((nxt_unit & (split - 1)) * n_requests) >> split_log2].channel * chan_stride,
                                 px_row_stride, 0, 1, nh, nw);
          }
        }
        if (rq.reduction == ABX_RED_DIV) continue;  // floating-point request: object_float.cu
        __syncwarp();  // every lane is done with the previous window
        if (lane == 0) {
          const u32 box_bytes = ((u32)kBoxRows << sh) * (u32)sizeof(PX);
          mbar_expect_tx(bar, (h8 << sh) * (u32)sizeof(PX));
          const int row = (int)row0 + rq.channel * chan_rows;
          for (u32 j = 0; j < h8; j += kBoxRows)
            tma_box_2d(win_addr + (j >> 3) * box_bytes, cls ? &maps.px[1] : &maps.px[0], col0 - (int)s_px, row + (int)j, bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        request_stats_win<PX>(rec.n, n_pad, slot_off, win_off + (s_px - s_lab) * (u32)sizeof(PX), hist_off, bins, sh, rq.features, chan + (i64)obj * n_requests + q);
      }
    }
    item = nxt_item;
    nxt_item = item < n_items ? qu.fetch() : n_items;
  }
}

// The four tensor maps, or false when the layout does not qualify for TMA (the caller then takes object_stats_warp).
bool make_maps(const abx_extract_args* a, TmaMaps* m) {
  EncodeFn encode = tensor_map_encoder();
  if (!encode || a->Z != 1 || a->pixel_elems <= 0) return false;
  const size_t es = a->pixel_dtype == ABX_U8 ? 1 : 2;
  const i64 lab_ps = a->n_planes > 1 ? a->label_plane_stride : (i64)a->H * a->label_row_stride;
  if ((reinterpret_cast<uintptr_t>(a->labels) & 15u) || a->label_row_stride % 8 || lab_ps % 8 || a->W < 64 || a->H < kBoxRows ||
      a->label_row_stride < a->W || lab_ps < (i64)a->H * a->label_row_stride)
    return false;
  if ((reinterpret_cast<uintptr_t>(a->pixels) & 15u) || (a->row_stride * (i64)es) % 16 || a->row_stride < 64 ||
      a->chan_stride % a->row_stride || a->chan_stride / a->row_stride > 0x7FFFFFFF / (a->C > 0 ? a->C : 1))
    return false;
  const i64 rows = a->pixel_elems / a->row_stride;  // whole rows inside the caller's buffer
  if (rows < kBoxRows || rows > 0x7FFFFFFF) return false;
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;  // boxes are exact windows: no over-fetch
  for (int cls = 0; cls < 2; ++cls) {
    const cuuint32_t bw = cls ? 64u : 32u;
    const cuuint64_t ldim[3] = {(cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->n_planes};
    const cuuint64_t lstr[2] = {(cuuint64_t)a->label_row_stride * 2u, (cuuint64_t)lab_ps * 2u};
    const cuuint32_t lbox[3] = {bw, (cuuint32_t)kBoxRows, 1u};
    if (encode(&m->lab[cls], CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(a->labels), ldim, lstr, lbox, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
    const cuuint64_t pdim[2] = {(cuuint64_t)a->row_stride, (cuuint64_t)rows};
    const cuuint64_t pstr[1] = {(cuuint64_t)a->row_stride * es};
    const cuuint32_t pbox[2] = {bw, (cuuint32_t)kBoxRows};
    if (encode(&m->px[cls], es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
               const_cast<void*>(a->pixels), pdim, pstr, pbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

template <typename PX>
int launch_tma(const abx_extract_args* a, const Workspace& ws, const TmaMaps& maps, const Common& cm, cudaStream_t st,
               int warps) {
  constexpr size_t smem_max = (size_t)kTmaWarps * kTmaSlot;
  const size_t smem = (size_t)warps * kTmaSlot;
  static thread_local bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(object_stats_tma<PX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(object_stats_tma<PX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return abx_check_cuda(e, "object_stats_tma smem attribute");
    done[dev] = true;
  }
  // fewer objects than two rounds of resident warps: several work items per object
  int split_log2 = 0;
  while (split_log2 < 3 && (cm.n_total << split_log2) < 2 * 148 * warps && (2 << split_log2) <= a->n_requests) ++split_log2;
  const int split = 1 << split_log2;
  int grid = (cm.n_total * split + warps - 1) / warps;
  if (grid > 148) grid = 148;  // persistent: one CTA per SM, warps pull objects from a counter
#define ABX_LAUNCH_TMA(SPLIT)                                                                                          \
  object_stats_tma<PX, SPLIT><<<grid, warps * 32, smem, st>>>(                                                          \
      maps, cm, static_cast<const PX*>(a->pixels), reinterpret_cast<const i64*>(a->tile_offset), a->chan_stride,        \
      a->row_stride, (int)(a->chan_stride / a->row_stride), a->requests, a->n_requests, ws.chan, ws.stats_list,         \
      ws.list_counts, ws.list_counts + 3, a->n_objects + a->n_planes, split_log2)
  if (split_log2) ABX_LAUNCH_TMA(true);
  else ABX_LAUNCH_TMA(false);
#undef ABX_LAUNCH_TMA
  return abx_check_cuda(cudaGetLastError(), "object_stats_tma");
}

}  // namespace

// Whether abx_extract will take the TMA kernel for this call (layout and dtype qualify).
bool abx_stats_tma_ok(const abx_extract_args* a) {
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  if (n_total == 0 || a->n_requests == 0) return false;
  if (a->pixel_dtype != ABX_U16 && a->pixel_dtype != ABX_U8) return false;
  TmaMaps maps;
  memset(&maps, 0, sizeof(maps));
  return make_maps(a, &maps);
}

// Returns ABX_OK and *launched = true when the TMA kernel took the statistics of the window-sized objects.
// shared_sm: the shape kernel runs at the same time on another stream — half the warps per CTA, so that one CTA of each
// kernel fits an SM.
int launch_object_stats_tma(const abx_extract_args* a, const Workspace& ws, cudaStream_t st, bool shared_sm, bool* launched) {
  *launched = false;
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  if (n_total == 0 || a->n_requests == 0) return ABX_OK;
  if (a->pixel_dtype != ABX_U16 && a->pixel_dtype != ABX_U8) return ABX_OK;
  TmaMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (!make_maps(a, &maps)) return ABX_OK;
  Common cm;
  cm.labels = static_cast<const uint16_t*>(a->labels);
  cm.lab_plane_stride = a->label_plane_stride;
  cm.lab_row_stride = a->label_row_stride;
  cm.plane_tile = a->plane_tile;
  cm.plane_base = a->plane_base;
  cm.n_planes = a->n_planes;
  cm.n_objects = a->n_objects;
  cm.n_total = n_total;
  cm.recs = ws.recs;
  cm.counters = ws.list_counts + 2;
  *launched = true;
  const int warps = shared_sm ? kTmaWarps / 2 : kTmaWarps;
  if (a->pixel_dtype == ABX_U16) return launch_tma<uint16_t>(a, ws, maps, cm, st, warps);
  return launch_tma<uint8_t>(a, ws, maps, cm, st, warps);
}
