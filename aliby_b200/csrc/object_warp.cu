// Per-object statistics of objects whose bounding box fits 64 x 64, by PLAIN GATHERS: one warp per object.
//
// object_stats_warp covers, for one object, everything the reference computes with |instructions| full-plane passes
// (src/extraction/extract.py:346-359): the intensity statistics of every (channel, Z-reduction) request
// (cell.py:43-157,232-265; distributors.py:19-21 fused into the load; tile crop of tiler.py:309-366 fused through the
// tile offset).  It is the twin of object_sweep (object_sweep.cu, TMA-staged windows) and runs
//   (a) for whole launches whose layout TMA cannot address (unaligned bases / row strides, planes narrower than 64,
//       Z stacks whose rows are not 16-byte multiples),
//   (b) as one work item per (object, request) for the few objects object_sweep leaves over,
//   (c) on the uint32 sum planes of Z-add requests (zreduce.cu).
//
//   phase M  label window -> compact list of the object's pixel offsets ((r << 6) | c, row-major) from warp ballots:
//            no atomics, deterministic order; padded to a multiple of 128 with copies of entry 0
//   phase S  per request: gather through the list (coalesced along rows, a software pipeline of four loads per lane
//            and stage), moments in registers, values staged in shared memory when they fit, range-adaptive 1024-bin
//            histogram (ATOMS.POPC.INC), the four ranks located by warp scans; 7-bit refinement sweeps only when the
//            value range exceeds 1023
//
// SMALL CODE FOOTPRINT matters more than anything else here (one path per phase, loops not unrolled beyond what
// memory-level parallelism needs, only the cold variants — wide sums, re-gather — out of line): the first version of
// this file compiled to 165 KB of SASS and stalled on instruction fetch (profiles/r01e_summary.md), and a five-deep
// gather pipeline did so again.  Hot phases are inlined at their single call site, because
// a __noinline__ call cost a frame in local memory (profiles/r01_final_summary.md).  No __syncthreads: the warps of a
// CTA are independent.  Larger windows and the per-plane background go to work lists that the CTA-per-object kernels
// (object_stats.cu, shape_edt.cu) and background.cu consume.
#include <type_traits>

#include "common.cuh"

namespace {

#include "warp_common.cuh"

// phase M: row bitmasks (optional), row bases (optional) and the compact offset list
template <bool kMasks>
__device__ __forceinline__ void build_list(const uint16_t* __restrict__ lab, i64 lab_rs, u32 label, int h, int w, u32 offs_off,
                                        u32 rowmask_off, u32 rowbase_off) {
  unsigned short* offs = reinterpret_cast<unsigned short*>(dyn + offs_off);
  u64* rowmask = reinterpret_cast<u64*>(dyn + rowmask_off);
  unsigned short* rowbase = reinterpret_cast<unsigned short*>(dyn + rowbase_off);
  const u32 lane = lane_id();
  const u32 lt = (1u << lane) - 1u;
  const bool two = w > 32;
  u32 base = 0;
  constexpr int kRows = 8;  // label rows in flight per iteration (16 loads per lane)
#pragma unroll 1
  for (int r0 = 0; r0 < h; r0 += kRows) {
    u32 l0[kRows], l1[kRows];
#pragma unroll
    for (int u = 0; u < kRows; ++u) {  // all loads of the row group first
      const uint16_t* lrow = lab + (i64)(r0 + u) * lab_rs;
      const bool in = r0 + u < h;
      l0[u] = (in && lane < (u32)w) ? (u32)__ldg(lrow + lane) : kFull;
      l1[u] = (in && two && lane + 32 < (u32)w) ? (u32)__ldg(lrow + lane + 32) : kFull;
    }
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const int r = r0 + u;
      if (r >= h) break;
      const bool hit0 = l0[u] == label, hit1 = l1[u] == label;
      const u32 b0 = __ballot_sync(kFull, hit0);
      const u32 b1 = __ballot_sync(kFull, hit1);
      if (kMasks && lane == 0) { rowmask[r] = (u64)b0 | ((u64)b1 << 32); rowbase[r] = (unsigned short)base; }
      if (hit0) offs[base + __popc(b0 & lt)] = (unsigned short)((r << 6) | lane);
      base += __popc(b0);
      if (hit1) offs[base + __popc(b1 & lt)] = (unsigned short)((r << 6) | (lane + 32));
      base += __popc(b1);
    }
  }
  if (kMasks) {  // rows beyond the window read as empty
    if (lane >= (u32)h) rowmask[lane] = 0;
    if (lane + 32 >= (u32)h) rowmask[lane + 32] = 0;
  }
  __syncwarp();
}

// One pixel of a request, Z-reduced (slow path: values that were not staged).
template <typename PX>
__device__ __noinline__ u32 gather_reduced(const PX* __restrict__ p, int Z, i64 z_stride, int red) {
  u32 x = (u32)__ldg(p);
  if (red == ABX_RED_MAX) {
    for (int z = 1; z < Z; ++z) x = max(x, (u32)__ldg(p + (i64)z * z_stride));
  } else {
    for (int z = 1; z < Z; ++z) x += (u32)__ldg(p + (i64)z * z_stride);
  }
  return x;
}

template <typename PX>
struct ValueSource {  // value i of the current request
  const unsigned short* vals;
  const unsigned short* offs;
  const PX* px;
  i64 z_stride;
  u32 rs;
  int Z, red;
  bool staged;
  __device__ __forceinline__ u32 operator()(u32 i) const {
    if (staged) return vals[i];
    const u32 k = offs[i];
    return gather_reduced(px + ((k >> 6) * rs + (k & 63u)), Z, z_stride, red);
  }
};

// Four pixels per lane of the padded list, Z-reduced with max (the only narrow reduction): values of
// list entries i0 + 32 u.  No predicates anywhere: the list is padded with copies of its first entry.
template <typename PX>
__device__ __forceinline__ void gather4(u32 i0, const unsigned short* __restrict__ offs, const PX* __restrict__ px, u32 rs,
                                        i64 z_stride, int Z, u32 (&xq)[4]) {
  u32 go[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const u32 kq = offs[i0 + 32u * u];
    go[u] = (kq >> 6) * rs + (kq & 63u);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) xq[u] = (u32)__ldg(px + go[u]);
  if (Z > 1) {
#pragma unroll 1
    for (int z = 1; z < Z; ++z) {
      const PX* pz = px + (i64)z * z_stride;
#pragma unroll
      for (int u = 0; u < 4; ++u) xq[u] = max(xq[u], (u32)__ldg(pz + go[u]));
    }
  }
}

// pass 1 (narrow values: one plane, or the Z maximum): moments and extrema of one request over the PADDED
// list (n_pad entries, a multiple of 128; entries >= n repeat entry 0, whose contribution is taken out of
// the sums afterwards and which cannot move the extrema).  kDepth gathers of four pixels per lane rotate
// through registers, so that eight loads are in flight while four values are accumulated.  Per value:
// sum, sum of squares (64-bit mad), sum of (x^2 >> bits) by a high multiply — NumPy's wrapped v**2 is
// sum(x^2) - 2^bits * that — min, max; for moment_of_inertia x*c, x*r and x*(c^2 + r^2).  Values are staged
// in shared memory for pass 2 when `stage`.
template <typename PX>
__device__ __forceinline__ void moments_pass(u32 n, u32 n_pad, const unsigned short* __restrict__ offs,
                                             unsigned short* __restrict__ vals, bool stage, const PX* __restrict__ px, u32 rs,
                                             i64 z_stride, int Z, bool want_moi, ChanStats& cs, u32& v0_out) {
  constexpr int kShift = (sizeof(PX) == 1) ? 12 : 8;  // (x << kShift)^2 >> 32 == x^2 >> bits(PX)
  const u32 lane = lane_id();
  u32 f_sum = 0, f_wh = 0, f_m10 = 0, f_m01 = 0;
  u64 f_sq = 0, f_q = 0;
  u32 a_min = kFull, a_max = 0;
  auto accumulate = [&](u32 i0, const u32 (&xq)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const u32 v = xq[u];
      const u32 i = i0 + 32u * u;
      f_sum += v;
      f_sq += (u64)v * (u64)v;
      const u32 a = v << kShift;
      f_wh += __umulhi(a, a);
      a_min = min(a_min, v);
      a_max = max(a_max, v);
      if (want_moi) {
        const u32 k = offs[i];
        const u32 c = k & 63u, r = k >> 6;
        f_m10 += v * c;
        f_m01 += v * r;
        f_q += (u64)v * (u64)(c * c + r * r);
      }
      if (stage) vals[i] = (unsigned short)v;
    }
  };
  // kDepth quads rotate through registers: while one is accumulated, kDepth - 1 gathers (four loads each) are in
  // flight — the loads hit L2 (prefetched one request ahead) and L2 latency is what a warp has to cover
  const u32 nq = n_pad >> 7;  // quads of the list (>= 1)
  u32 x[kDepth][4];
#pragma unroll
  for (int d = 0; d < kDepth; ++d)
    if ((u32)d < nq) gather4<PX>(lane + 128u * d, offs, px, rs, z_stride, Z, x[d]);
  const u32 first = x[0][0];
#pragma unroll 1
  for (u32 q = 0; q < nq; q += kDepth) {
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
      if (q + d < nq) {  // warp-uniform
        const u32 i0 = lane + ((q + d) << 7);
        accumulate(i0, x[d]);
        if (q + d + kDepth < nq) gather4<PX>(i0 + 128u * kDepth, offs, px, rs, z_stride, Z, x[d]);
      }
    }
  }
  // take the padding (n_pad - n copies of entry 0) out of the sums
  const u32 v0 = __shfl_sync(kFull, first, 0);
  const u64 p = (u64)(n_pad - n);
  const u32 k0 = offs[0];
  const u32 c0 = k0 & 63u, r0 = k0 >> 6;
  const u32 a0 = v0 << kShift;
  cs.sum = (u64)__reduce_add_sync(kFull, f_sum) - p * v0;  // n * 65535 < 2^32
  cs.sumsq = warp_sum64(f_sq) - p * ((u64)v0 * v0);
  const u64 wh = (u64)__reduce_add_sync(kFull, f_wh) - p * __umulhi(a0, a0);
  cs.wrapsq = cs.sumsq - (wh << (32 - 2 * kShift));
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
  v0_out = v0;
}

// pass 1, wide values (Z reduction "add": up to 20 bits): 64-bit accumulators over the exact list, nothing staged.
// Out of line; its results travel through the (still unused) histogram area of the slot, so that the caller's
// ChanStats never has to live in local memory for the sake of a by-reference argument.
template <typename PX>
__device__ __noinline__ void moments_wide(u32 n, u32 slot_off, const PX* __restrict__ px, u32 rs, i64 z_stride, int Z, int red,
                                          bool want_moi) {
  const unsigned short* offs = reinterpret_cast<const unsigned short*>(dyn + slot_off);
  const u32 lane = lane_id();
  u64 f_sum = 0, f_sq = 0, f_m10 = 0, f_m01 = 0, f_q = 0;
  u32 a_min = kFull, a_max = 0;
#pragma unroll 1
  for (u32 i = lane; i < n; i += 32) {
    const u32 k = offs[i];
    const u32 c = k & 63u, r = k >> 6;
    const u32 v = gather_reduced(px + (r * rs + c), Z, z_stride, red);
    f_sum += v;
    f_sq += (u64)v * (u64)v;
    a_min = min(a_min, v);
    a_max = max(a_max, v);
    if (want_moi) {
      f_m10 += (u64)v * c;
      f_m01 += (u64)v * r;
      f_q += (u64)v * (u64)(c * c + r * r);
    }
  }
  const u64 sum = warp_sum64(f_sum), sq = warp_sum64(f_sq);
  const u64 m10 = warp_sum64(f_m10), m01 = warp_sum64(f_m01), mq = warp_sum64(f_q);
  const u32 vmin = __reduce_min_sync(kFull, a_min), vmax = __reduce_max_sync(kFull, a_max);
  __syncwarp();
  if (lane == 0) {
    u64* o = reinterpret_cast<u64*>(dyn + slot_off + kStatsHistOff);
    o[0] = sum; o[1] = sq; o[2] = sq; o[3] = m10; o[4] = m01; o[5] = mq; o[6] = 0;
    o[7] = (u64)vmin | ((u64)vmax << 32);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// phase S: one (channel, reduction) request
// ------------------------------------------------------------------------------------------------
template <typename PX>
__device__ __forceinline__ void request_stats(u32 n, u32 n_pad, u32 slot_off, const PX* __restrict__ px, u32 rs, i64 z_stride,
                                              int Z, int reduction, u32 feats, ChanStats* __restrict__ dst) {
  const unsigned short* offs = reinterpret_cast<const unsigned short*>(dyn + slot_off);
  unsigned short* vals = reinterpret_cast<unsigned short*>(dyn + slot_off) + kCapSmall;
  u32* hist = reinterpret_cast<u32*>(dyn + slot_off + kStatsHistOff);
  u32* t = reinterpret_cast<u32*>(dyn + slot_off + kStatsTOff);
  const u32 lane = lane_id();
  const bool wide = (reduction == ABX_RED_ADD && Z > 1) || sizeof(PX) == 4;  // uint32 pixels: Z-add sum planes (zreduce.cu)
  const bool staged = !wide && n_pad <= (u32)kCapSmall;  // larger lists occupy the staging area themselves
  ChanStats cs;
  u32 v0 = 0;
  if (wide) {
    moments_wide<PX>(n, slot_off, px, rs, z_stride, Z, reduction, (feats & ABX_F_MOI) != 0);
    const u64* o = reinterpret_cast<const u64*>(hist);
    cs.sum = o[0]; cs.sumsq = o[1]; cs.wrapsq = o[2]; cs.m10 = o[3]; cs.m01 = o[4]; cs.m20 = o[5]; cs.m02 = o[6];
    cs.vmin = (u32)o[7]; cs.vmax = (u32)(o[7] >> 32);
    __syncwarp();
  } else {
    moments_pass<PX>(n, n_pad, offs, vals, staged, px, rs, z_stride, Z, (feats & ABX_F_MOI) != 0, cs, v0);
  }
  cs.med_lo = cs.med_hi = 0;
  cs.top2p5_sum = cs.top5_sum = 0;

  if (feats & (ABX_F_MEDIAN | ABX_F_TOP2P5 | ABX_F_TOP5)) {
    ValueSource<PX> value{vals, offs, px, z_stride, rs, Z, reduction, staged};
    const u32 vmin = cs.vmin;
    // ---- pass 2: range-adaptive histogram ----
    const u32 range = cs.vmax - vmin;
    int s0 = 0;
    while ((range >> s0) >= (u32)kBins) ++s0;
    const u32 nb = (range >> s0) + 1;
    __syncwarp();
    hist_zero(hist, 32u * bins_per_lane(nb));
    __syncwarp();
    const u32 hbase = smem_addr(hist);
    if (staged) {  // staged values, two per 32-bit shared-memory load
      const u32* v32 = reinterpret_cast<const u32*>(vals);
      const u32 pairs = n >> 1;
      if (s0 == 0) {
        const u32 base = hbase - 4u * vmin;  // bin address = 4 * value + base
#pragma unroll 4
        for (u32 w = lane; w < pairs; w += 32) {
          const u32 xx = v32[w];
          hist_inc(4u * (xx & 0xFFFFu) + base);
          hist_inc(4u * (xx >> 16) + base);
        }
      } else {
#pragma unroll 4
        for (u32 w = lane; w < pairs; w += 32) {
          const u32 xx = v32[w];
          hist_inc(hbase + 4u * (((xx & 0xFFFFu) - vmin) >> s0));
          hist_inc(hbase + 4u * (((xx >> 16) - vmin) >> s0));
        }
      }
      if ((n & 1u) && lane == 0) hist_inc(hbase + 4u * (((u32)vals[n - 1] - vmin) >> s0));
    } else if (!wide) {  // list too long to stage: gather again (the lines are in L1 / L2), padding taken out afterwards
#pragma unroll 1
      for (u32 i0 = lane; i0 < n_pad; i0 += 128) {
        u32 xq[4];
        gather4<PX>(i0, offs, px, rs, z_stride, Z, xq);
#pragma unroll
        for (int u = 0; u < 4; ++u) hist_inc(hbase + 4u * ((xq[u] - vmin) >> s0));
      }
      __syncwarp();
      if (lane == 0) hist[(v0 - vmin) >> s0] -= n_pad - n;
    } else {
#pragma unroll 1
      for (u32 i = lane; i < n; i += 32) hist_add(hist, (value(i) - vmin) >> s0);
    }
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
#pragma unroll 2
        for (u32 i = lane; i < n; i += 32) {
          const u32 d = value(i) - vmin;
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
#pragma unroll 2
        for (u32 i = lane; i < n; i += 32) {
          const u32 x = value(i);
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

template <typename PX>
__global__ void __launch_bounds__(kStatsWarps * 32, 2)
object_stats_warp(const Common cm, const PX* __restrict__ pixels, const i64* __restrict__ tile_offset, i64 chan_stride,
                  i64 z_stride, i64 px_row_stride, int Z, const abx_request* __restrict__ requests, int n_requests,
                  ChanStats* __restrict__ chan, int* __restrict__ stats_list, u32* __restrict__ stats_count,
                  const u32* __restrict__ todo_count, const int* __restrict__ pair_list, const u32* __restrict__ gate,
                  int bookkeeping) {
  // gate: the launch has nothing to do when *gate == 0 (zreduce.cu: no request left to the stack).  bookkeeping = 0:
  // another kernel already zero-filled the absent labels and handed the large objects over.
  if (gate != nullptr && *gate == 0u) return;
  // Two modes.  todo_count == nullptr: every object of the launch, all its requests (the path for layouts TMA cannot
  // address).  Otherwise: the (object, request) pairs the sweep kernel (object_sweep.cu) left over, one work item per
  // pair so that the few of them finish in the time of one request.
  const u32 lane = lane_id();
  const u32 slot_off = (threadIdx.x >> 5) * kStatsSlot;
  const bool by_list = todo_count != nullptr;
  const int n_items = by_list ? (int)(*todo_count) : cm.n_total;
  Queue qu{cm.counters, n_items, 0};
  int item = qu.fetch();
  int nxt = item < n_items ? qu.fetch() : n_items;
  while (item < n_items) {
    int obj = item, q_lo = 0, q_hi = n_requests;
    if (by_list) {
      const int pair = pair_list[item];
      obj = pair / n_requests;
      q_lo = pair - obj * n_requests;
      q_hi = q_lo + 1;
    }
    const abx_object_rec rec = cm.recs[obj];
    const bool is_bg = obj >= cm.n_objects;
    const int h = (int)(rec.rmax - rec.rmin) + 1, w = (int)(rec.cmax - rec.cmin) + 1;
    if (rec.n == 0) {  // absent label (or empty background): zero records -> NaN in finalize
      for (int q = lane; bookkeeping && q < n_requests; q += 32) {
        ChanStats z;
        z.sum = z.sumsq = z.wrapsq = z.m10 = z.m01 = z.m20 = z.m02 = z.top2p5_sum = z.top5_sum = 0;
        z.vmin = z.vmax = z.med_lo = z.med_hi = 0;
        chan[(i64)obj * n_requests + q] = z;
      }
    } else if (is_bg || h > kSide || w > kSide) {  // hand over to the CTA-per-object kernel
      if (bookkeeping && lane == 0) stats_list[atomicAdd(stats_count, 1u)] = obj;
    } else {
      const int p = find_plane(cm.plane_base, cm.n_planes, obj);
      __syncwarp();
      if (by_list)  // nothing was prefetched for this item: start its one request towards L2 before the label loads
        prefetch_request<PX>(pixels + tile_offset[cm.plane_tile[p]] + (i64)rec.rmin * px_row_stride + rec.cmin +
                                 (i64)requests[q_lo].channel * chan_stride,
                             px_row_stride, z_stride, Z, h, w);
      build_list<false>(cm.labels + (i64)p * cm.lab_plane_stride + (i64)rec.rmin * cm.lab_row_stride + rec.cmin,
                        cm.lab_row_stride, (u32)(obj - cm.plane_base[p] + 1), h, w, slot_off, 0, 0);
      // pad the list to a whole number of 128-pixel steps with copies of its first entry
      const u32 n_pad = (rec.n + (u32)kPad - 1u) & ~((u32)kPad - 1u);
      {
        unsigned short* offs = reinterpret_cast<unsigned short*>(dyn + slot_off);
        const unsigned short first = offs[0];
        for (u32 i = rec.n + lane; i < n_pad; i += 32) offs[i] = first;
        __syncwarp();
      }
      const PX* px0 = pixels + tile_offset[cm.plane_tile[p]] + (i64)rec.rmin * px_row_stride + rec.cmin;
#pragma unroll 1
      for (int q = q_lo; q < q_hi; ++q) {
        const abx_request rq = requests[q];
        // L2 prefetch one request ahead (a longer distance does not survive: at 2 TB/s the 126 MB L2 turns over in
        // the time a warp spends on one object): the next request of this object, or the label window and the first
        // request of the warp's next object
        if (q + 1 < q_hi) {
          prefetch_request<PX>(px0 + (i64)requests[q + 1].channel * chan_stride, px_row_stride, z_stride, Z, h, w);
        } else if (!by_list && nxt < cm.n_objects) {
          const abx_object_rec nr = cm.recs[nxt];
          const int nh = (int)(nr.rmax - nr.rmin) + 1, nw = (int)(nr.cmax - nr.cmin) + 1;
          if (nr.n > 0 && nh <= kSide && nw <= kSide) {
            const int np = find_plane(cm.plane_base, cm.n_planes, nxt);
            prefetch_rows(cm.labels + (i64)np * cm.lab_plane_stride + (i64)nr.rmin * cm.lab_row_stride + nr.cmin,
                          cm.lab_row_stride * 2, nh, (u32)nw * 2u);
            prefetch_request<PX>(pixels + tile_offset[cm.plane_tile[np]] + (i64)nr.rmin * px_row_stride + nr.cmin +
                                     (i64)requests[0].channel * chan_stride,
                                 px_row_stride, z_stride, Z, nh, nw);
          }
        }
        if (rq.reduction == ABX_RED_DIV) continue;  // floating-point request: object_float.cu
        request_stats<PX>(rec.n, n_pad, slot_off, px0 + (i64)rq.channel * chan_stride, (u32)px_row_stride, z_stride, Z,
                          rq.reduction, rq.features, chan + (i64)obj * n_requests + q);
      }
    }
    item = nxt;
    nxt = item < n_items ? qu.fetch() : n_items;
  }
}

template <typename K>
int set_smem(K kernel, size_t smem, bool* done) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return abx_check_cuda(e, "object_warp smem attribute");
    done[dev] = true;
  }
  return ABX_OK;
}

template <typename PX>
int launch_stats(const abx_extract_args* a, const Workspace& ws, const Common& cm, cudaStream_t st, bool todo,
                 const abx_request* requests, const u32* gate, int bookkeeping) {
  constexpr size_t smem = (size_t)kStatsWarps * kStatsSlot;
  static thread_local bool done[64] = {false};
  int rc = set_smem(object_stats_warp<PX>, smem, done);
  if (rc) return rc;
  int grid = (cm.n_total + kStatsWarps - 1) / kStatsWarps;
  if (grid > 148 * 2) grid = 148 * 2;  // persistent: 2 CTAs per SM, warps pull objects from a counter
  // todo: the length of the list is only known on the device — a fixed grid of single-warp CTAs (12 KB of shared memory
  // each, so that they fit next to the CTAs of the shape kernel running on the caller's stream)
  if (todo) grid = 148;
  object_stats_warp<PX><<<grid, todo ? 32 : kStatsWarps * 32, todo ? (size_t)kStatsSlot : smem, st>>>(
      cm, static_cast<const PX*>(a->pixels), reinterpret_cast<const i64*>(a->tile_offset), a->chan_stride, a->z_stride,
      a->row_stride, a->Z, requests, a->n_requests, ws.chan, ws.stats_list, ws.list_counts + kCntStatsList,
      todo ? ws.list_counts + kCntLeftover : nullptr, ws.pair_list, gate, bookkeeping);
  return abx_check_cuda(cudaGetLastError(), "object_stats_warp");
}

}  // namespace

static Common make_common(const abx_extract_args* a, const Workspace& ws, int n_total) {
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
  cm.counters = nullptr;
  return cm;
}

// todo = false: every window-sized object; todo = true: the (object, request) pairs the sweep kernel left over
int launch_object_stats_warp(const abx_extract_args* a, const Workspace& ws, cudaStream_t st, bool todo) {
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  if (n_total == 0 || a->n_requests == 0) return ABX_OK;
  Common cm = make_common(a, ws, n_total);
  cm.counters = ws.list_counts + (todo ? kCntLeftoverWork : kCntGather);
  if (a->pixel_dtype == ABX_U16) return launch_stats<uint16_t>(a, ws, cm, st, todo, a->requests, nullptr, 1);
  if (a->pixel_dtype == ABX_U8) return launch_stats<uint8_t>(a, ws, cm, st, todo, a->requests, nullptr, 1);
  return ABX_OK;  // float pixels: every request belongs to object_float.cu
}

// After a Z-reduced launch (zreduce.cu): the Z-add requests of every window-sized object, from their uint32 sum planes
// (one value per pixel instead of Z).  Does nothing (one gate load per CTA) when there is no such request.
int launch_object_stats_rest(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  const int n_total = a->n_objects + (a->with_background ? a->n_planes : 0);
  if (n_total == 0 || a->n_requests == 0) return ABX_OK;
  abx_extract_args sums = *a;  // the sum planes as a dense (tiles, requests, 1, H, W) uint32 array
  sums.pixels = ws.zplanes;
  sums.tile_offset = reinterpret_cast<const int64_t*>(ws.ztile_offset + a->n_tiles);
  sums.chan_stride = sums.z_stride = (i64)a->H * a->W;
  sums.row_stride = a->W;
  sums.Z = 1;
  Common cm = make_common(a, ws, n_total);
  cm.counters = ws.list_counts + kCntRest;
  return launch_stats<uint32_t>(&sums, ws, cm, st, false, ws.req_rest, ws.zflags, 0);
}
