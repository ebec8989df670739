// Device helpers shared by the warp-per-object kernels (object_warp.cu, object_sweep.cu, object_edt.cu, object_pair.cu).  Include inside the
// translation unit's anonymous namespace after common.cuh.
#pragma once

constexpr int kSide = 64;        // maximum window side
constexpr int kCapSmall = 2048;  // gather kernel: objects up to this many pixels stage their values
constexpr int kCapLarge = 4096;  // pixels per object, large slot (= kSide * kSide)
constexpr int kBins = 1024;      // level-0 histogram bins (32-bit counters)
constexpr int kStatsWarps = 9;   // nine identical 12.1 KB slots: 109 KB per CTA, 2 CTAs per SM
constexpr int kDepth = 3;        // gathers of four pixels per lane in the pass-1 pipeline (code size: the I-cache is the limit)
constexpr int kPad = 128;        // the offset list is padded to a multiple of this (four pixels per lane)
// statistics slot: offs u16[4096] (objects <= kCapSmall pixels: offs u16[2048] | vals u16[2048]) | hist u32[1024] | t u32[16]
constexpr u32 kStatsHistOff = kCapLarge * 2, kStatsTOff = kStatsHistOff + kBins * 4, kStatsSlot = kStatsTOff + 64;
constexpr u32 kFull = 0xFFFFFFFFu;

// Every device function derives its shared-memory pointers from this array plus a byte offset, so that the
// compiler keeps them in the shared address space (generic pointers passed through __noinline__ calls
// compiled to LD.E/ST.E with 64-bit address arithmetic: 24 instructions per EDT step instead of 8).
extern __shared__ __align__(128) unsigned char dyn[];

// Sum of a 64-bit quantity over the warp from three independent 32-bit REDUX reductions of its
// 24/24/16-bit slices (each slice sum < 2^29): shorter and far less latency than five dependent
// 64-bit shuffle steps.  Exact modulo 2^64 for any input.
__device__ __forceinline__ u64 warp_sum64(u64 v) {
  const u32 a = __reduce_add_sync(kFull, (u32)v & 0xFFFFFFu);
  const u32 b = __reduce_add_sync(kFull, (u32)(v >> 24) & 0xFFFFFFu);
  const u32 c = __reduce_add_sync(kFull, (u32)(v >> 48));
  return (u64)a + ((u64)b << 24) + ((u64)c << 48);
}

// ------------------------------------------------------------------------------------------------
// Work distribution: persistent warps pull work items from a global counter, one atomic per item; a warp always
// holds the NEXT item too, so that it can prefetch that item's windows into L2 while it works.
// ------------------------------------------------------------------------------------------------
struct Queue {
  u32* counters;  // [2]
  int n_total;
  int phase;      // which of the two counters (always 0 since the slots became uniform)
  __device__ __forceinline__ int fetch() {
    int v = 0;
    if (lane_id() == 0) v = (int)atomicAdd(&counters[phase], 1u);
    return __shfl_sync(kFull, v, 0);
  }
};

// ------------------------------------------------------------------------------------------------
// phase S helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
// one count at a shared-memory byte address (no return value: a reduction, not an atomic round trip)
__device__ __forceinline__ void hist_inc(u32 addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void hist_add(u32* hist, u32 bin) { hist_inc(smem_addr(hist) + 4u * bin); }
__device__ __forceinline__ void hist_zero(u32* hist, u32 n_bins) {  // n_bins: a multiple of 4
  uint4* h4 = reinterpret_cast<uint4*>(hist);
#pragma unroll 1
  for (u32 k = lane_id(); k < n_bins / 4u; k += 32) h4[k] = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ u32 bins_per_lane(u32 nb) { return (((nb + 31u) >> 5) + 7u) & ~7u; }  // 8, 16, 24 or 32

// Locate four ranks in the histogram h[0, nb).  Level A: each lane sums `per` consecutive bins (a multiple
// of 8, read as uint4) and a warp scan finds the owning lane; level B: eight lanes per rank scan the owner's
// bins.  For rank t: key = its bin, rank = t - (count below the bin), cnt = count below the bin,
// cb = sum over the bins below of count * bin index.  out: t[0..16).  Bins [nb, 32 * per) must be zero.
__device__ __forceinline__ void find_ranks32(const u32* h, u32 nb, const u32 (&ranks)[4], u32* t) {
  const u32 lane = lane_id();
  const u32 per = bins_per_lane(nb);
  const u32 b0 = lane * per;
  u32 cnt = 0, cb = 0;
#pragma unroll 1
  for (u32 k = 0; k < per; k += 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(h + b0 + k);
    const uint4 w = *reinterpret_cast<const uint4*>(h + b0 + k + 4);
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
  // level B, the four ranks at once: eight lanes per rank walk the (<= 32) bins of the rank's owner lane
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
  const u32 lb = own * per + sub * nper;  // first bin of this lane
  u32 c4[4], lc = 0, lq = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c4[k] = ((u32)k < nper) ? h[lb + k] : 0u;
    lc += c4[k];
    lq += c4[k] * (lb + k);
  }
  u32 ic = lc, iq = lq;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const u32 a = __shfl_up_sync(kFull, ic, o, 8);
    const u32 b = __shfl_up_sync(kFull, iq, o, 8);
    if (sub >= (u32)o) { ic += a; iq += b; }
  }
  u32 acc = e + ic - lc, accq = eb + iq - lq;  // counts / weighted counts below this lane's first bin
  if (tr >= acc && tr < acc + lc) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (tr >= acc && tr < acc + c4[k]) {
        t[grp] = lb + k;          // key
        t[4 + grp] = tr - acc;    // rank inside the bin
        t[8 + grp] = acc;         // count below
        t[12 + grp] = accq;       // sum(count * bin) below
      }
      acc += c4[k];
      accq += c4[k] * (lb + k);
    }
  }
  __syncwarp();
}

// Four independent rank searches at once: eight lanes per 128-bin sub-histogram (refinement).
// in: t[4 + g] = rank inside group g; out: t[g] = sub-bin, t[4 + g] = rank inside the sub-bin.
__device__ __forceinline__ void find_ranks32_x4(const u32* h, u32* t) {
  const u32 lane = lane_id();
  const u32 grp = lane >> 3, sub = lane & 7u;
  const u32 b0 = grp * 128u + sub * 16u;
  u32 cnt = 0;
#pragma unroll
  for (int k = 0; k < 16; k += 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(h + b0 + k);
    cnt += v.x + v.y + v.z + v.w;
  }
  u32 icnt = cnt;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const u32 c = __shfl_up_sync(kFull, icnt, o, 8);
    if (sub >= (u32)o) icnt += c;
  }
  const u32 ecnt = icnt - cnt;
  const u32 tr = t[4 + grp];
  __syncwarp();
  if (tr >= ecnt && tr < ecnt + cnt) {
    u32 acc = ecnt;
#pragma unroll 1
    for (u32 bq = b0; bq < b0 + 16u; ++bq) {
      const u32 c = h[bq];
      if (tr < acc + c) { t[grp] = bq - grp * 128u; t[4 + grp] = tr - acc; break; }
      acc += c;
    }
  }
  __syncwarp();
}

// L2 prefetch of window rows (lane <-> row): one request per 32-byte sector the row can touch, the last one clamped
// to the row's own last byte.  (A bulk prefetch per row — cp.async.bulk.prefetch.L2 — takes its address from uniform
// registers and compiles to a lane-by-lane loop: 280 instead of 25 instructions per window.)
__device__ __forceinline__ void prefetch_rows(const void* first_row, i64 row_stride_bytes, int nh, u32 row_bytes) {
#ifndef ABX_NO_PREFETCH
  const u32 ns = (row_bytes + 62u) >> 5;  // sectors a row of row_bytes at any alignment can span
  for (int r = (int)lane_id(); r < nh; r += 32) {
    const char* a = static_cast<const char*>(first_row) + (i64)r * row_stride_bytes;
    const char* last = a + (row_bytes - 1u);
#pragma unroll 1
    for (u32 s = 0; s < ns; ++s) {
      const char* q = a + 32u * s;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q < last ? q : last));
    }
  }
#endif
}
// the window of one request (every Z plane it reduces)
template <typename PX>
__device__ __forceinline__ void prefetch_request(const PX* __restrict__ px, i64 px_rs, i64 z_stride, int Z, int nh, int nw) {
  const int zmax = Z < 16 ? Z : 16;
  for (int z = 0; z < zmax; ++z) prefetch_rows(px + (i64)z * z_stride, px_rs * (i64)sizeof(PX), nh, (u32)nw * (u32)sizeof(PX));
}

struct Common {  // kernel arguments shared by both kernels
  const uint16_t* labels;
  i64 lab_plane_stride, lab_row_stride;
  const int32_t* plane_tile;
  const int32_t* plane_base;
  int n_planes, n_objects, n_total;
  const abx_object_rec* recs;
  u32* counters;  // [2] work counters of this kernel
};

