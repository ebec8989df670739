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
// One CTA of 4 warps per (object, pair); warps over the rows of the bounding box, lanes over its columns; the object is
// where the label plane holds its id (objects of any size: nothing here depends on the 64 x 64 bitmaps).  Z stacks are
// reduced per pixel on the fly (max or add).  Pixels come from global memory / L2 (twice when ranks are wanted): this
// kernel is the "next" row of the hot path, built for parity first — see DESIGN.md for its measured cost.
#include "common.cuh"

namespace {

constexpr int kPairWarps = 4;
constexpr int kPairThreads = kPairWarps * 32;
constexpr u32 kRankWords = 2048;  // 65 536 presence bits

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

__device__ __forceinline__ u64 warp_sum(u64 v) {
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, k);
  return v;
}

template <typename PX>
__global__ void __launch_bounds__(kPairThreads)
object_pair_kernel(const PairArgs a) {
  __shared__ u32 bits[2][kRankWords];
  __shared__ uint16_t prefix[2][kRankWords];
  __shared__ u64 partial[kPairWarps][10];
  __shared__ u32 scan_tot[2][kPairWarps];
  __shared__ u32 n_both_w[kPairWarps];
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const i64 n_items = (i64)a.n_objects * a.n_pairs;
  for (i64 item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int obj = (int)(item / a.n_pairs), pi = (int)(item - (i64)obj * a.n_pairs);
    const abx_object_rec rec = a.recs[obj];
    const abx_pair pr = a.pairs[pi];
    PairStats ps;
    memset(&ps, 0, sizeof(ps));
    const ChanStats* ca = a.chan + (i64)obj * a.n_requests + pr.request_a;
    const ChanStats* cb = a.chan + (i64)obj * a.n_requests + pr.request_b;
    const u32 amin = ca->vmin, amax = ca->vmax, bmin = cb->vmin, bmax = cb->vmax;
    if (rec.n == 0 || amax >= 65536u || bmax >= 65536u) {  // (block-uniform)
      if (rec.n) {
        ps.flags = 1u;
        if (threadIdx.x == 0) atomicOr(a.err, 4u);
      }
      if (threadIdx.x == 0) a.out[item] = ps;
      continue;
    }
    const int plane = find_plane(a.plane_base, a.n_planes, obj);
    const u32 id = (u32)(obj - a.plane_base[plane]) + 1u;
    const uint16_t* lab = a.labels + (i64)plane * a.label_plane_stride;
    const abx_request qa = a.requests[pr.request_a], qb = a.requests[pr.request_b];
    const PX* base = static_cast<const PX*>(a.pixels) + a.tile_offset[a.plane_tile[plane]];
    const PX* pa = base + (i64)qa.channel * a.chan_stride;
    const PX* pb = base + (i64)qb.channel * a.chan_stride;
    const double tx = pr.threshold_fraction * (double)amax, ty = pr.threshold_fraction * (double)bmax;
    const bool ranks = (pr.features & ABX_PF_RWC) != 0;
    u32 big_r = 0;
    if (ranks) {
      // ---- presence bits of both images over [vmin, vmax], then the exclusive prefix of the word popcounts ----
      const u32 wa = (amax - amin) / 32u + 1u, wb = (bmax - bmin) / 32u + 1u;
      for (u32 i = threadIdx.x; i < wa; i += kPairThreads) bits[0][i] = 0;
      for (u32 i = threadIdx.x; i < wb; i += kPairThreads) bits[1][i] = 0;
      __syncthreads();
      for (u32 r = rec.rmin + warp; r <= rec.rmax; r += kPairWarps) {
        for (u32 c = rec.cmin + lane; c <= rec.cmax; c += 32u) {
          if (lab[(i64)r * a.label_row_stride + c] != id) continue;
          const i64 off = (i64)r * a.row_stride + c;
          const u32 x = value_at(pa + off, a.Z, a.z_stride, qa.reduction) - amin;
          const u32 y = value_at(pb + off, a.Z, a.z_stride, qb.reduction) - bmin;
          atomicOr(&bits[0][x >> 5], 1u << (x & 31u));
          atomicOr(&bits[1][y >> 5], 1u << (y & 31u));
        }
      }
      __syncthreads();
#pragma unroll
      for (int im = 0; im < 2; ++im) {
        const u32 nw = im ? wb : wa;
        // thread t owns the words [16 t, 16 t + 16)
        u32 mine = 0;
        for (u32 k = 0; k < 16u; ++k) {
          const u32 i = threadIdx.x * 16u + k;
          if (i < nw) mine += __popc(bits[im][i]);
        }
        u32 incl = mine;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
          const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, k);
          if ((int)lane >= k) incl += t;
        }
        if (lane == 31u) scan_tot[im][warp] = incl;
        __syncthreads();
        u32 before = incl - mine;
        for (u32 w = 0; w < warp; ++w) before += scan_tot[im][w];
        for (u32 k = 0; k < 16u; ++k) {
          const u32 i = threadIdx.x * 16u + k;
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
    // ---- the sums ----
    u64 sxy = 0, tot_x = 0, tot_y = 0, cx = 0, cy = 0, cxy = 0, cxx = 0, cyy = 0, wx = 0, wy = 0;
    u32 n_both = 0;
    for (u32 r = rec.rmin + warp; r <= rec.rmax; r += kPairWarps) {
      for (u32 c = rec.cmin + lane; c <= rec.cmax; c += 32u) {
        if (lab[(i64)r * a.label_row_stride + c] != id) continue;
        const i64 off = (i64)r * a.row_stride + c;
        const u32 x = value_at(pa + off, a.Z, a.z_stride, qa.reduction);
        const u32 y = value_at(pb + off, a.Z, a.z_stride, qb.reduction);
        const u64 xy = (u64)x * y;
        sxy += xy;
        const bool ox = (double)x >= tx, oy = (double)y >= ty;
        if (ox) tot_x += x;
        if (oy) tot_y += y;
        if (ox && oy) {
          ++n_both;
          cx += x;
          cy += y;
          cxy += xy;
          cxx += (u64)x * x;
          cyy += (u64)y * y;
          if (ranks) {
            const u32 xr = x - amin, yr = y - bmin;
            const u32 ra = prefix[0][xr >> 5] + __popc(bits[0][xr >> 5] & ((1u << (xr & 31u)) - 1u));
            const u32 rb = prefix[1][yr >> 5] + __popc(bits[1][yr >> 5] & ((1u << (yr & 31u)) - 1u));
            const u32 wgt = big_r - (ra > rb ? ra - rb : rb - ra);
            wx += (u64)x * wgt;
            wy += (u64)y * wgt;
          }
        }
      }
    }
    u64 v[10] = {sxy, tot_x, tot_y, cx, cy, cxy, cxx, cyy, wx, wy};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const u64 s = warp_sum(v[k]);
      if (lane == 0) partial[warp][k] = s;
    }
    const u32 nb = __reduce_add_sync(0xFFFFFFFFu, n_both);
    if (lane == 0) n_both_w[warp] = nb;
    __syncthreads();
    if (threadIdx.x == 0) {
      u64 t[10];
      for (int k = 0; k < 10; ++k) {
        t[k] = 0;
        for (int w = 0; w < kPairWarps; ++w) t[k] += partial[w][k];
      }
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
  const unsigned grid = (unsigned)(n_items < 148 * 16 ? n_items : 148 * 16);  // 16 CTAs of 4 warps per SM, grid-stride
  if (a->pixel_dtype == ABX_U8) object_pair_kernel<uint8_t><<<grid, kPairThreads, 0, st>>>(p);
  else object_pair_kernel<uint16_t><<<grid, kPairThreads, 0, st>>>(p);
  return abx_check_cuda(cudaGetLastError(), "object_pair");
}
