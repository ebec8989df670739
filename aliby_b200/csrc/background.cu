// Per-plane background (label 0) of LARGE planes: median and top-5 mean of the pixels outside every cell
// (trap.py:6-43: imBackground, background_max5) from one streaming pass.
//
// The per-object kernels treat a segment as one unit of work, which is right for cells and for the 96 x 96 trap
// tiles of the yeast pipelines, but the background of a whole 2160 x 2160 field is a 2-million-pixel segment: one CTA
// sweeping it several times takes milliseconds.  For integer pixels of at most 16 bits the complete value
// distribution fits a 65 536-bin histogram, from which both order statistics are exact:
//
//   bg_hist_kernel    all SMs stream the label + pixel planes once (Z reduction fused), one global atomic per
//                     background pixel and request into hist[(plane, request)][value]
//   bg_select_kernel  one CTA per (plane, request): ranks (n-1)/2, n/2 and the sum of the 5 largest values
//
// Planes of at most kBigBackground pixels (the 96 x 96 trap tiles of the yeast pipelines) take bg_tile_kernel: one CTA per
// (plane, request) with the whole 65 536-value histogram in shared memory as packed 16-bit counters (a plane has at most
// 16 384 pixels) — a time point of 40 tiles used to spend 0.32 of its 0.42 ms on these backgrounds in the generic
// CTA-per-object kernel.  Wide values (Z-add of more than one plane) stay with object_stats.cu.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kValues = 65536;

template <typename PX>
__global__ void __launch_bounds__(kThreads)
bg_hist_kernel(const uint16_t* __restrict__ labels, i64 lab_plane_stride, i64 lab_row_stride,
               const int32_t* __restrict__ plane_tile, int H, int W, const PX* __restrict__ pixels,
               const i64* __restrict__ tile_offset, i64 chan_stride, i64 z_stride, i64 px_row_stride, int Z,
               const abx_request* __restrict__ requests, int n_requests, u32* __restrict__ hist) {
  const int p = blockIdx.y;
  const uint16_t* lab = labels + (i64)p * lab_plane_stride;
  const PX* px = pixels + tile_offset[plane_tile[p]];
  const i64 n = (i64)H * W;
  for (i64 i = (i64)blockIdx.x * kThreads + threadIdx.x; i < n; i += (i64)gridDim.x * kThreads) {
    const int r = (int)(i / W), c = (int)(i - (i64)r * W);
    if (__ldg(lab + (i64)r * lab_row_stride + c) != 0) continue;
    for (int q = 0; q < n_requests; ++q) {
      const abx_request rq = requests[q];
      if (rq.bg_features == 0 || rq.reduction == ABX_RED_DIV || (rq.reduction == ABX_RED_ADD && Z > 1)) continue;
      const PX* s = px + (i64)rq.channel * chan_stride + (i64)r * px_row_stride + c;
      u32 x = (u32)__ldg(s);
      for (int z = 1; z < Z; ++z) x = max(x, (u32)__ldg(s + (i64)z * z_stride));
      atomicAdd(&hist[((i64)p * n_requests + q) * kValues + x], 1u);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
bg_select_kernel(const u32* __restrict__ hist, const abx_request* __restrict__ requests, int n_requests, int Z,
                 int n_objects, const abx_object_rec* __restrict__ recs, ChanStats* __restrict__ out) {
  __shared__ u32 part[kThreads];
  __shared__ u32 s_med[2];
  __shared__ u64 s_top;
  const int p = blockIdx.y, q = blockIdx.x;
  const abx_request rq = requests[q];
  if (rq.bg_features == 0 || rq.reduction == ABX_RED_DIV || (rq.reduction == ABX_RED_ADD && Z > 1)) return;
  const u32* h = hist + ((i64)p * n_requests + q) * kValues;
  const u32 n = recs[n_objects + p].n;
  ChanStats* dst = out + (i64)(n_objects + p) * n_requests + q;
  constexpr int kPer = kValues / kThreads;  // 256 consecutive values per thread
  const u32 v0 = threadIdx.x * kPer;
  u32 cnt = 0;
  for (int k = 0; k < kPer; ++k) cnt += h[v0 + k];
  part[threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {  // exclusive prefix over 256 partial counts; tiny
    u32 acc = 0;
    for (int t = 0; t < kThreads; ++t) { const u32 c = part[t]; part[t] = acc; acc += c; }
    s_top = 0;
  }
  __syncthreads();
  if (n == 0) {
    if (threadIdx.x == 0) {
      ChanStats z;
      z.sum = z.sumsq = z.wrapsq = z.m10 = z.m01 = z.m20 = z.m02 = z.top2p5_sum = z.top5_sum = 0;
      z.vmin = z.vmax = z.med_lo = z.med_hi = 0;
      *dst = z;
    }
    return;
  }
  const u32 below = part[threadIdx.x];
  const u32 ranks[2] = {(n - 1) / 2, n / 2};
  const u32 k5 = n < 5u ? n : 5u;
  const u32 top_from = n - k5;  // elements with rank >= top_from are the k5 largest
  u32 acc = below;
  u64 top = 0;
  for (int k = 0; k < kPer; ++k) {
    const u32 c = h[v0 + k];
    if (c) {
      for (int j = 0; j < 2; ++j)
        if (ranks[j] >= acc && ranks[j] < acc + c) s_med[j] = v0 + k;
      if (acc + c > top_from) {  // part of this bin belongs to the top k5
        const u32 first = acc > top_from ? acc : top_from;
        top += (u64)(acc + c - first) * (u64)(v0 + k);
      }
    }
    acc += c;
  }
  if (top) atomicAdd(reinterpret_cast<unsigned long long*>(&s_top), (unsigned long long)top);
  __syncthreads();
  if (threadIdx.x == 0) {
    ChanStats z;
    z.sum = z.sumsq = z.wrapsq = z.m10 = z.m01 = z.m20 = z.m02 = z.top2p5_sum = 0;
    z.vmin = z.vmax = 0;
    z.med_lo = s_med[0]; z.med_hi = s_med[1];
    z.top5_sum = s_top;
    *dst = z;
  }
}

// Small planes: (plane, request) per CTA.  The plane's background values are loaded once (batches of eight loads per
// thread in flight) and parked in shared memory; their range then sizes the histogram — packed 16-bit counters, a plane
// has at most 16 384 pixels — so that zeroing and scanning it cost what the value range costs, not 65 536 bins.
constexpr int kTilePixels = (int)kBigBackground;  // 16 384
constexpr int kChunkValues = 8192;                // values one histogram round covers (16 KB of packed counters)

template <typename PX>
__global__ void __launch_bounds__(kThreads)
bg_tile_kernel(const uint16_t* __restrict__ labels, i64 lab_plane_stride, i64 lab_row_stride,
               const int32_t* __restrict__ plane_tile, int H, int W, const PX* __restrict__ pixels,
               const i64* __restrict__ tile_offset, i64 chan_stride, i64 z_stride, i64 px_row_stride, int Z,
               const abx_request* __restrict__ requests, int n_requests, int n_objects,
               const abx_object_rec* __restrict__ recs, ChanStats* __restrict__ out, int px_slots) {
  // shared memory sized by the launch for the planes' pixel count (px_slots = H W rounded up to 16): a 96 x 96 tile
  // takes 16 + 18 + 9 KB and five CTAs share an SM
  extern __shared__ __align__(16) u32 h32[];                                      // kChunkValues / 2 words
  unsigned short* vals = reinterpret_cast<unsigned short*>(h32 + kChunkValues / 2);  // [px_slots]
  unsigned char* isbg = reinterpret_cast<unsigned char*>(vals + px_slots);      // [px_slots]
  __shared__ u32 part[kThreads];
  __shared__ u32 s_lo[kThreads / 32], s_hi[kThreads / 32];
  __shared__ u32 s_med[2], s_total;
  __shared__ u64 s_top;
  const int p = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
  const abx_request rq = requests[q];
  if (rq.bg_features == 0 || rq.reduction == ABX_RED_DIV || (rq.reduction == ABX_RED_ADD && Z > 1)) return;
  const u32 n = recs[n_objects + p].n;
  ChanStats* dst = out + (i64)(n_objects + p) * n_requests + q;
  if (n == 0) {
    if (tid == 0) {
      ChanStats z;
      z.sum = z.sumsq = z.wrapsq = z.m10 = z.m01 = z.m20 = z.m02 = z.top2p5_sum = z.top5_sum = 0;
      z.vmin = z.vmax = z.med_lo = z.med_hi = 0;
      *dst = z;
    }
    return;
  }
  if (tid == 0) s_top = 0;
  const uint16_t* lab = labels + (i64)p * lab_plane_stride;
  const PX* px = pixels + tile_offset[plane_tile[p]] + (i64)rq.channel * chan_stride;
  const int n_px = H * W;
  // ---- the background values, once ----
  u32 lo = 0xFFFFFFFFu, hi = 0;
  const u32 inv_w = 0xFFFFFFFFu / (u32)W + 1u;  // row of pixel i = umulhi(i, inv_w): exact for i < 2^32 / W (n_px <= 16 384)
  for (int base = 0; base < n_px; base += kThreads * 8) {
    int idx[8];
    bool bg[8];
    const PX* src[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      idx[u] = base + u * kThreads + tid;
      const int i = idx[u] < n_px ? idx[u] : 0;
      const int r = (int)__umulhi((u32)i, inv_w), c = i - r * W;
      bg[u] = idx[u] < n_px && __ldg(lab + (i64)r * lab_row_stride + c) == 0;
      src[u] = px + (i64)r * px_row_stride + c;
    }
    u32 x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = bg[u] ? (u32)__ldg(src[u]) : 0u;
    for (int z = 1; z < Z; ++z)  // Z-max (Z-add of a single plane is the plane itself)
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (bg[u]) x[u] = max(x[u], (u32)__ldg(src[u] + (i64)z * z_stride));
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (idx[u] < n_px) {
        vals[idx[u]] = (unsigned short)x[u];
        isbg[idx[u]] = bg[u] ? 1 : 0;
        if (bg[u]) { lo = min(lo, x[u]); hi = max(hi, x[u]); }
      }
  }
  lo = __reduce_min_sync(0xFFFFFFFFu, lo);
  hi = __reduce_max_sync(0xFFFFFFFFu, hi);
  if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; }
  __syncthreads();
  for (int k = 0; k < kThreads / 32; ++k) { lo = min(lo, s_lo[k]); hi = max(hi, s_hi[k]); }
  const u32 ranks[2] = {(n - 1) / 2, n / 2};
  const u32 top_from = n - (n < 5u ? n : 5u);  // elements with rank >= top_from are the (up to) five largest
  // The histogram covers kChunkValues values at a time (16 KB of packed counters); wider ranges —
  // rare for a background — take several rounds over the parked values, in value order, with a running rank.
  u32 acc0 = 0;  // background pixels below the current chunk (block-uniform)
  for (u32 cmin = lo & ~1u; cmin <= hi; cmin += (u32)kChunkValues) {
    const u32 cmax = min(hi, cmin + (u32)kChunkValues - 1u);
    const int words = (int)((cmax - cmin) >> 1) + 1;  // <= kChunkValues / 2
    const int words4 = (words + 3) & ~3;
    __syncthreads();
    for (int k = tid; k < words4 / 4; k += kThreads) reinterpret_cast<uint4*>(h32)[k] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < n_px; i += kThreads)
      if (isbg[i]) {
        const u32 d = (u32)vals[i] - cmin;  // wraps for values below the chunk
        if (d < (u32)kChunkValues) atomicAdd(&h32[d >> 1], 1u << ((d & 1u) << 4));
      }
    __syncthreads();
    // contiguous chunks of words per thread, prefix over their counts, then the few owners walk theirs in value order
    const int per = (words + kThreads - 1) / kThreads;
    const int w0 = tid * per;
    u32 cnt = 0;
    for (int k = 0; k < per; ++k) {
      int j = k + (tid & 31);  // rotated start: the lanes of a warp hit different banks
      if (j >= per) j -= per * (j / per);
      const int w = w0 + j;
      if (w < words) { const u32 v = h32[w]; cnt += (v & 0xFFFFu) + (v >> 16); }
    }
    // exclusive prefix over the 256 partial counts: warp scans, then the warps' totals
    u32 incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) part[tid >> 5] = incl;
    __syncthreads();
    u32 below = acc0 + incl - cnt, total = acc0;
    for (int k = 0; k < kThreads / 32; ++k) {
      const u32 t = part[k];
      if (k < (tid >> 5)) below += t;
      total += t;
    }
    if (tid == 0) s_total = total;
    if (cnt > 0 && ((ranks[0] >= below && ranks[0] < below + cnt) || (ranks[1] >= below && ranks[1] < below + cnt) ||
                    below + cnt > top_from)) {
      u32 acc = below;
      u64 top = 0;
      for (int k = 0; k < 2 * per && w0 + (k >> 1) < words; ++k) {
        const u32 word = h32[w0 + (k >> 1)];
        const u32 c = (k & 1) ? (word >> 16) : (word & 0xFFFFu);
        const u32 value = cmin + 2u * (u32)w0 + (u32)k;
        if (c) {
          for (int j = 0; j < 2; ++j)
            if (ranks[j] >= acc && ranks[j] < acc + c) s_med[j] = value;
          if (acc + c > top_from) {
            const u32 first = acc > top_from ? acc : top_from;
            top += (u64)(acc + c - first) * (u64)value;
          }
        }
        acc += c;
      }
      if (top) atomicAdd(reinterpret_cast<unsigned long long*>(&s_top), (unsigned long long)top);
    }
    __syncthreads();
    acc0 = s_total;
  }
  __syncthreads();
  if (tid == 0) {
    ChanStats z;
    z.sum = z.sumsq = z.wrapsq = z.m10 = z.m01 = z.m20 = z.m02 = z.top2p5_sum = 0;
    z.vmin = z.vmax = 0;
    z.med_lo = s_med[0]; z.med_hi = s_med[1];
    z.top5_sum = s_top;
    *dst = z;
  }
}

}  // namespace

// background.cu takes the per-plane background of integer pixels: streaming histogram for large planes, shared-memory
// histogram per (plane, request) for small ones (the caller's CTA kernel then only keeps the wide Z-add requests)
bool abx_big_background(const abx_extract_args* a) {
  return a->with_background && a->n_requests > 0 && (a->pixel_dtype == ABX_U8 || a->pixel_dtype == ABX_U16);
}

static bool streaming_background(const abx_extract_args* a) {
  return abx_big_background(a) && (i64)a->H * a->W > kBigBackground;
}

size_t abx_big_background_bytes(const abx_extract_args* a) {
  return streaming_background(a) ? (size_t)a->n_planes * (size_t)a->n_requests * kValues * sizeof(u32) : 0;
}

template <typename PX>
static int launch_tile_background(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  constexpr size_t smem_max = (size_t)kChunkValues * 2 + (size_t)kTilePixels * 3;  // 16-bit counters | u16 values | flags
  const int px_slots = (a->H * a->W + 15) & ~15;
  const size_t smem = (size_t)kChunkValues * 2 + (size_t)px_slots * 3;
  static thread_local bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(bg_tile_kernel<PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return abx_check_cuda(e, "bg_tile smem attribute");
    done[dev] = true;
  }
  if (a->n_requests > 65535) return abx_set_error(ABX_ERR_INVALID, "more than 65535 requests with a background metric");
  bg_tile_kernel<PX><<<dim3(a->n_planes, a->n_requests), kThreads, smem, st>>>(
      static_cast<const uint16_t*>(a->labels), a->label_plane_stride, a->label_row_stride, a->plane_tile, a->H, a->W,
      static_cast<const PX*>(a->pixels), reinterpret_cast<const i64*>(a->tile_offset), a->chan_stride, a->z_stride,
      a->row_stride, a->Z, a->requests, a->n_requests, a->n_objects, ws.recs, ws.chan, px_slots);
  return abx_check_cuda(cudaGetLastError(), "bg_tile");
}

int launch_big_background(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  if (!abx_big_background(a) || a->n_planes == 0) return ABX_OK;
  if (!streaming_background(a))
    return a->pixel_dtype == ABX_U16 ? launch_tile_background<uint16_t>(a, ws, st) : launch_tile_background<uint8_t>(a, ws, st);
  cudaError_t e = cudaMemsetAsync(ws.bg_hist, 0, abx_big_background_bytes(a), st);
  if (e != cudaSuccess) return abx_check_cuda(e, "big_background memset");
  const i64 n = (i64)a->H * a->W;
  int bx = (int)((n + kThreads * 8 - 1) / (kThreads * 8));
  const int cap = (148 * 8 + a->n_planes - 1) / a->n_planes;
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  dim3 grid(bx, a->n_planes);
#define ABX_LAUNCH_BG(PX)                                                                                            \
  bg_hist_kernel<PX><<<grid, kThreads, 0, st>>>(                                                                     \
      static_cast<const uint16_t*>(a->labels), a->label_plane_stride, a->label_row_stride, a->plane_tile, a->H, a->W, \
      static_cast<const PX*>(a->pixels), reinterpret_cast<const i64*>(a->tile_offset), a->chan_stride, a->z_stride,  \
      a->row_stride, a->Z, a->requests, a->n_requests, ws.bg_hist)
  if (a->pixel_dtype == ABX_U16) ABX_LAUNCH_BG(uint16_t);
  else ABX_LAUNCH_BG(uint8_t);
#undef ABX_LAUNCH_BG
  bg_select_kernel<<<dim3(a->n_requests, a->n_planes), kThreads, 0, st>>>(ws.bg_hist, a->requests, a->n_requests, a->Z,
                                                                         a->n_objects, ws.recs, ws.chan);
  return abx_check_cuda(cudaGetLastError(), "big_background");
}
