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
// Planes of at most kBigBackground pixels (trap tiles) and wide values (Z-add of more than one plane) stay with
// object_stats.cu.
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

}  // namespace

bool abx_big_background(const abx_extract_args* a) {
  return a->with_background && a->n_requests > 0 && (i64)a->H * a->W > kBigBackground &&
         (a->pixel_dtype == ABX_U8 || a->pixel_dtype == ABX_U16);
}

size_t abx_big_background_bytes(const abx_extract_args* a) {
  return abx_big_background(a) ? (size_t)a->n_planes * (size_t)a->n_requests * kValues * sizeof(u32) : 0;
}

int launch_big_background(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  if (!abx_big_background(a) || a->n_planes == 0) return ABX_OK;
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
