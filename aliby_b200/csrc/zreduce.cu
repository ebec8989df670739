// Z stacks: a streaming Z-max pass in front of the per-object kernels.
//
// reduce_z (src/extraction/core/functions/distributors.py:6-24) is a ufunc.reduce over axis 0 of the (Z, Y, X) stack of
// one channel; the reference recomputes it for every (object, metric) call (extract.py:105-106).  The per-object
// kernels can fuse it into their gathers (object_warp.cu), but then every VALUE costs Z dependent loads: a C4 field
// (5 channels x 16 z x 2048^2) took 3.9 ms, 2.7 % of HBM.  Here each requested (tile, channel) stack is reduced ONCE,
// streaming (128-bit loads, HBM-bound), into a dense plane of the workspace — plane (tile, q) for request q — and the
// Z = 1 kernels then run on those planes: `max` planes keep the pixel dtype and go to the TMA-staged window kernel,
// `add` planes are uint32 sums (the 16-bit window kernel does not take them) and go to the gather kernel, which now
// reads one value per pixel instead of Z.  `div` requests (floating point) stay with object_float.cu.
//
// Every plane slot holds H x W x 4 bytes.  Two rewritten request lists make the split, both with channel = q (the plane
// index): req_tma keeps the `max` requests, req_rest the `add` requests; the others are marked as skipped with
// ABX_RED_DIV, which both kernels already leave to object_float.cu (that kernel reads the caller's own list).
#include "common.cuh"

namespace {

template <typename PX>
__global__ void __launch_bounds__(256)
zreduce_kernel(const PX* __restrict__ pixels, const i64* __restrict__ tile_offset, i64 chan_stride, i64 z_stride, i64 row_stride,
               int Z, int H, int W, const abx_request* __restrict__ requests, int n_requests, unsigned char* __restrict__ out) {
  constexpr int kVec = 16 / (int)sizeof(PX);  // pixels per 128-bit load; W is a multiple of it
  const int tile = blockIdx.y / n_requests, q = blockIdx.y - tile * n_requests;
  const abx_request rq = requests[q];
  const bool is_max = rq.reduction == ABX_RED_MAX;
  if (!is_max && rq.reduction != ABX_RED_ADD) return;
  const PX* src = pixels + tile_offset[tile] + (i64)rq.channel * chan_stride;
  unsigned char* slot = out + ((i64)tile * n_requests + q) * H * W * 4;  // 256-byte aligned: H * W * 4 is a multiple of 256
  const int wv = W / kVec;
  const i64 n_vec = (i64)H * wv;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(row_stride * sizeof(PX)) |
                         (uintptr_t)(z_stride * sizeof(PX))) & 15u) == 0;
  for (i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (i64)gridDim.x * blockDim.x) {
    const int r = (int)(v / wv), xv = (int)(v - (i64)r * wv);
    const PX* p = src + (i64)r * row_stride + (i64)xv * kVec;
    if (is_max) {
      uint4 m;
      if (aligned) {
        m = __ldg(reinterpret_cast<const uint4*>(p));
#pragma unroll 4
        for (int z = 1; z < Z; ++z) {
          const uint4 y = __ldg(reinterpret_cast<const uint4*>(p + (i64)z * z_stride));
          if (sizeof(PX) == 2) { m.x = __vmaxu2(m.x, y.x); m.y = __vmaxu2(m.y, y.y); m.z = __vmaxu2(m.z, y.z); m.w = __vmaxu2(m.w, y.w); }
          else { m.x = __vmaxu4(m.x, y.x); m.y = __vmaxu4(m.y, y.y); m.z = __vmaxu4(m.z, y.z); m.w = __vmaxu4(m.w, y.w); }
        }
      } else {  // any alignment: element by element
        PX e[kVec];
#pragma unroll
        for (int j = 0; j < kVec; ++j) e[j] = __ldg(p + j);
        for (int z = 1; z < Z; ++z)
#pragma unroll
          for (int j = 0; j < kVec; ++j) {
            const PX y = __ldg(p + (i64)z * z_stride + j);
            e[j] = y > e[j] ? y : e[j];
          }
        memcpy(&m, e, 16);
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<PX*>(slot) + (i64)r * W + (i64)xv * kVec) = m;
    } else {  // uint32 sums (np.add.reduce of uint16 gives uint64; Z * 65535 < 2^32)
      u32 acc[kVec];
#pragma unroll
      for (int j = 0; j < kVec; ++j) acc[j] = 0;
#pragma unroll 2
      for (int z = 0; z < Z; ++z) {
        PX e[kVec];
        if (aligned) {
          const uint4 y = __ldg(reinterpret_cast<const uint4*>(p + (i64)z * z_stride));
          memcpy(e, &y, 16);
        } else {
#pragma unroll
          for (int j = 0; j < kVec; ++j) e[j] = __ldg(p + (i64)z * z_stride + j);
        }
#pragma unroll
        for (int j = 0; j < kVec; ++j) acc[j] += (u32)e[j];
      }
      u32* d = reinterpret_cast<u32*>(slot) + (i64)r * W + (i64)xv * kVec;
#pragma unroll
      for (int j = 0; j < kVec; j += 4) *reinterpret_cast<uint4*>(d + j) = make_uint4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
  }
}

__global__ void zsetup_kernel(const abx_request* __restrict__ requests, int n_requests, int n_tiles, i64 plane_elems,
                              int elem_size, abx_request* __restrict__ req_tma, abx_request* __restrict__ req_rest,
                              i64* __restrict__ ztile_offset, u32* __restrict__ any_rest) {
  for (int q = threadIdx.x; q < n_requests; q += blockDim.x) {
    const abx_request rq = requests[q];
    abx_request a = rq, b = rq;
    a.channel = b.channel = q;  // the reduced plane of request q
    if (rq.reduction != ABX_RED_MAX) a.reduction = ABX_RED_DIV;  // skipped by the window kernel on the max planes
    if (rq.reduction != ABX_RED_ADD) b.reduction = ABX_RED_DIV;  // skipped by the gather pass on the uint32 sum planes
    if (rq.reduction == ABX_RED_ADD) atomicOr(any_rest, 1u);
    req_tma[q] = a;
    req_rest[q] = b;
  }
  // [0, n_tiles): tile offsets in pixel-dtype elements (4 / elem_size elements per slot pixel); [n_tiles, 2 n_tiles): in uint32
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
    ztile_offset[t] = (i64)t * n_requests * plane_elems * (4 / elem_size);
    ztile_offset[n_tiles + t] = (i64)t * n_requests * plane_elems;
  }
}

}  // namespace

// Whether abx_extract reduces the Z stacks of this call up front (host decision from the layout alone).
bool abx_zreduce_ok(const abx_extract_args* a) {
  if (a->Z <= 1 || a->n_requests <= 0 || a->n_tiles <= 0 || a->n_objects <= 0) return false;
  if (a->pixel_dtype != ABX_U16 && a->pixel_dtype != ABX_U8) return false;
  const int es = a->pixel_dtype == ABX_U8 ? 1 : 2;
  if ((i64)a->n_tiles * a->n_requests > 65535) return false;  // grid.y of the reduction; such calls keep the fused gathers
  return (a->W * es) % 16 == 0 && a->W >= 64 && a->H >= 8 && ((i64)a->H * a->W * 4) % 256 == 0;
}

size_t abx_zreduce_bytes(const abx_extract_args* a) {
  if (!abx_zreduce_ok(a)) return 0;
  const size_t es = a->pixel_dtype == ABX_U8 ? 1 : 2;
  (void)es;
  return (size_t)a->n_tiles * (size_t)a->n_requests * (size_t)a->H * (size_t)a->W * 4u;  // uint32-sized slots
}

// Fills ws.zplanes / req_tma / req_rest / ztile_offset / zflags on the stream.
int launch_zreduce(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(ws.zflags, 0, sizeof(u32), st);
  if (e != cudaSuccess) return abx_check_cuda(e, "zreduce flags");
  const i64 plane = (i64)a->H * a->W;
  const int es = a->pixel_dtype == ABX_U8 ? 1 : 2;
  zsetup_kernel<<<1, 256, 0, st>>>(a->requests, a->n_requests, a->n_tiles, plane, es, ws.req_tma, ws.req_rest,
                                   ws.ztile_offset, ws.zflags);
  const i64 n_vec = plane / (16 / es);
  i64 bx = (n_vec + 255) / 256;
  if (bx > 148 * 8) bx = 148 * 8;
  const dim3 grid((unsigned)bx, (unsigned)(a->n_tiles * a->n_requests));
  if (a->pixel_dtype == ABX_U16)
    zreduce_kernel<uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(a->pixels), reinterpret_cast<const i64*>(a->tile_offset),
                                                a->chan_stride, a->z_stride, a->row_stride, a->Z, a->H, a->W, a->requests,
                                                a->n_requests, static_cast<unsigned char*>(ws.zplanes));
  else
    zreduce_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(a->pixels), reinterpret_cast<const i64*>(a->tile_offset),
                                               a->chan_stride, a->z_stride, a->row_stride, a->Z, a->H, a->W, a->requests,
                                               a->n_requests, static_cast<unsigned char*>(ws.zplanes));
  return abx_check_cuda(cudaGetLastError(), "zreduce");
}
