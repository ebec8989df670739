// Shape metrics from three chained exact Euclidean distance transforms, one CTA per object.
//
// Restates cell.py:207-229 (min_maj_approximation) and cell.py:176-187 (conical_volume):
//   nn       = EDT(mask)            distance of every object pixel to the nearest non-object pixel
//   dn       = EDT(nn != max nn)    distance to the nearest "cone top" pixel (where nn is maximal)
//   cone_top = EDT(dn == 0)         for cone-top pixels: distance to the nearest other object pixel
// The reference runs each EDT on the whole padded plane (0.9 s per object at 2160^2); here the
// transforms run on the bounding box plus the one-pixel frame the reference pads with, which is
// exact because the nearest zero of an object pixel never lies outside that window.  Squared
// distances are integers (row pass + column pass, separable and exact); the square roots are
// taken in fp64 exactly as SciPy does.  The one plane-dependent case — an object made only of
// cone-top pixels, for which SciPy's transform of an input without zeros measures the distance
// to index (-1, 0) of the padded plane — is reproduced explicitly.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr u32 kNone = 0xFFFFu;  // "no zero in this row"

struct Buf {
  u32* d2;
  unsigned short* g;
  unsigned char* mask;  // 1 = object
  unsigned char* top;   // 1 = cone top
};

__device__ __forceinline__ Buf carve(unsigned char* base, size_t A) {
  Buf b;
  b.d2 = reinterpret_cast<u32*>(base);
  b.g = reinterpret_cast<unsigned short*>(base + 4 * A);
  b.mask = base + 6 * A;
  b.top = base + 7 * A;
  return b;
}

// g[i][j] = distance along row i to the nearest pixel with zero(idx) (kNone if the row has none)
template <class Zero>
__device__ __forceinline__ void row_pass(const Buf& b, int hp, int wp, Zero zero) {
  for (int i = threadIdx.x; i < hp; i += kThreads) {
    const size_t o = (size_t)i * wp;
    u32 d = kNone;
    for (int j = 0; j < wp; ++j) {
      d = zero(o + j) ? 0u : min(d + 1u, kNone);
      b.g[o + j] = (unsigned short)d;
    }
    d = kNone;
    for (int j = wp - 1; j >= 0; --j) {
      d = zero(o + j) ? 0u : min(d + 1u, kNone);
      if (d < b.g[o + j]) b.g[o + j] = (unsigned short)d;
    }
  }
}

// d2[idx] = min over rows i' of g[i'][j]^2 + (i - i')^2 for pixels with want(idx)
template <class Want>
__device__ __forceinline__ void col_pass(const Buf& b, int hp, int wp, Want want) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < hp; i += kWarps) {
    for (int j = lane; j < wp; j += 32) {
      const size_t idx = (size_t)i * wp + j;
      if (!want(idx)) continue;
      const u32 g0 = b.g[idx];
      u32 best = (g0 == kNone) ? 0xFFFFFFFFu : g0 * g0;
      for (u32 dr = 1; dr * dr < best; ++dr) {
        const bool up = (int)dr <= i, down = i + (int)dr < hp;
        if (!up && !down) break;
        if (up) {
          const u32 gg = b.g[idx - (size_t)dr * wp];
          if (gg != kNone) best = min(best, gg * gg + dr * dr);
        }
        if (down) {
          const u32 gg = b.g[idx + (size_t)dr * wp];
          if (gg != kNone) best = min(best, gg * gg + dr * dr);
        }
      }
      b.d2[idx] = best;
    }
  }
}

struct Red {
  double dsum[kWarps];
  u32 umax[kWarps];
  u32 ucnt[kWarps];
};

__device__ __forceinline__ void block_reduce(Red& r, double& dsum, u32& umax, u32& ucnt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xFFFFFFFFu, dsum, o);
  umax = __reduce_max_sync(0xFFFFFFFFu, umax);
  ucnt = __reduce_add_sync(0xFFFFFFFFu, ucnt);
  __syncthreads();
  if (lane == 0) { r.dsum[warp] = dsum; r.umax[warp] = umax; r.ucnt[warp] = ucnt; }
  __syncthreads();
  dsum = 0; umax = 0; ucnt = 0;
  for (int w = 0; w < kWarps; ++w) { dsum += r.dsum[w]; umax = max(umax, r.umax[w]); ucnt += r.ucnt[w]; }
}

__global__ void __launch_bounds__(kThreads)
shape_edt_kernel(const uint16_t* __restrict__ labels, i64 plane_stride, i64 row_stride,
                 const int32_t* __restrict__ plane_base, int n_planes, int n_objects,
                 const abx_object_rec* __restrict__ recs, ShapeStats* __restrict__ out, int large_mode,
                 unsigned char* scratch, size_t scratch_per_cta, const int* __restrict__ work_list,
                 const u32* __restrict__ work_count) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ Red red;
  __shared__ abx_object_rec rec;
  __shared__ int s_plane;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const u32 n_work = *work_count;  // objects handed over by the warp-per-object kernel
  for (u32 wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const int obj = work_list[wi];
    __syncthreads();
    if (threadIdx.x == 0) { rec = recs[obj]; s_plane = find_plane(plane_base, n_planes, obj); }
    __syncthreads();
    const u32 n = rec.n;
    if (n == 0) {
      if (!large_mode && threadIdx.x == 0) { ShapeStats z; z.sum_nn = 0; z.sum_top = 0; z.max_nn2 = 0; z.max_dn2 = 0; out[obj] = z; }
      continue;
    }
    const int hp = (int)(rec.rmax - rec.rmin) + 3, wp = (int)(rec.cmax - rec.cmin) + 3;
    const size_t A = (size_t)hp * wp;
    const bool is_large = A > (size_t)kEdtSmemWindow;
    if (is_large != (large_mode != 0)) continue;
    const Buf b = carve(is_large ? scratch + (size_t)blockIdx.x * scratch_per_cta : dyn, A);
    const u32 label = (u32)(obj - plane_base[s_plane] + 1);
    const uint16_t* lab = labels + (i64)s_plane * plane_stride;

    // mask of the framed window
    for (int i = warp; i < hp; i += kWarps) {
      const bool row_in = i >= 1 && i <= hp - 2;
      const uint16_t* lrow = lab + (i64)(rec.rmin + i - 1) * row_stride + ((i64)rec.cmin - 1);
      for (int j = lane; j < wp; j += 32) {
        const bool in = row_in && j >= 1 && j <= wp - 2;
        b.mask[(size_t)i * wp + j] = (in && (u32)__ldg(lrow + j) == label) ? 1 : 0;
      }
    }
    __syncthreads();

    // ---- EDT 1: distance to the background ----
    row_pass(b, hp, wp, [&](size_t k) { return b.mask[k] == 0; });
    __syncthreads();
    col_pass(b, hp, wp, [&](size_t k) { return b.mask[k] != 0; });
    __syncthreads();
    double s_nn = 0; u32 m_nn2 = 0, dummy = 0;
    for (size_t k = threadIdx.x; k < A; k += kThreads)
      if (b.mask[k]) { const u32 d = b.d2[k]; m_nn2 = max(m_nn2, d); s_nn += sqrt((double)d); }
    block_reduce(red, s_nn, m_nn2, dummy);
    u32 n_top = 0; double dz = 0; u32 uz = 0;
    for (size_t k = threadIdx.x; k < A; k += kThreads) {
      const unsigned char t = (b.mask[k] && b.d2[k] == m_nn2) ? 1 : 0;
      b.top[k] = t;
      n_top += t;
    }
    block_reduce(red, dz, uz, n_top);  // also orders top[] writes before the next pass

    // ---- EDT 2: distance to the cone top ----
    row_pass(b, hp, wp, [&](size_t k) { return b.top[k] != 0; });
    __syncthreads();
    col_pass(b, hp, wp, [&](size_t k) { return b.mask[k] != 0; });
    __syncthreads();
    u32 m_dn2 = 0; dz = 0; uz = 0;
    for (size_t k = threadIdx.x; k < A; k += kThreads)
      if (b.mask[k]) m_dn2 = max(m_dn2, b.d2[k]);
    block_reduce(red, dz, m_dn2, uz);

    // ---- EDT 3: plateau size = distance from cone-top pixels to the rest of the object ----
    double s_top = 0;
    if (n_top == n) {
      // no zero anywhere in `dn == 0`: SciPy measures to index (-1, 0) of the padded plane
      for (int i = warp; i < hp; i += kWarps)
        for (int j = lane; j < wp; j += 32)
          if (b.mask[(size_t)i * wp + j]) {
            const double dr = (double)rec.rmin + (double)i + 1.0, dc = (double)rec.cmin + (double)j;
            s_top += sqrt(dr * dr + dc * dc);
          }
    } else {
      row_pass(b, hp, wp, [&](size_t k) { return b.mask[k] != 0 && b.top[k] == 0; });
      __syncthreads();
      col_pass(b, hp, wp, [&](size_t k) { return b.top[k] != 0; });
      __syncthreads();
      for (size_t k = threadIdx.x; k < A; k += kThreads)
        if (b.top[k]) s_top += sqrt((double)b.d2[k]);
    }
    uz = 0; u32 uz2 = 0;
    block_reduce(red, s_top, uz, uz2);
    if (threadIdx.x == 0) {
      ShapeStats o;
      o.sum_nn = s_nn; o.sum_top = s_top; o.max_nn2 = m_nn2; o.max_dn2 = m_dn2;
      out[obj] = o;
    }
  }
}

}  // namespace

int launch_shape_edt(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  if (!(a->need_edt & 3) || a->n_objects == 0) return ABX_OK;
  const size_t smem = (size_t)kEdtSmemWindow * kEdtBytesPerPixel;
  static thread_local bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(shape_edt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return abx_check_cuda(e, "shape_edt smem attribute");
    attr_done[dev] = true;
  }
  const int grid = a->n_objects < 148 * 3 ? a->n_objects : 148 * 3;
  shape_edt_kernel<<<grid, kThreads, smem, st>>>(static_cast<const uint16_t*>(a->labels), a->label_plane_stride,
                                                 a->label_row_stride, a->plane_base, a->n_planes, a->n_objects,
                                                 ws.recs, ws.shape, 0, nullptr, 0, ws.edt_list, ws.list_counts + 1);
  if (ws.edt_scratch_per_cta) {
    shape_edt_kernel<<<kEdtLargeCtas, kThreads, 0, st>>>(static_cast<const uint16_t*>(a->labels),
                                                         a->label_plane_stride, a->label_row_stride, a->plane_base,
                                                         a->n_planes, a->n_objects, ws.recs, ws.shape, 1,
                                                         ws.edt_scratch, ws.edt_scratch_per_cta, ws.edt_list,
                                                         ws.list_counts + 1);
  }
  return abx_check_cuda(cudaGetLastError(), "shape_edt");
}
