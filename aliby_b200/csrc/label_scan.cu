// Label scan: one streaming pass over the uint16 label planes that produces, per
// (plane, label), the area, the bounding box and the 1-based coordinate sums.
//
// Replaces the one-hot expansion of src/agora/utils/masks.py:35-37 (L*Y*X bytes) and
// the full-plane products of cell.py:18-27 (area) and cell.py:282-303 (centroid).
//
// Mapping: a warp owns a 256-pixel row chunk per iteration, each lane 8 consecutive
// labels (one 128-bit load when the row pitch allows).  Lanes whose 8 labels agree are
// merged with their neighbours through two ballots, so one lane per label run issues the
// atomics: no shuffles, no shared memory.  Bounding-box atomics are skipped when a plain
// (possibly stale, hence conservative) read shows they cannot change the record.
#include "common.cuh"

namespace {

constexpr int kScanThreads = 256;

__global__ void init_records_kernel(abx_object_rec* recs, int n_objects, int n_planes, int H, int W, u32* err) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *err = 0;
  if (i < n_objects) {
    abx_object_rec r;
    r.sum_row = 0; r.sum_col = 0; r.n = 0;
    r.rmin = 0xFFFFFFFFu; r.rmax = 0; r.cmin = 0xFFFFFFFFu; r.cmax = 0; r.pad_ = 0;
    recs[i] = r;
  } else if (i < n_objects + n_planes) {
    abx_object_rec r;  // background object of a plane: label 0, bbox = whole plane
    r.sum_row = 0; r.sum_col = 0; r.n = 0;
    r.rmin = 0; r.rmax = (u32)(H - 1); r.cmin = 0; r.cmax = (u32)(W - 1); r.pad_ = 0;
    recs[i] = r;
  }
}

__device__ __forceinline__ void emit_run(abx_object_rec* __restrict__ recs, const int32_t* __restrict__ plane_base,
                                         int p, u32 label, u32 row, u32 cs, u32 ce, u32* err) {
  const int base = plane_base[p];
  const u32 n_labels = (u32)(plane_base[p + 1] - base);
  if (label > n_labels) { atomicOr(err, 1u); return; }
  abx_object_rec* rec = recs + base + (label - 1);
  const u32 count = ce - cs + 1;
  atomicAdd(&rec->n, count);
  atomicAdd(reinterpret_cast<u64*>(&rec->sum_row), (u64)count * (u64)(row + 1));
  atomicAdd(reinterpret_cast<u64*>(&rec->sum_col), ((u64)(cs + 1) + (u64)(ce + 1)) * (u64)count / 2);
  // bbox: the fields are monotone, so a stale read can only cause a redundant atomic
  if (row < __ldcg(&rec->rmin)) atomicMin(&rec->rmin, row);
  if (row > __ldcg(&rec->rmax)) atomicMax(&rec->rmax, row);
  if (cs < __ldcg(&rec->cmin)) atomicMin(&rec->cmin, cs);
  if (ce > __ldcg(&rec->cmax)) atomicMax(&rec->cmax, ce);
}

__global__ void __launch_bounds__(kScanThreads)
label_scan_kernel(const uint16_t* __restrict__ labels, int n_planes, int H, int W, i64 plane_stride, i64 row_stride,
                  const int32_t* __restrict__ plane_base, abx_object_rec* __restrict__ recs, int n_objects,
                  int with_bg, int vec_ok, u32* err) {
  const u32 lane = lane_id();
  const int warps_per_block = blockDim.x >> 5;
  const i64 gwarp = (i64)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const i64 nwarps = (i64)gridDim.x * warps_per_block;
  const int chunks_per_row = (W + 255) >> 8;
  const i64 total = (i64)n_planes * H * chunks_per_row;
  u32 bg_count = 0;
  int bg_plane = -1;

  for (i64 chunk = gwarp; chunk < total; chunk += nwarps) {
    const int cx = (int)(chunk % chunks_per_row);
    const i64 t = chunk / chunks_per_row;
    const u32 row = (u32)(t % H);
    const int p = (int)(t / H);
    if (with_bg && p != bg_plane) {  // warp-uniform: flush the background count of the previous plane
      const u32 tot = __reduce_add_sync(0xFFFFFFFFu, bg_count);
      if (lane == 0 && bg_plane >= 0 && tot) atomicAdd(&recs[n_objects + bg_plane].n, tot);
      bg_count = 0;
      bg_plane = p;
    }
    const u32 c0 = (u32)cx * 256u + lane * 8u;
    const int cnt = (int)min((i64)8, max((i64)0, (i64)W - (i64)c0));
    const uint16_t* src = labels + (i64)p * plane_stride + (i64)row * row_stride + c0;
    u32 l[8];
    if (vec_ok && cnt == 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
      l[0] = v.x & 0xFFFFu; l[1] = v.x >> 16; l[2] = v.y & 0xFFFFu; l[3] = v.y >> 16;
      l[4] = v.z & 0xFFFFu; l[5] = v.z >> 16; l[6] = v.w & 0xFFFFu; l[7] = v.w >> 16;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) l[i] = (i < cnt) ? (u32)__ldg(src + i) : 0xFFFFFFFFu;
    }
    bool uniform = cnt > 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) uniform = uniform && (i >= cnt || l[i] == l[0]);
    const u32 L = l[0];
    const u32 unif = __ballot_sync(0xFFFFFFFFu, uniform);
    const u32 prevL = __shfl_up_sync(0xFFFFFFFFu, L, 1);
    const bool head = uniform && (lane == 0 || !((unif >> (lane - 1)) & 1u) || prevL != L);
    const u32 heads = __ballot_sync(0xFFFFFFFFu, head);
    if (head) {
      const u32 brk = heads | ~unif;
      const u32 above = (lane == 31) ? 0u : (brk & ~((2u << lane) - 1u));
      const int nxt = above ? (__ffs(above) - 1) : 32;
      const u32 c_tail = (u32)cx * 256u + (u32)(nxt - 1) * 8u;
      const u32 ce = min(c_tail + 8u, (u32)W) - 1u;
      if (L == 0) { if (with_bg) bg_count += ce - c0 + 1; }
      else emit_run(recs, plane_base, p, L, row, c0, ce, err);
    } else if (!uniform && cnt > 0) {
      // mixed lane: walk its (up to 8) runs
      int s = 0;
      for (int i = 1; i <= cnt; ++i) {
        if (i == cnt || l[i] != l[s]) {
          if (l[s] == 0) { if (with_bg) bg_count += (u32)(i - s); }
          else emit_run(recs, plane_base, p, l[s], row, c0 + s, c0 + i - 1, err);
          s = i;
        }
      }
    }
  }
  if (with_bg) {
    const u32 tot = __reduce_add_sync(0xFFFFFFFFu, bg_count);
    if (lane == 0 && bg_plane >= 0 && tot) atomicAdd(&recs[n_objects + bg_plane].n, tot);
  }
}

__global__ void label_max_kernel(const uint16_t* __restrict__ labels, int n_planes, int H, int W, i64 plane_stride,
                                 i64 row_stride, int32_t* out_max) {
  // grid: (blocks_per_plane, n_planes); out_max zeroed by the caller
  const int p = blockIdx.y;
  const i64 n = (i64)H * W;
  u32 m = 0;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const i64 r = i / W, c = i - r * W;
    m = max(m, (u32)__ldg(labels + (i64)p * plane_stride + r * row_stride + c));
  }
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  if (lane_id() == 0 && m) atomicMax(out_max + p, (int32_t)m);
}

}  // namespace

int launch_label_scan(const abx_extract_args* a, abx_object_rec* recs, u32* err, cudaStream_t st) {
  const int n_rec = a->n_objects + a->n_planes;
  init_records_kernel<<<(n_rec + 255) / 256, 256, 0, st>>>(recs, a->n_objects, a->n_planes, a->H, a->W, err);
  const i64 chunks = (i64)a->n_planes * a->H * ((a->W + 255) / 256);
  if (chunks == 0) return abx_check_cuda(cudaGetLastError(), "init_records");
  const int warps_per_block = kScanThreads / 32;
  i64 blocks = (chunks + warps_per_block - 1) / warps_per_block;
  const i64 cap = 148 * 8 * 4;  // a few waves of 8 resident CTAs per SM
  if (blocks > cap) blocks = cap;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(a->labels) & 15u) == 0) && (a->label_row_stride % 8 == 0) &&
                     (a->label_plane_stride % 8 == 0);
  label_scan_kernel<<<(int)blocks, kScanThreads, 0, st>>>(
      static_cast<const uint16_t*>(a->labels), a->n_planes, a->H, a->W, a->label_plane_stride, a->label_row_stride,
      a->plane_base, recs, a->n_objects, a->with_background, vec_ok, err);
  return abx_check_cuda(cudaGetLastError(), "label_scan");
}

extern "C" int abx_label_max(const void* labels, int32_t label_dtype, int32_t n_planes, int32_t H, int32_t W,
                             int64_t plane_stride, int64_t row_stride, int32_t* out_max, void* stream) {
  if (label_dtype != ABX_U16) return abx_set_error(ABX_ERR_UNSUPPORTED, "abx_label_max: labels must be uint16");
  if (!labels || !out_max || n_planes < 0 || H <= 0 || W <= 0)
    return abx_set_error(ABX_ERR_INVALID, "abx_label_max: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_planes == 0) return ABX_OK;
  cudaError_t e = cudaMemsetAsync(out_max, 0, sizeof(int32_t) * n_planes, st);
  if (e != cudaSuccess) return abx_check_cuda(e, "label_max memset");
  const i64 n = (i64)H * W;
  int bx = (int)min((i64)64, (n + 1023) / 1024);
  dim3 grid(bx < 1 ? 1 : bx, n_planes);
  label_max_kernel<<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(labels), n_planes, H, W, plane_stride,
                                         row_stride, out_max);
  return abx_check_cuda(cudaGetLastError(), "label_max");
}
