// Label scan: one streaming pass over the uint16 label planes that produces, per
// (plane, label), the area, the bounding box and the 1-based coordinate sums.
//
// Replaces the one-hot expansion of src/agora/utils/masks.py:35-37 (L*Y*X bytes) and
// the full-plane products of cell.py:18-27 (area) and cell.py:282-303 (centroid).
//
// Mapping: a warp owns a 256-pixel row chunk per iteration.  The chunk is staged in shared
// memory with one 128-bit load per lane, then read back with a stride-32 mapping so that the
// eight ballots of "this pixel starts a run" form a 256-bit mask in pixel order.  The lane
// that owns a run start finds the run end with bit scans and issues the atomics: one lane
// per label run, no shuffles.  Bounding-box atomics are skipped when one 128-bit read of the
// record (possibly stale, hence conservative) shows they cannot change it.
#include "common.cuh"

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;

__global__ void init_records_kernel(abx_object_rec* recs, int n_objects, int n_planes, int H, int W, u32* err) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 8) err[i] = 0;  // error flags, the two work-list lengths, the two object counters
  if (i < n_objects + n_planes) {
    abx_object_rec r;
    r.sum_row = 0; r.sum_col = 0; r.n = 0;
    r.pad_[0] = r.pad_[1] = r.pad_[2] = 0;
    if (i < n_objects) {
      r.rmin = 0xFFFFFFFFu; r.rmax = 0; r.cmin = 0xFFFFFFFFu; r.cmax = 0;
    } else {  // background object of a plane: label 0, bbox = whole plane
      r.rmin = 0; r.rmax = (u32)(H - 1); r.cmin = 0; r.cmax = (u32)(W - 1);
    }
    recs[i] = r;
  }
}

__device__ __forceinline__ void emit_run(abx_object_rec* __restrict__ recs, int base, u32 n_labels, u32 label,
                                         u32 row, u32 cs, u32 ce, u32* err) {
  if (label > n_labels) { atomicOr(err, 1u); return; }
  abx_object_rec* rec = recs + base + (label - 1);
  const u32 count = ce - cs + 1;
  atomicAdd(&rec->n, count);
  atomicAdd(reinterpret_cast<u64*>(&rec->sum_row), (u64)count * (u64)(row + 1));
  atomicAdd(reinterpret_cast<u64*>(&rec->sum_col), ((u64)(cs + 1) + (u64)(ce + 1)) * (u64)count / 2);
  // bbox: the fields are monotone, so a stale read can only cause a redundant atomic
  const uint4 bb = __ldcg(reinterpret_cast<const uint4*>(&rec->rmin));  // rmin, rmax, cmin, cmax
  if (row < bb.x) atomicMin(&rec->rmin, row);
  if (row > bb.y) atomicMax(&rec->rmax, row);
  if (cs < bb.z) atomicMin(&rec->cmin, cs);
  if (ce > bb.w) atomicMax(&rec->cmax, ce);
}

__global__ void __launch_bounds__(kScanThreads)
label_scan_kernel(const uint16_t* __restrict__ labels, int n_planes, int H, int W, i64 plane_stride, i64 row_stride,
                  const int32_t* __restrict__ plane_base, abx_object_rec* __restrict__ recs, int n_objects,
                  int with_bg, int vec_ok, u32* err) {
  __shared__ __align__(16) uint16_t stage_all[kScanWarps][256];
  uint16_t* stage = stage_all[threadIdx.x >> 5];
  const u32 lane = lane_id();
  const i64 gwarp = (i64)blockIdx.x * kScanWarps + (threadIdx.x >> 5);
  const i64 nwarps = (i64)gridDim.x * kScanWarps;
  const int chunks_per_row = (W + 255) >> 8;
  const i64 total = (i64)n_planes * H * chunks_per_row;
  u32 bg_count = 0;  // warp-uniform
  int bg_plane = -1;

  for (i64 chunk = gwarp; chunk < total; chunk += nwarps) {
    const int cx = (int)(chunk % chunks_per_row);
    const i64 t = chunk / chunks_per_row;
    const u32 row = (u32)(t % H);
    const int p = (int)(t / H);
    if (with_bg && p != bg_plane) {  // flush the background count of the previous plane
      if (lane == 0 && bg_plane >= 0 && bg_count) atomicAdd(&recs[n_objects + bg_plane].n, bg_count);
      bg_count = 0;
      bg_plane = p;
    }
    const u32 cbase = (u32)cx * 256u;
    const int len = min(256, W - (int)cbase);  // valid pixels of this chunk
    const uint16_t* src = labels + (i64)p * plane_stride + (i64)row * row_stride + cbase;
    __syncwarp();
    {
      const int o = (int)lane * 8;
      if (vec_ok && o + 8 <= len) {
        *reinterpret_cast<uint4*>(stage + o) = __ldg(reinterpret_cast<const uint4*>(src + o));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (o + i < len) stage[o + i] = __ldg(src + o + i);
      }
    }
    __syncwarp();
    // run starts in pixel order: bit `lane` of sflag[k] <-> pixel 32 k + lane
    u32 sflag[8], mine[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int q = 32 * k + (int)lane;
      const bool valid = q < len;
      const u32 cur = valid ? (u32)stage[q] : 0u;
      const u32 prev = (valid && q > 0) ? (u32)stage[q - 1] : 0xFFFFFFFFu;
      mine[k] = cur;
      sflag[k] = __ballot_sync(0xFFFFFFFFu, valid && cur != prev);
      if (with_bg) bg_count += __popc(__ballot_sync(0xFFFFFFFFu, valid && cur == 0u));
    }
    const int base = plane_base[p];
    const u32 n_labels = (u32)(plane_base[p + 1] - base);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (!((sflag[k] >> lane) & 1u) || mine[k] == 0u) continue;
      // run end = next start - 1 (or the end of the chunk)
      u32 nxt = (lane == 31) ? 0u : (sflag[k] & ~((2u << lane) - 1u));
      int endq = -1;
      if (nxt) endq = 32 * k + __ffs(nxt) - 1;
#pragma unroll
      for (int kk = 1; kk < 8; ++kk)
        if (k + kk < 8 && endq < 0 && sflag[(k + kk) & 7]) endq = 32 * (k + kk) + __ffs(sflag[(k + kk) & 7]) - 1;
      if (endq < 0 || endq > len) endq = len;
      emit_run(recs, base, n_labels, mine[k], row, cbase + 32u * k + lane, cbase + (u32)endq - 1u, err);
    }
  }
  if (with_bg && lane == 0 && bg_plane >= 0 && bg_count) atomicAdd(&recs[n_objects + bg_plane].n, bg_count);
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
  const int warps_per_block = kScanWarps;
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
