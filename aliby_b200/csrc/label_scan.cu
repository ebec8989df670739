// Label scan: one streaming pass over the uint16 label planes that produces, per
// (plane, label), the area, the bounding box, the 1-based coordinate sums and a 64 x 64
// "torus" BITMAP of the object: bit (c & 63) of word (r & 63) <-> pixel (r, c).  An object whose bounding box fits
// 64 x 64 maps one-to-one into its bitmap, so the per-object kernels (object_sweep.cu, object_edt.cu) rebuild
// their row masks from 512 bytes instead of re-reading label windows: the label planes leave HBM once.
//
// Replaces the one-hot expansion of src/agora/utils/masks.py:35-37 (L*Y*X bytes) and
// the full-plane products of cell.py:18-27 (area) and cell.py:282-303 (centroid).
//
// Mapping: see label_scan_kernel — lanes walk down 8-pixel column strips with 128-bit coalesced
// loads and keep the object they are inside of in registers; one set of global atomics per
// (strip, band, object).  The first version (one warp per 256-pixel row chunk, one set of atomics
// per row run) executed 2.35 warp instructions per pixel and ran at 0.6 TB/s
// (profiles/r01e_summary.md).
#include <cstring>

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace {

#include "tma.cuh"

constexpr int kScanThreads = 128;  // 4 warps x 8 KB of staging = 32 KB static shared memory per CTA
constexpr int kScanWarps = kScanThreads / 32;

// Records, counters, and — in the same launch, so that a small call has one node less in its chain — the zeroed torus
// bitmaps (grid-stride over 16-byte words) and the sqrt table of the shape kernel.
__global__ void init_records_kernel(abx_object_rec* recs, int n_objects, int n_planes, int H, int W, u32* err,
                                    uint4* __restrict__ bitmap_words, size_t n_bitmap_words, double* __restrict__ sqrt_tab,
                                    int n_sqrt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t k = (size_t)i; k < n_bitmap_words; k += (size_t)gridDim.x * blockDim.x) bitmap_words[k] = make_uint4(0, 0, 0, 0);
  if (i < n_sqrt) sqrt_tab[i] = sqrt((double)i);
  if (i < kCounterWords) err[i] = 0;  // error flags, the work-list lengths, the work counters of the per-object kernels
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

// One object piece found by a lane: `n` pixels, coordinate sums and bounding box in plane coordinates.
struct Piece {
  u32 label, n, sum_row, sum_col;  // sums of (row + 1) and (col + 1)
  u32 rmin, rmax, cmask;           // cmask: bit j <-> column c0 + j holds a pixel of the piece
};

// The 32-bit word of the object's torus bitmap that holds (row 0, the 8-pixel strip at c0), and the strip's bit offset
// inside it: strips are 8-aligned, so the eight pixels of a strip row are exactly one byte.  The byte is OR-ed in
// (red.global.or, no return value): a cell 58 to 64 pixels wide can touch nine strips, and then its first and its
// last strip share a byte column.  Labels above n_labels (caller error, flagged by emit_piece) land in the dummy
// bitmap behind the last object.
struct BitCol {
  u32* word;  // row r lives 2 * (r & 63) words further
  u32 shift;
};
__device__ __forceinline__ BitCol bitmap_col(u64* __restrict__ bitmaps, int base, u32 n_labels, int n_objects, u32 label,
                                             u32 c0) {
  const u32 row = label <= n_labels ? (u32)base + label - 1u : (u32)n_objects;
  const u32 byte = (c0 >> 3) & 7u;
  BitCol b;
  b.word = reinterpret_cast<u32*>(bitmaps + (size_t)row * 64u) + (byte >> 2);
  b.shift = (byte & 3u) << 3;
  return b;
}
__device__ __forceinline__ void bitmap_or(const BitCol& b, u32 r, u32 m) { atomicOr(b.word + ((r & 63u) << 1), m << b.shift); }

__device__ __forceinline__ void emit_piece(abx_object_rec* __restrict__ recs, int base, u32 n_labels, const Piece& p,
                                           u32 c0, u32* err) {
  if (p.label > n_labels) { atomicOr(err, 1u); return; }
  abx_object_rec* rec = recs + base + (p.label - 1);
  atomicAdd(&rec->n, p.n);
  atomicAdd(reinterpret_cast<u64*>(&rec->sum_row), (u64)p.sum_row);
  atomicAdd(reinterpret_cast<u64*>(&rec->sum_col), (u64)p.sum_col);
  atomicMin(&rec->rmin, p.rmin);
  atomicMax(&rec->rmax, p.rmax);
  atomicMin(&rec->cmin, c0 + (u32)__ffs(p.cmask) - 1u);
  atomicMax(&rec->cmax, c0 + 31u - (u32)__clz(p.cmask));
}

// bit j of the result <-> halfword j of (w0..w3) equals `label`
__device__ __forceinline__ u32 eq_mask8(const u32 (&w)[4], u32 label) {
  const u32 pair = label | (label << 16);
  // "halfword != 0" as min(halfword, 1) on both halves at once (VIMNMX.U16x2): bit 0 / 16 of nz[j] is set iff the low /
  // high halfword of word j differs from the label
  u32 nz[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) nz[j] = __vminu2(w[j] ^ pair, 0x00010001u);
  // gather the eight flag bits (disjoint, so sums are ors — and the multiply-adds run on the FMA pipe, not the ALU):
  // low halfwords -> even positions, high halfwords -> odd positions
  const u32 a = nz[0] + 4u * nz[1] + 16u * nz[2] + 64u * nz[3];  // bits {0,16},{2,18},{4,20},{6,22}
  const u32 ne = (a & 0x55u) | ((a >> 15) & 0xAAu);
  return ne ^ 0xFFu;
}

__device__ __forceinline__ u32 bitpos_sum8(u32 m) {  // sum of the positions of the set bits of an 8-bit mask
  return (u32)__popc(m & 0xAAu) + 2u * (u32)__popc(m & 0xCCu) + 4u * (u32)__popc(m & 0xF0u);
}

// Column-strip walker.  A warp owns a band of band_rows (32; 16 or 8 for small launches, so that a time point of a few
// tiles still fills the GPU with short walks) rows x 256 columns; lane l walks DOWN the
// 8-pixel strip [c0, c0 + 8): one 128-bit cp.async per row brings the strip into the lane's private
// shared-memory slots (the 32 lanes of a row form one coalesced 512-byte request; kRowBatch rows per
// group, two groups in flight), and the object the lane is inside of lives in registers: a row whose
// 8 labels all equal the current label costs a handful of instructions, and an object is flushed to
// its global record with one set of atomics per (strip, band) instead of one per row run.
constexpr int kBandRows = 32;
constexpr int kRowBatch = 8;  // rows per cp.async group

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool kTma, bool kBits>
__global__ void __launch_bounds__(kScanThreads)
label_scan_kernel(const __grid_constant__ CUtensorMap tmap, const uint16_t* __restrict__ labels, int n_planes, int H, int W,
                  i64 plane_stride, i64 row_stride, const int32_t* __restrict__ plane_base,
                  abx_object_rec* __restrict__ recs, int n_objects, int vec_ok, u32* err, u64* __restrict__ bitmaps,
                  int count_background, int band_rows) {
  __shared__ __align__(128) uint4 stage_all[kScanWarps][2][kRowBatch][32];  // 8 KB per warp
  __shared__ __align__(8) u64 bars[kScanWarps][2];
  const u32 lane = lane_id();
  const int warp = threadIdx.x >> 5;
  uint4 (*stage)[kRowBatch][32] = stage_all[warp];
  const i64 gwarp = (i64)blockIdx.x * kScanWarps + warp;
  const i64 nwarps = (i64)gridDim.x * kScanWarps;
  const int col_groups = (W + 255) >> 8;
  const int bands = (H + band_rows - 1) / band_rows;
  const i64 total = (i64)n_planes * bands * col_groups;
  u32 parity0 = 0, parity1 = 0;  // phase of the two mbarriers (warp-uniform)
  if (kTma) {
    if (lane == 0) {
      mbar_init(smem_addr_of(&bars[warp][0]), 1);
      mbar_init(smem_addr_of(&bars[warp][1]), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }

  for (i64 unit = gwarp; unit < total; unit += nwarps) {
    const int cg = (int)(unit % col_groups);
    const i64 t = unit / col_groups;
    const int band = (int)(t % bands);
    const int p = (int)(t / bands);
    const u32 c0 = (u32)cg * 256u + lane * 8u;
    if (!kTma && c0 >= (u32)W) continue;  // (TMA: the whole warp stays converged; such lanes read zeros)
    const bool full = vec_ok && c0 + 8u <= (u32)W;  // aligned 128-bit copies are legal for this strip
    const int r_begin = band * band_rows, r_end = min(H, r_begin + band_rows);
    const uint16_t* src = labels + (i64)p * plane_stride + c0;
    const int base = plane_base[p];
    const u32 n_labels = (u32)(plane_base[p + 1] - base);

    // rows [r0, r0 + kRowBatch) -> stage[buf]
    auto issue = [&](int r0, int buf) {
      if constexpr (kTma) {
        __syncwarp();  // every lane has finished reading the buffer that is overwritten
        if (lane == 0) {
          const u32 bar = smem_addr_of(&bars[warp][buf]);
          mbar_expect_tx(bar, kRowBatch * 512u);
          tma_box_3d(smem_addr_of(&stage[buf][0][0]), &tmap, cg * 256, r0, p, bar);
        }
      } else {
      // fallback (unaligned strides, planes smaller than a box): the lane copies what it will read back itself
#pragma unroll
      for (int u = 0; u < kRowBatch; ++u) {
        const int r = r0 + u;
        if (r >= r_end) break;
        const uint16_t* row = src + (i64)r * row_stride;
        if (full) {
          cp_async16(&stage[buf][u][lane], row);
        } else {  // ragged right edge or unaligned rows: element loads, zero beyond the plane
          uint16_t* dst = reinterpret_cast<uint16_t*>(&stage[buf][u][lane]);
#pragma unroll 1
          for (int j = 0; j < 8; ++j) dst[j] = (c0 + j < (u32)W) ? __ldg(row + j) : (uint16_t)0;
        }
      }
      cp_async_commit();
      }
    };
    auto wait_for = [&](int buf, bool more_in_flight) {
      if (kTma) {
        if (buf == 0) { mbar_wait(smem_addr_of(&bars[warp][0]), parity0); parity0 ^= 1u; }
        else { mbar_wait(smem_addr_of(&bars[warp][1]), parity1); parity1 ^= 1u; }
      } else if (more_in_flight) {
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
    };

    Piece cur;
    cur.label = 0; cur.n = 0; cur.sum_row = 0; cur.sum_col = 0; cur.rmin = 0; cur.rmax = 0; cur.cmask = 0;
    BitCol cur_bits{nullptr, 0};  // the tracked object's bitmap column for this strip
    // label-0 pixels of this strip that lie inside the plane (a box that sticks out is zero-filled: not background)
    const u32 vmask = c0 >= (u32)W ? 0u : (c0 + 8u <= (u32)W ? 0xFFu : (0xFFu >> (c0 + 8u - (u32)W)));
    const u32 vcount = (u32)__popc(vmask);
    u32 zeros = 0;
    issue(r_begin, 0);
    int buf = 0;
#pragma unroll 1
    for (int r0 = r_begin; r0 < r_end; r0 += kRowBatch, buf ^= 1) {
      const bool more = r0 + kRowBatch < r_end;
      if (more) issue(r0 + kRowBatch, buf ^ 1);
      wait_for(buf, more);
      const int rows = min(kRowBatch, r_end - r0);
#pragma unroll 1
      for (int u = 0; u < rows; ++u) {
        const u32 r = (u32)(r0 + u);
        const uint4 q = stage[buf][u][lane];
        const u32 w[4] = {q.x, q.y, q.z, q.w};
        const u32 pair = cur.label | (cur.label << 16);
        if (((w[0] ^ pair) | (w[1] ^ pair) | (w[2] ^ pair) | (w[3] ^ pair)) == 0u) {  // all 8 == current label
          if (cur.label) {
            cur.n += 8u; cur.sum_row += 8u * (r + 1u); cur.sum_col += 8u * c0 + 36u;
            cur.cmask = 0xFFu; cur.rmax = r;
            if (kBits) bitmap_or(cur_bits, r, 0xFFu);
          } else {
            zeros += vcount;
          }
          continue;
        }
        u32 rest = 0xFFu & ~eq_mask8(w, 0u);  // labelled pixels of this row
        zeros += (u32)__popc(~rest & vmask);
        if (cur.label) {
          const u32 m = eq_mask8(w, cur.label);
          if (m) {
            const u32 k = (u32)__popc(m);
            cur.n += k; cur.sum_row += k * (r + 1u); cur.sum_col += k * (c0 + 1u) + bitpos_sum8(m);
            cur.cmask |= m; cur.rmax = r;
            rest &= ~m;
            if (kBits) bitmap_or(cur_bits, r, m);
          } else {  // the object ended above this row
            emit_piece(recs, base, n_labels, cur, c0, err);
            cur.label = 0;
          }
        }
        while (rest) {  // other labels inside the 8 pixels
          const u32 pos = (u32)__ffs(rest) - 1u;
          const u32 wsel = pos < 4u ? (pos < 2u ? w[0] : w[1]) : (pos < 6u ? w[2] : w[3]);
          const u32 lbl = (wsel >> ((pos & 1u) << 4)) & 0xFFFFu;
          const u32 m = eq_mask8(w, lbl);
          const u32 k = (u32)__popc(m);
          Piece np;
          np.label = lbl; np.n = k; np.sum_row = k * (r + 1u); np.sum_col = k * (c0 + 1u) + bitpos_sum8(m);
          np.rmin = r; np.rmax = r; np.cmask = m;
          BitCol bits{nullptr, 0};
          if (kBits) {
            bits = bitmap_col(bitmaps, base, n_labels, n_objects, lbl, c0);
            bitmap_or(bits, r, m);
          }
          if (cur.label == 0) { cur = np; cur_bits = bits; }      // becomes the tracked object
          else emit_piece(recs, base, n_labels, np, c0, err);     // second object in the strip: per-row piece
          rest &= ~m;
        }
      }
    }
    if (cur.label) emit_piece(recs, base, n_labels, cur, c0, err);
    if (count_background) {  // the plane's background record counts its own pixels (one atomic per warp and unit)
      zeros = __reduce_add_sync(0xFFFFFFFFu, zeros);
      if (lane == 0 && zeros) atomicAdd(&recs[n_objects + p].n, zeros);
    }
  }
}

// per-plane maximum label: one warp per row at a time, 128-bit loads when the layout allows
__global__ void label_max_kernel(const uint16_t* __restrict__ labels, int n_planes, int H, int W, i64 plane_stride,
                                 i64 row_stride, int vec_ok, int32_t* out_max) {
  // grid: (blocks_per_plane, n_planes); out_max zeroed by the caller
  const int p = blockIdx.y;
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5;
  const u32 lane = lane_id();
  const uint16_t* plane = labels + (i64)p * plane_stride;
  u32 m = 0;
  for (int r = blockIdx.x * warps + warp; r < H; r += gridDim.x * warps) {
    const uint16_t* row = plane + (i64)r * row_stride;
    int c = 0;
    if (vec_ok) {
      u32 a = 0, b = 0;
      for (c = (int)lane * 8; c + 8 <= W; c += 256) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + c));
        a = __vmaxu2(a, __vmaxu2(q.x, q.y));
        b = __vmaxu2(b, __vmaxu2(q.z, q.w));
      }
      a = __vmaxu2(a, b);
      m = max(m, max(a & 0xFFFFu, a >> 16));
      c = W & ~7;  // the ragged tail below
    }
    for (int x = c + (int)lane; x < W; x += 32) m = max(m, (u32)__ldg(row + x));
  }
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  if (lane == 0 && m) atomicMax(out_max + p, (int32_t)m);
}

}  // namespace

// The tensor map of the label planes (u16, dims W x H x P, box 256 x kRowBatch x 1), or false when the layout
// does not qualify for TMA (unaligned base / strides, planes smaller than one box) or the driver lacks the encoder.
static bool make_label_tensor_map(const abx_extract_args* a, CUtensorMap* tm) {
  EncodeFn encode = tensor_map_encoder();
  if (!encode) return false;
  const i64 plane_stride = a->n_planes > 1 ? a->label_plane_stride : (i64)a->H * a->label_row_stride;
  if ((reinterpret_cast<uintptr_t>(a->labels) & 15u) || a->label_row_stride % 8 || plane_stride % 8 || a->W < 256 ||
      a->H < kRowBatch || a->label_row_stride < a->W || plane_stride < (i64)a->H * a->label_row_stride)
    return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->n_planes};
  const cuuint64_t gstr[2] = {(cuuint64_t)a->label_row_stride * 2u, (cuuint64_t)plane_stride * 2u};
  const cuuint32_t box[3] = {256u, (cuuint32_t)kRowBatch, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(a->labels), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// bitmaps: [n_objects + 1][64] u64 torus bitmaps (the last one is a dummy for out-of-range labels), or nullptr
int launch_label_scan(const abx_extract_args* a, abx_object_rec* recs, u32* err, u64* bitmaps, double* sqrt_tab, int n_sqrt,
                      cudaStream_t st) {
  const int n_rec = a->n_objects + a->n_planes;
  const size_t n_words = bitmaps ? ((size_t)a->n_objects + 1) * 32 : 0;  // 512 bytes per bitmap
  size_t threads = (size_t)n_rec > n_words / 8 ? (size_t)n_rec : n_words / 8;  // <= 8 words per thread
  if (threads < (size_t)n_sqrt) threads = (size_t)n_sqrt;
  if (threads < (size_t)kCounterWords) threads = kCounterWords;
  init_records_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(recs, a->n_objects, a->n_planes, a->H, a->W, err,
                                                                          reinterpret_cast<uint4*>(bitmaps), n_words, sqrt_tab,
                                                                          sqrt_tab ? n_sqrt : 0);
  // rows per warp: 32, or fewer (a multiple of the 8-row load batch) while the launch has less than one unit per resident warp
  int band_rows = kBandRows;
  auto n_units = [&](int br) { return (i64)a->n_planes * ((a->H + br - 1) / br) * ((a->W + 255) / 256); };
  while (band_rows > kRowBatch && n_units(band_rows) < 148 * 7 * kScanWarps) band_rows /= 2;
  const i64 units = n_units(band_rows);
  if (units == 0) return abx_check_cuda(cudaGetLastError(), "init_records");
  i64 blocks = (units + kScanWarps - 1) / kScanWarps;
  const i64 cap = 148 * 7;  // 7 CTAs of 4 warps per SM (shared memory); more units are walked in a grid-stride loop
  if (blocks > cap) blocks = cap;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(a->labels) & 15u) == 0) && (a->label_row_stride % 8 == 0) &&
                     (a->label_plane_stride % 8 == 0);
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  const bool tma = make_label_tensor_map(a, &tm);
#define ABX_SCAN(TMA, BITS)                                                                                              \
  label_scan_kernel<TMA, BITS><<<(int)blocks, kScanThreads, 0, st>>>(                                                     \
      tm, static_cast<const uint16_t*>(a->labels), a->n_planes, a->H, a->W, a->label_plane_stride, a->label_row_stride,  \
      a->plane_base, recs, a->n_objects, vec_ok, err, bitmaps, a->with_background, band_rows)
  if (tma && bitmaps) ABX_SCAN(true, true);
  else if (tma) ABX_SCAN(true, false);
  else if (bitmaps) ABX_SCAN(false, true);
  else ABX_SCAN(false, false);
#undef ABX_SCAN
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
  int bx = (H + 7) / 8;  // 8 warps per CTA, a row per warp at a time
  if (bx > 64) bx = 64;
  dim3 grid(bx < 1 ? 1 : bx, n_planes);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(labels) & 15u) == 0) && (row_stride % 8 == 0) && (plane_stride % 8 == 0);
  label_max_kernel<<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(labels), n_planes, H, W, plane_stride,
                                         row_stride, vec_ok, out_max);
  return abx_check_cuda(cudaGetLastError(), "label_max");
}
