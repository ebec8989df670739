// Shape metrics of window-sized objects (bounding box <= 64 x 64): ONE WARP PER OBJECT, everything on bitmasks
// and a 16-bit grid in shared memory — no pixel list.  The three chained exact EDTs of
// src/extraction/core/functions/cell.py:176-229 (conical_volume, min_maj_approximation -> eccentricity, volume):
//
//   phase M  label window -> 64-bit row masks.  The window comes in by TMA (cp.async.bulk.tensor boxes of
//            64 columns x 8 rows, started at a 16-byte aligned column left of the bbox) when the label layout
//            qualifies and the shifted window still fits 64 columns; by plain loads otherwise.
//   phase R  squared row distances g^2 (u16) of every cell, from the run ends of the row mask
//   phase C  EDT 1, exact column pass with packed 16-bit arithmetic: a lane owns two adjacent columns and four
//            rows at a time; one 32-bit shared-memory load brings a source row for both columns and one
//            VIADDMNMX.U16x2 (min(g^2 + d^2, best) on both halves) updates a target row.  Four zero rows above
//            and below the window bound every walk (a finished pixel's candidates are >= d^2 >= its best).
//   phase T  cone top = pixels that attain the maximum, as a second set of row masks
//   phase 2  EDT 2 only where it is needed, max over the object of the distance to the nearest top pixel:
//            one top (half of all cells) -> the row ends decide; up to four -> packed column terms kept in
//            registers; more -> per-row loop over the tops
//   phase 3  EDT 3 on the (small) cone top, lanes over rows
//
// Slot: row masks 512 B | top masks 512 B | mbarrier 128 B | run ends 256 B | grid u16 [72][64] = 10 624 B per warp,
// 10 warps per CTA, 2 CTAs per SM.  Larger windows go to the CTA-per-object kernel (shape_edt.cu).
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <cstring>

#include "common.cuh"

namespace {

#include "warp_common.cuh"
#include "tma.cuh"
#include "edt_phases.cuh"

constexpr int kGridWarps = 10;
constexpr u32 kTopOff = 512, kGBarOff = 1024, kInfoOff = 1152, kGOff = 1408;
constexpr u32 kGridSlot = kGOff + kEdtGridBytes;  // 10 624
static_assert(kGOff % 128 == 0, "TMA destinations are 128-byte aligned");

__device__ __forceinline__ void shape_object(const abx_object_rec& rec, int p, u32 label, const Common& cm,
                                             const CUtensorMap* lab_map, bool use_tma, u32 slot_off, u32 bar, u32& parity,
                                             bool want_conical, const double* __restrict__ sqrt_tab,
                                             ShapeStats* __restrict__ dst) {
  u64* rowmask = reinterpret_cast<u64*>(dyn + slot_off);
  u64* topmask = reinterpret_cast<u64*>(dyn + slot_off + kTopOff);
  const u32 g_off = slot_off + kGOff;  // grid row j of the window lives at row j + kMargin
  const u32 lane = lane_id();
  const int h = (int)(rec.rmax - rec.rmin) + 1, w = (int)(rec.cmax - rec.cmin) + 1;
  const u32 n = rec.n;
  // ---- phase M: row masks ----
  u32 s_lab = rec.cmin & 7u;  // the TMA box starts at a 16-byte multiple: s_lab columns left of the bbox
  const bool by_tma = use_tma && (u32)w + s_lab <= 64u;
  if (!by_tma) s_lab = 0;
  __syncwarp();
  if (by_tma) {
    const u32 h8 = ((u32)h + 7u) & ~7u;
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the grid was written by generic stores
      mbar_expect_tx(bar, h8 * 128u);
      const u32 dst0 = smem_addr(dyn + g_off + kMargin * 128u);
      for (u32 j = 0; j < h8; j += 8) tma_box_3d(dst0 + j * 128u, lab_map, (int)(rec.cmin - s_lab), (int)(rec.rmin + j), p, bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    const unsigned short* lw = reinterpret_cast<const unsigned short*>(dyn + g_off + kMargin * 128u);
#pragma unroll 4
    for (int r = 0; r < h; ++r) {
      // columns outside the bbox hold other labels (or the hardware's zero fill): they never match
      const u32 b0 = __ballot_sync(kFull, (u32)lw[r * 64 + lane] == label);
      const u32 b1 = __ballot_sync(kFull, (u32)lw[r * 64 + 32 + lane] == label);
      if (lane == 0) rowmask[r] = (u64)b0 | ((u64)b1 << 32);
    }
  } else {
    const uint16_t* lab = cm.labels + (i64)p * cm.lab_plane_stride + (i64)rec.rmin * cm.lab_row_stride + rec.cmin;
#pragma unroll 4
    for (int r = 0; r < h; ++r) {
      const uint16_t* lrow = lab + (i64)r * cm.lab_row_stride;
      const u32 l0 = lane < (u32)w ? (u32)__ldg(lrow + lane) : kFull;
      const u32 l1 = lane + 32u < (u32)w ? (u32)__ldg(lrow + lane + 32) : kFull;
      const u32 b0 = __ballot_sync(kFull, l0 == label);
      const u32 b1 = __ballot_sync(kFull, l1 == label);
      if (lane == 0) rowmask[r] = (u64)b0 | ((u64)b1 << 32);
    }
  }
  shape_from_masks(rec, s_lab, slot_off, slot_off + kTopOff, slot_off + kInfoOff, g_off, false, want_conical, sqrt_tab, dst);
}

__global__ void __launch_bounds__(kGridWarps * 32, 2)
object_edt_grid(const __grid_constant__ CUtensorMap lab_map, int use_tma, const Common cm, int want_conical,
                const double* __restrict__ sqrt_tab, ShapeStats* __restrict__ shape, int* __restrict__ edt_list,
                u32* __restrict__ edt_count) {
  const u32 lane = lane_id();
  const u32 slot_off = (threadIdx.x >> 5) * kGridSlot;
  const u32 bar = smem_addr(dyn + slot_off + kGBarOff);
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the margin above the window stays zero for the life of the kernel
  *reinterpret_cast<uint4*>(dyn + slot_off + kGOff + 16u * lane) = make_uint4(0, 0, 0, 0);
  __syncwarp();
  u32 parity = 0;
  Queue qu{cm.counters, cm.n_objects, 0};
  int obj = qu.fetch();
  int nxt = obj < cm.n_objects ? qu.fetch() : cm.n_objects;
  while (obj < cm.n_objects) {
    if (nxt < cm.n_objects) {  // L2 prefetch of the next object's label window
      const abx_object_rec nr = cm.recs[nxt];
      const int nh = (int)(nr.rmax - nr.rmin) + 1, nw = (int)(nr.cmax - nr.cmin) + 1;
      if (nr.n > 0 && nh <= kSide && nw <= kSide) {
        const int np = find_plane(cm.plane_base, cm.n_planes, nxt);
        prefetch_rows(cm.labels + (i64)np * cm.lab_plane_stride + (i64)nr.rmin * cm.lab_row_stride + nr.cmin,
                      cm.lab_row_stride * 2, nh, (u32)nw * 2u);
      }
    }
    const abx_object_rec rec = cm.recs[obj];
    const int h = (int)(rec.rmax - rec.rmin) + 1, w = (int)(rec.cmax - rec.cmin) + 1;
    if (rec.n == 0) {
      if (lane == 0) { ShapeStats z; z.sum_nn = 0; z.sum_top = 0; z.max_nn2 = 0; z.max_dn2 = 0; shape[obj] = z; }
    } else if (h > kSide || w > kSide) {
      if (lane == 0) edt_list[atomicAdd(edt_count, 1u)] = obj;  // hand over to the CTA-per-object kernel
    } else {
      const int p = find_plane(cm.plane_base, cm.n_planes, obj);
      shape_object(rec, p, (u32)(obj - cm.plane_base[p] + 1), cm, &lab_map, use_tma != 0, slot_off, bar, parity,
                   want_conical != 0, sqrt_tab, shape + obj);
    }
    obj = nxt;
    nxt = obj < cm.n_objects ? qu.fetch() : cm.n_objects;
  }
}

// Label planes (W, H, P) with a box of 64 columns x 8 rows, or false when the layout does not qualify for TMA.
bool make_label_map(const abx_extract_args* a, CUtensorMap* tm) {
  EncodeFn encode = tensor_map_encoder();
  if (!encode) return false;
  const i64 lab_ps = a->n_planes > 1 ? a->label_plane_stride : (i64)a->H * a->label_row_stride;
  if ((reinterpret_cast<uintptr_t>(a->labels) & 15u) || a->label_row_stride % 8 || lab_ps % 8 || a->W < 64 || a->H < 8 ||
      a->label_row_stride < a->W || lab_ps < (i64)a->H * a->label_row_stride)
    return false;
  const cuuint64_t ldim[3] = {(cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->n_planes};
  const cuuint64_t lstr[2] = {(cuuint64_t)a->label_row_stride * 2u, (cuuint64_t)lab_ps * 2u};
  const cuuint32_t lbox[3] = {64u, 8u, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(a->labels), ldim, lstr, lbox, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void sqrt_table_kernel(double* tab, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tab[i] = sqrt((double)i);
}

}  // namespace

// sqrt(d2) for every squared distance the first EDT of a 64 x 64 window can produce (exact: IEEE sqrt)
int abx_sqrt_table_entries() { return (kSide / 2) * (kSide / 2) + 1; }  // row distances are <= 32

int launch_object_edt_warp(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  if (!a->need_edt || a->n_objects == 0) return ABX_OK;
  constexpr size_t smem = (size_t)kGridWarps * kGridSlot;
  static thread_local bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(object_edt_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return abx_check_cuda(e, "object_edt_grid smem attribute");
    done[dev] = true;
  }
  const int n_tab = abx_sqrt_table_entries();
  sqrt_table_kernel<<<(n_tab + 255) / 256, 256, 0, st>>>(ws.sqrt_tab, n_tab);
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  const int use_tma = make_label_map(a, &tm) ? 1 : 0;
  Common cm;
  cm.labels = static_cast<const uint16_t*>(a->labels);
  cm.lab_plane_stride = a->label_plane_stride;
  cm.lab_row_stride = a->label_row_stride;
  cm.plane_tile = a->plane_tile;
  cm.plane_base = a->plane_base;
  cm.n_planes = a->n_planes;
  cm.n_objects = a->n_objects;
  cm.n_total = a->n_objects;
  cm.recs = ws.recs;
  cm.counters = ws.list_counts + 4;
  int grid = (a->n_objects + kGridWarps - 1) / kGridWarps;
  if (grid > 148 * 2) grid = 148 * 2;  // persistent: 2 CTAs per SM, warps pull objects from a counter
  object_edt_grid<<<grid, kGridWarps * 32, smem, st>>>(tm, use_tma, cm, (a->need_edt & 2) != 0, ws.sqrt_tab, ws.shape,
                                                       ws.edt_list, ws.list_counts + 1);
  return abx_check_cuda(cudaGetLastError(), "object_edt_grid");
}
