// Shape metrics of window-sized objects (bounding box <= 64 x 64): ONE WARP PER OBJECT, everything on bitmasks
// and a 16-bit grid in shared memory — no pixel list, and no label window: the 64-bit row masks come straight from
// the torus bitmap the label scan wrote for the object (label_scan.cu; 512 bytes, two coalesced 64-bit loads per
// lane, rotated by the bounding box origin).  The three chained exact EDTs of
// src/extraction/core/functions/cell.py:176-229 (conical_volume, min_maj_approximation -> eccentricity, volume):
//
//   phase M  torus bitmap -> 64-bit row masks
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
// Slot: row masks 512 B | top masks 512 B | run ends 256 B | grid u16 [72][64] = 10 496 B per warp,
// 10 warps per CTA, 2 CTAs per SM.  Objects come in the order the plan kernel wrote (object_sweep.cu: big ones first);
// larger windows go to the CTA-per-object kernel (shape_edt.cu).
#include <cstring>

#include "common.cuh"

namespace {

#include "warp_common.cuh"
#include "edt_phases.cuh"

constexpr int kGridWarps = 10;
constexpr u32 kTopOff = 512, kInfoOff = 1024, kGOff = 1280;
constexpr u32 kGridSlot = kGOff + kEdtGridBytes;  // 10 496

__global__ void __launch_bounds__(kGridWarps * 32, 2)
object_edt_grid(const abx_object_rec* __restrict__ recs, const u64* __restrict__ bitmaps, const int* __restrict__ order,
                const u32* __restrict__ order_counts /* [0] big, [1] small */, int n_objects, u32* __restrict__ work_counter,
                int want_conical, const double* __restrict__ sqrt_tab, ShapeStats* __restrict__ shape) {
  const u32 lane = lane_id();
  const u32 slot_off = (threadIdx.x >> 5) * kGridSlot;
  u64* rowmask = reinterpret_cast<u64*>(dyn + slot_off);
  // the margin above the window stays zero for the life of the kernel
  *reinterpret_cast<uint4*>(dyn + slot_off + kGOff + 16u * lane) = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const u32 n_big = order_counts[0];
  const int n_items = (int)(n_big + order_counts[1]);
  Queue qu{work_counter, n_items, 0};
  int item = qu.fetch();
  int nxt = item < n_items ? qu.fetch() : n_items;
  while (item < n_items) {
    const int obj = order[(u32)item < n_big ? item : n_objects - 1 - (item - (int)n_big)];
    const abx_object_rec rec = recs[obj];
    // ---- phase M: row masks from the torus bitmap, bit c of row r <-> pixel (rmin + r, cmin + c) ----
    const u64* bm = bitmaps + (size_t)obj * 64u;
    u64 m0 = bm[(rec.rmin + lane) & 63u], m1 = bm[(rec.rmin + lane + 32u) & 63u];
    const u32 rot = rec.cmin & 63u;
    m0 = (m0 >> rot) | (rot ? (m0 << (64u - rot)) : 0ull);
    m1 = (m1 >> rot) | (rot ? (m1 << (64u - rot)) : 0ull);
    __syncwarp();
    rowmask[lane] = m0;
    rowmask[lane + 32] = m1;
    shape_from_masks(rec, 0u, slot_off, slot_off + kTopOff, slot_off + kInfoOff, slot_off + kGOff, false, want_conical != 0,
                     sqrt_tab, shape + obj);
    item = nxt;
    nxt = item < n_items ? qu.fetch() : n_items;
  }
}

}  // namespace

// sqrt(d2) for every squared distance the first EDT of a 64 x 64 window can produce (exact: IEEE sqrt)
int abx_sqrt_table_entries() { return (kSide / 2) * (kSide / 2) + 1; }  // row distances are <= 32

int launch_object_edt_warp(const abx_extract_args* a, const Workspace& ws, cudaStream_t st) {
  if (!(a->need_edt & 3) || a->n_objects == 0) return ABX_OK;
  constexpr size_t smem = (size_t)kGridWarps * kGridSlot;
  static thread_local bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(object_edt_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return abx_check_cuda(e, "object_edt_grid smem attribute");
    done[dev] = true;
  }
  // (ws.sqrt_tab was filled by the label scan's first kernel)
  int grid = (a->n_objects + kGridWarps - 1) / kGridWarps;
  if (grid > 148 * 2) grid = 148 * 2;  // persistent: 2 CTAs per SM, warps pull objects from a counter
  object_edt_grid<<<grid, kGridWarps * 32, smem, st>>>(ws.recs, ws.bitmaps, ws.order_edt, ws.list_counts + kCntEdtBig,
                                                       a->n_objects, ws.list_counts + kCntEdtWork,
                                                       (a->need_edt & 2) != 0, ws.sqrt_tab, ws.shape);
  return abx_check_cuda(cudaGetLastError(), "object_edt_grid");
}
