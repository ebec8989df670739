// TMA (cp.async.bulk.tensor) and mbarrier plumbing shared by the kernels that stage tiles in shared memory
// (label_scan.cu, object_sweep.cu).  Include inside the translation unit's anonymous namespace, after <cuda.h> and
// common.cuh.  One elected lane issues the boxes; the 32 lanes wait on the slot's mbarrier.
// Measured on B200 (tools/probes/tma_probe.cu): the innermost box coordinate times the element size must be a
// multiple of 16 bytes, otherwise the copy faults with "illegal instruction"; rows may start anywhere.
#pragma once

__device__ __forceinline__ u32 smem_addr_of(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
  u32 done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_box_2d(u32 dst, const CUtensorMap* tmap, int x, int y, u32 bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_box_3d(u32 dst, const CUtensorMap* tmap, int x, int y, int z, u32 bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
// L2 prefetch of one box (same tensor map and coordinates as the copy that follows later): fire and forget
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tmap, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tmap), "r"(x), "r"(y) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn tensor_map_encoder() {
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  return encode;
}

