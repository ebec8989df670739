// Standalone probe: which (box width, promotion, smem offset, param form) combinations of a 3-D u16 TMA box load work.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
typedef unsigned int u32;
extern __shared__ __align__(128) unsigned char dyn[];
struct Maps { CUtensorMap m[4]; };
__device__ __forceinline__ u32 saddr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
template <int kMode>
__global__ void probe(const __grid_constant__ Maps maps, const __grid_constant__ CUtensorMap single, int which, int bw, int x, int y, int z, u32 off, unsigned short* out) {
  const u32 lane = threadIdx.x & 31;
  const u32 bar = saddr(dyn + 64);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const u32 dst = saddr(dyn + off);
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bw * 8 * 2) : "memory");
    const CUtensorMap* tm = kMode == 0 ? &single : (kMode == 1 ? &maps.m[1] : (kMode == 2 ? (which == 0 ? &maps.m[0] : which == 1 ? &maps.m[1] : which == 2 ? &maps.m[2] : &maps.m[3]) : &maps.m[which]));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
  }
  u32 done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(0) : "memory");
  } while (!done);
  const unsigned short* w = reinterpret_cast<const unsigned short*>(dyn + off);
  for (int i = lane; i < bw * 8; i += 32) out[i] = w[i];
}
#include <cstdlib>
int main(int argc, char** argv) {
  const int W = 2160, H = 2160, P = 8;
  std::vector<unsigned short> h((size_t)W * H * P);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned short)(i * 2654435761u >> 16);
  unsigned short *d, *out;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&out, 256 * 8 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn encode = (EncodeFn)fn;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  // argv: bw promo mode off x
  const int bw = atoi(argv[1]), promo = atoi(argv[2]), mode = atoi(argv[3]), x = atoi(argv[5]);
  const u32 off = (u32)atoi(argv[4]);
  const int y = 777, z = 3;
  Maps maps; memset(&maps, 0, sizeof(maps));
  const int bws[4] = {32, 64, 128, 256};
  int which = 0;
  for (int i = 0; i < 4; ++i) {
    if (bws[i] == bw) which = i;
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P};
    const cuuint64_t gstr[2] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2};
    const cuuint32_t box[3] = {(cuuint32_t)bws[i], 8u, 1u};
    const cuuint32_t es[3] = {1, 1, 1};
    encode(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (mode == 0) probe<0><<<1, 32, 200 * 1024>>>(maps, maps.m[which], which, bw, x, y, z, off, out);
  if (mode == 1) probe<1><<<1, 32, 200 * 1024>>>(maps, maps.m[which], which, bw, x, y, z, off, out);
  if (mode == 2) probe<2><<<1, 32, 200 * 1024>>>(maps, maps.m[which], which, bw, x, y, z, off, out);
  if (mode == 3) probe<3><<<1, 32, 200 * 1024>>>(maps, maps.m[which], which, bw, x, y, z, off, out);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned short> o(bw * 8);
  int bad = -1;
  if (e == cudaSuccess) {
    cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost);
    bad = 0;
    for (int r = 0; r < 8; ++r)
      for (int c = 0; c < bw; ++c)
        bad += o[r * bw + c] != (x + c < W ? h[((size_t)z * H + y + r) * W + x + c] : 0);
  }
  printf("mode=%d bw=%3d promo=%d off=%6u x=%d : %s bad=%d\n", mode, bw, promo, off, x, cudaGetErrorString(e), bad);
  return 0;
}
