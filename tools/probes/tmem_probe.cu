#include <cstdio>
typedef unsigned int u32;
extern __shared__ __align__(128) unsigned char dyn[];
__global__ void __launch_bounds__(128, 4) k(u32* out) {
  u32* slot = reinterpret_cast<u32*>(dyn);
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(slot)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const u32 tbase = slot[0] + ((warp * 32u) << 16);
  u32 a = lane, b = lane * 2, c = lane * 3, d = lane * 4;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tbase + 4u), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  u32 r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(tbase + 4u) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[threadIdx.x + blockIdx.x * blockDim.x] = r0 + r1 + r2 + r3;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot[0]), "n"(64));
}
int main() {
  u32* d; cudaMalloc(&d, 4 * 128 * 148 * 4);
  k<<<148 * 4, 128, 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  u32 h[128]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int ok = 1; for (int i = 0; i < 128; ++i) ok &= h[i] == (u32)((i & 31) * 10);
  printf("%s ok=%d\n", cudaGetErrorString(e), ok);
}
