// Microbenchmark: TMEM -> register read throughput (tcgen05.ld 32x32b.x32) per SM as a function of the number of warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tmem_ld_bench tools/micro/tmem_ld_bench.cu && ./tools/micro/tmem_ld_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__global__ void bench(int iters, int loads_in_flight, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t a[32], b[32];
    ld32(base + ((i * 64) & 511 & ~63), a);
    if (loads_in_flight > 1) ld32(base + (((i * 64) + 32) & 511), b);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int e = 0; e < 32; ++e) acc ^= a[e];
    if (loads_in_flight > 1) {
#pragma unroll
      for (int e = 0; e < 32; ++e) acc ^= b[e];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int lif = 1; lif <= 2; ++lif)
    for (int warps : {1, 2, 4, 8, 16}) {
      bench<<<148, warps * 32, 0>>>(iters, lif, out, sink);
      cudaDeviceSynchronize();
      bench<<<148, warps * 32, 0>>>(iters, lif, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * lif * warps * 32 * 32 * 4;
      printf("warps %2d loads/iter %d: %lld clk, %.1f B/clk/SM (%s)\n", warps, lif, h, bytes / h, cudaGetErrorString(e));
    }
  return 0;
}
