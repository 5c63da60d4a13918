// Microbenchmark: cost of fence.proxy.async.shared::cta after a few st.shared (the hand-off of a thread-written tcgen05 operand tile),
// per warp, as a function of the number of warps doing it.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fence_bench tools/micro/fence_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void bench(int iters, int mode, long long* out) {
  extern __shared__ uint8_t smem[];
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + threadIdx.x * 64;
  __syncthreads();
  const long long t0 = clock64();
  uint32_t v = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + q * 16), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3) : "memory");
    if (mode == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (mode == 2) __threadfence_block();
    v += 4;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
  long long* out;
  cudaMalloc(&out, 148 * 8);
  const int iters = 2000;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      bench<<<148, warps * 32, warps * 32 * 64>>>(iters, mode, out);
      cudaDeviceSynchronize();
      bench<<<148, warps * 32, warps * 32 * 64>>>(iters, mode, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("mode %d (%s) warps %2d: %.1f clk per iteration (4 x STS.128%s) %s\n", mode, mode == 0 ? "no fence" : mode == 1 ? "fence.proxy.async" : "membar.cta",
             warps, (double)h / iters, mode ? " + fence" : "", cudaGetErrorString(e));
    }
  return 0;
}
