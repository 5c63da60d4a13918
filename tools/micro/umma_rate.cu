// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N, with the A operand in
// shared memory (SS) or in tensor memory (TS), issued back to back by one thread into rotating accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I mastermetastyletransfer_b200/csrc -o tools/micro/umma_rate tools/micro/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace mst;


template <int N, int TS, int SAME>
__global__ void k(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (__shfl_sync(0xffffffffu, warp, 0) == 0) {  // whole warp, convergent: descriptors stay in uniform registers
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = umma_desc_sw128(base), bd = umma_desc_sw128(base + 16384);
    constexpr int NACC = SAME ? 1 : (256 / N < 1 ? 1 : (256 / N > 4 ? 4 : 256 / N));  // independent accumulators (columns 0..255)
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {  // four K steps into one accumulator, then the next (as the kernels do); constant offsets only
        const uint32_t d = tm + ((j >> 2) % NACC) * N;
        if (TS) umma_ts_pred(d, tm + 256 + (j & 3) * 8, bd + (uint64_t)((j & 3) * 2), idesc, 1);
        else umma_bf16_pred(d, ad + (uint64_t)((j & 3) * 2), bd + (uint64_t)((j & 3) * 2), idesc, 1);
      }
    }
    umma_commit_pred(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int TS, int SAME>
static double run(int iters, long long* out) {
  cudaFuncSetAttribute(k<N, TS, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k<N, TS, SAME><<<148, 128, 64 * 1024>>>(iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  return (double)mx / iters;
}
template <int N>
static void row(int iters, long long* out) {
  printf("%6d %14.1f %14.1f %14.1f %14.1f   (math floor 128*N/256 = %d)\n", N, run<N, 0, 0>(iters, out), run<N, 0, 1>(iters, out),
         run<N, 1, 0>(iters, out), run<N, 1, 1>(iters, out), N / 2);
}
int main() {
  long long* out;
  cudaMalloc(&out, 148 * 8);
  const int iters = 4000;
  printf("cycles per tcgen05.mma (M=128, K=16, bf16), %d back-to-back instructions from one warp (uniform issue), all 148 SMs busy\n", iters);
  printf("%6s %14s %14s %14s %14s\n", "N", "SS rotating", "SS same acc", "TS rotating", "TS same acc");
  row<16>(iters, out); row<32>(iters, out); row<64>(iters, out); row<96>(iters, out); row<128>(iters, out); row<192>(iters, out); row<256>(iters, out);
  return 0;
}
