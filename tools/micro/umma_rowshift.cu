// Micro test: may the start address of a swizzled K-major tcgen05 operand be offset by whole ROWS that are not a multiple of the
// 8-row swizzle atom?  (A 3x3 convolution wants tap kx = the same staged input row shifted by kx pixels.)  The operand is written
// with the swizzle taken from ABSOLUTE shared-memory address bits; the test issues M=128 x N=16 MMAs with the A start moved by
// `shift` rows, with the descriptor's base-offset field 0 or (start >> 7) & 7, and compares with the host result.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I mastermetastyletransfer_b200/csrc -o tools/micro/umma_rowshift tools/micro/umma_rowshift.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "common.cuh"
using namespace mst;
typedef __nv_bfloat16 bf16;

__host__ __device__ inline float aval(int r, int k) { return (float)(((r * 7 + k * 3) % 17) - 8); }
__host__ __device__ inline float bval(int n, int k) { return (float)(((n * 5 + k) % 13) - 6); }

// rowb = bytes per row (128: SWIZZLE_128B, 64: SWIZZLE_64B)
__global__ void k(int rowb, int shift, int use_bo, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t mask = rowb == 128 ? 7u : 3u;
  const int K = rowb / 2, cpr = rowb / 16;
  const uint32_t a_off = 0, b_off = 32768;
  for (int i = threadIdx.x; i < 160 * cpr; i += blockDim.x) {
    const int r = i / cpr, c = i % cpr;
    uint32_t addr = base + a_off + r * rowb + c * 16;
    addr ^= ((addr >> 7) & mask) << 4;
    bf16* p = reinterpret_cast<bf16*>(gen + (addr - base));
    for (int e = 0; e < 8; ++e) p[e] = __float2bfloat16(aval(r, c * 8 + e));
  }
  for (int i = threadIdx.x; i < 16 * cpr; i += blockDim.x) {
    const int r = i / cpr, c = i % cpr;
    uint32_t addr = base + b_off + r * rowb + c * 16;
    addr ^= ((addr >> 7) & mask) << 4;
    bf16* p = reinterpret_cast<bf16*>(gen + (addr - base));
    for (int e = 0; e < 8; ++e) p[e] = __float2bfloat16(bval(r, c * 8 + e));
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint64_t layout = rowb == 128 ? 2ull : 4ull;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a_addr = base + a_off + shift * rowb + ks * 32, b_addr = base + b_off + ks * 32;
      uint64_t ad = ((uint64_t)((a_addr & 0x3FFFF) >> 4)) | (1ull << 16) | ((uint64_t)((8 * rowb) >> 4) << 32) | (1ull << 46) | (layout << 61);
      uint64_t bd = ((uint64_t)((b_addr & 0x3FFFF) >> 4)) | (1ull << 16) | ((uint64_t)((8 * rowb) >> 4) << 32) | (1ull << 46) | (layout << 61);
      if (use_bo) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
      umma_bf16_pred(tm, ad, bd, umma_idesc_bf16(128, 16), ks != 0);
    }
    umma_commit_pred(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tm + ((uint32_t)(warp * 32) << 16), v);
  tmem_wait_ld();
  for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 32);
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 16 * 4);
  static float h[128 * 16];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rowb : {128, 64})
    for (int shift = 0; shift < 10; ++shift)
      for (int bo = 0; bo < 2; ++bo) {
        cudaMemset(out, 0, sizeof(h));
        k<<<1, 128, 64 * 1024>>>(rowb, shift, bo, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rowb %d shift %d bo %d: CUDA error %s\n", rowb, shift, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 16; ++n) {
            float ref = 0.f;
            for (int kk = 0; kk < rowb / 2; ++kk) ref += aval(m + shift, kk) * bval(n, kk);
            if (ref != h[m * 16 + n]) ++bad;
          }
        printf("rowb %3d shift %d base_offset %s: %s (%d of 2048 wrong)\n", rowb, shift, bo ? "set " : "zero", bad ? "MISMATCH" : "ok", bad);
      }
  return 0;
}
