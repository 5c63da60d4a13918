// Micro test: the A operand of tcgen05.mma read from TENSOR MEMORY.  A [128 x 64] bf16 (K-major, 128B-swizzled shared-memory
// tile, as a streamed weight k-block arrives) is copied to TMEM with tcgen05.cp.128x256b (one copy per K = 16 step: 128 lanes x
// 8 columns), then D = A . B^T is issued with the A operand given as a TMEM address; compared with the SS form and the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I mastermetastyletransfer_b200/csrc -o tools/micro/umma_a_tmem tools/micro/umma_a_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "common.cuh"
using namespace mst;
typedef __nv_bfloat16 bf16;

__host__ __device__ inline float aval(int r, int k) { return (float)(((r * 7 + k * 3) % 17) - 8); }
__host__ __device__ inline float bval(int n, int k) { return (float)(((n * 5 + k) % 13) - 6); }

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void k(int mode, int shift, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t a_off = 0, b_off = 32768;
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i / 8, c = i % 8;
    bf16* p = reinterpret_cast<bf16*>(gen + a_off + sw128_offset(r, c));
    for (int e = 0; e < 8; ++e) p[e] = __float2bfloat16(aval(r, c * 8 + e));
  }
  for (int i = threadIdx.x; i < 80 * 8; i += blockDim.x) {  // B rows: 64 + room for a row shift
    const int r = i / 8, c = i % 8;
    bf16* p = reinterpret_cast<bf16*>(gen + b_off + sw128_offset(r, c));
    for (int e = 0; e < 8; ++e) p[e] = __float2bfloat16(bval(r, c * 8 + e));
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 64);
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t ad = umma_desc_sw128(base + a_off + ks * 32);
      const uint64_t bd = umma_desc_sw128(base + b_off + shift * 128 + ks * 32);
      if (mode == 0) {
        umma_bf16(tm, ad, bd, idesc, ks != 0);
      } else {
        tmem_cp_128x256b(tm + 64 + ks * 8, ad);
        umma_ts(tm, tm + 64 + ks * 8, bd, idesc, ks != 0);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + h * 32, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 64 + h * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 64 * 4);
  static float h[128 * 64];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 1, 2, 9}) {
      cudaMemset(out, 0, sizeof(h));
      k<<<1, 128, 64 * 1024>>>(mode, shift, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0.f;
          for (int kk = 0; kk < 64; ++kk) ref += aval(m, kk) * bval(n + shift, kk);
          if (ref != h[m * 64 + n]) ++bad;
        }
      printf("%s B row shift %d: %s (%d of 8192 wrong)\n", mode ? "A in TMEM (tcgen05.cp + .ts mma)" : "A in shared memory          ", shift,
             bad ? "MISMATCH" : "ok", bad);
    }
  return 0;
}
