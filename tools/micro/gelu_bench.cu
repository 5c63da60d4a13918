// micro-benchmark: cost of GELU variants on registers (ALU-bound loop)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float gelu_erff(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678f)); }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float gelu_as(float x) {
  const float z = x * 0.70710678f, az = fabsf(z);
  const float t = rcp_approx(fmaf(0.3275911f, az, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f); p = fmaf(p, t, -0.284496736f); p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(-az * az * 1.4426950408889634f);
  const float erfabs = fmaf(-p * t, e, 1.0f);
  const float er = copysignf(erfabs, z);
  return 0.5f * x * (1.0f + er);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  float u = 0.7978845608f * fmaf(0.044715f * x, x * x, x), t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}
template <int V> __global__ void k(const float* in, float* out, int iters) {
  float acc = 0.f; float x = in[threadIdx.x + blockIdx.x * blockDim.x];
  for (int i = 0; i < iters; ++i) {
    float y = V == 0 ? gelu_erff(x) : (V == 1 ? gelu_as(x) : gelu_tanh(x));
    acc += y; x += 0.001f;
  }
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void errk(float* maxerr) {
  float m1 = 0, m2 = 0;
  for (int i = threadIdx.x; i < 1600000; i += blockDim.x) {
    float x = -8.f + i * 1e-5f; float r = gelu_erff(x);
    m1 = fmaxf(m1, fabsf(gelu_as(x) - r)); m2 = fmaxf(m2, fabsf(gelu_tanh(x) - r));
  }
  atomicMax((int*)&maxerr[0], __float_as_int(m1)); atomicMax((int*)&maxerr[1], __float_as_int(m2));
}
int main() {
  const int n = 148 * 8 * 256; float *in, *out, *me; cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&me, 8); cudaMemset(in, 0, n * 4); cudaMemset(me, 0, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int v = 0; v < 3; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      if (v == 0) k<0><<<148 * 8, 256>>>(in, out, 4096); else if (v == 1) k<1><<<148 * 8, 256>>>(in, out, 4096); else k<2><<<148 * 8, 256>>>(in, out, 4096);
      cudaEventRecord(b); cudaEventSynchronize(b);
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("variant %d: %.3f ms  -> %.2f G gelu/s\n", v, ms, (double)n * 4096 / ms / 1e6);
  }
  errk<<<1, 1024>>>(me); float h[2]; cudaMemcpy(h, me, 8, cudaMemcpyDeviceToHost);
  printf("max abs err vs erff-GELU on [-8,8]: A&S %.3e  tanh %.3e\n", h[0], h[1]);
  return 0;
}
