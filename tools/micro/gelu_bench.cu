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
__device__ __forceinline__ float gelu_tanh_fit(float x) {  // degree-2 fit of atanh(erf)/x, x^2 clamped (gelu_fit.py)
  const float x2 = fminf(x * x, 36.0f);
  float q = fmaf(-0.0003519023928f, x2, 0.03700801998f), t;
  q = fmaf(q, x2, 0.7975052754f);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * q));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_exp2(float x) {  // the shipped form (common.cuh gelu_erf)
  const float a = fabsf(x);
  float r = fmaf(0.0004881656787f, a, -0.007198856212f);
  r = fmaf(r, a, 0.05214627460f); r = fmaf(r, a, 0.4595968127f); r = fmaf(r, a, 1.151000023f);
  return fmaf(-fabsf(0.5f * x), ex2_approx(-a * r), fmaxf(x, 0.0f));
}
template <int V> __global__ void k(const float* in, float* out, int iters) {
  float acc = 0.f; float x = in[threadIdx.x + blockIdx.x * blockDim.x];
  for (int i = 0; i < iters; ++i) {
    float y = V == 0 ? gelu_erff(x) : V == 1 ? gelu_as(x) : V == 2 ? gelu_tanh(x) : V == 3 ? gelu_tanh_fit(x) : gelu_exp2(x);
    acc += y; x += 0.001f;
  }
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
__global__ void errk(float* maxerr) {
  float m1 = 0, m2 = 0, m3 = 0, m4 = 0;
  for (int i = threadIdx.x; i < 1600000; i += blockDim.x) {
    float x = -8.f + i * 1e-5f; float r = gelu_erff(x);
    m1 = fmaxf(m1, fabsf(gelu_as(x) - r)); m2 = fmaxf(m2, fabsf(gelu_tanh(x) - r));
    m3 = fmaxf(m3, fabsf(gelu_tanh_fit(x) - r)); m4 = fmaxf(m4, fabsf(gelu_exp2(x) - r));
  }
  atomicMax((int*)&maxerr[0], __float_as_int(m1)); atomicMax((int*)&maxerr[1], __float_as_int(m2));
  atomicMax((int*)&maxerr[2], __float_as_int(m3)); atomicMax((int*)&maxerr[3], __float_as_int(m4));
}
int main() {
  const int n = 148 * 8 * 256; float *in, *out, *me; cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&me, 16); cudaMemset(in, 0, n * 4); cudaMemset(me, 0, 16);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int v = 0; v < 5; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      if (v == 0) k<0><<<148 * 8, 256>>>(in, out, 4096); else if (v == 1) k<1><<<148 * 8, 256>>>(in, out, 4096); else if (v == 2) k<2><<<148 * 8, 256>>>(in, out, 4096); else if (v == 3) k<3><<<148 * 8, 256>>>(in, out, 4096); else k<4><<<148 * 8, 256>>>(in, out, 4096);
      cudaEventRecord(b); cudaEventSynchronize(b);
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("variant %d: %.3f ms  -> %.2f G gelu/s\n", v, ms, (double)n * 4096 / ms / 1e6);
  }
  errk<<<1, 1024>>>(me); float h[4]; cudaMemcpy(h, me, 16, cudaMemcpyDeviceToHost);
  printf("max abs err vs erff-GELU on [-8,8]: A&S %.3e  tanh %.3e  tanh-fit %.3e  exp2-form %.3e\n", h[0], h[1], h[2], h[3]);
  return 0;
}
