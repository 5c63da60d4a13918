// InstanceNorm statistics + application in ONE kernel (codes/style_transformer.py:1056-1057, :468, :520-530: nn.InstanceNorm2d over
// the tokens of an image, applied once or twice, optionally affine).
//
// x [B,T,C] fp32 -> mean / rstd [B,C] (as mst_instnorm_stats_affine) and y16 [B,T,C] bf16 = (x - mean) * rstd + beta (as
// mst_instnorm_apply_affine).  The two-kernel sequence reads x twice and pays two launch ramps on a 33 MB tensor (both kernels were
// latency-bound: 35 % / 52 % of the HBM rate).  Here a CTA owns one (image, 32-channel group) slice -- T rows of 128 bytes, 128 KB at
// T = 1024 -- and fetches ALL of it with a handful of TMA tensor copies (box [32 channels x <= 256 rows], every byte of the slice in
// flight at once), then computes the statistics from shared memory with the SAME summation order as instnorm_stats_kernel
// (norm_misc.cu: shifted one-pass sums, 8 partial sums per thread, 16-way tree) and normalises out of shared memory with the same
// fused multiply-add as instnorm_apply_kernel: the results are bit-identical to the two-kernel sequence, x is read once.
// T * 128 B must fit in shared memory (T <= 1600) and T must be a multiple of a box height <= 256; mst_instnorm falls back to the two
// kernels otherwise.
#include "../../include/mst_b200.h"
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace mst {

constexpr int INF_WARPS = 16;
constexpr int INF_UNROLL = 8;
constexpr int INF_MAX_T = 1600;

__global__ void __launch_bounds__(INF_WARPS * 32, 1) instnorm_fused_kernel(const __grid_constant__ CUtensorMap tm, float* __restrict__ mean,
                                                                           float* __restrict__ rstd, bf16* __restrict__ y16, int T, int C,
                                                                           int box_rows, int twice, int n_pad,
                                                                           const float* __restrict__ pad_val, float* __restrict__ pad_norm,
                                                                           const float* __restrict__ gamma, const float* __restrict__ beta) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ float red[2][INF_WARPS][33];
  __shared__ __align__(16) float mrb[3][32];  // mean, rstd, beta of the CTA's 32 channels
  const uint32_t xs_addr = (smem_u32(smem_raw) + 127u) & ~127u;
  const float* xs = reinterpret_cast<const float*>(smem_raw + (xs_addr - smem_u32(smem_raw)));
  const int groups = C / 32;
  const int b = blockIdx.x / groups;
  const int cg = blockIdx.x - b * groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = cg * 32 + lane;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)T * 128u) : "memory");
    for (int r0 = 0; r0 < T; r0 += box_rows)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                       xs_addr + (uint32_t)r0 * 128u),
                   "l"(&tm), "r"(cg * 32), "r"(b * T + r0), "r"(smem_u32(&bar))
                   : "memory");
  }
  __syncthreads();  // barrier initialised before anyone polls it
  mbar_wait(smem_u32(&bar), 0);

  // ---- statistics: instnorm_stats_kernel's loop, reading shared memory ----
  const float K = xs[lane];
  float s[INF_UNROLL], q[INF_UNROLL];
#pragma unroll
  for (int u = 0; u < INF_UNROLL; ++u) s[u] = q[u] = 0.f;
  int t = warp;
  for (; t + (INF_UNROLL - 1) * INF_WARPS < T; t += INF_UNROLL * INF_WARPS) {
    float v[INF_UNROLL];
#pragma unroll
    for (int u = 0; u < INF_UNROLL; ++u) v[u] = xs[(t + u * INF_WARPS) * 32 + lane];
#pragma unroll
    for (int u = 0; u < INF_UNROLL; ++u) {
      const float d = v[u] - K;
      s[u] += d;
      q[u] = fmaf(d, d, q[u]);
    }
  }
  for (; t < T; t += INF_WARPS) {
    const float d = xs[t * 32 + lane] - K;
    s[0] += d;
    q[0] = fmaf(d, d, q[0]);
  }
  red[0][warp][lane] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  red[1][warp][lane] = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
  __syncthreads();
  if (warp == 0) {
    float S = 0.f, Q = 0.f;
#pragma unroll
    for (int w = 0; w < INF_WARPS; ++w) { S += red[0][w][lane]; Q += red[1][w][lane]; }
    const float pv = n_pad > 0 ? pad_val[c] : 0.f;
    const float n_tot = (float)(T + n_pad);
    const float dp = pv - K;
    S += (float)n_pad * dp;
    Q += (float)n_pad * dp * dp;
    const float ms = S / n_tot;
    const float var = fmaxf(Q / n_tot - ms * ms, 0.f);
    const float m = K + ms;
    float r = 1.0f / sqrtf(var + 1e-5f);
    const float g = gamma ? gamma[c] : 1.0f;
    if (twice) r *= g * g / sqrtf(g * g * var * r * r + 1e-5f);
    else r *= g;
    const float be = beta ? beta[c] : 0.f;
    mean[(long long)b * C + c] = m;
    rstd[(long long)b * C + c] = r;
    if (pad_norm) pad_norm[(long long)b * C + c] = fmaf(pv - m, r, be);
    mrb[0][lane] = m; mrb[1][lane] = r; mrb[2][lane] = be;
  }
  __syncthreads();

  // ---- application: instnorm_apply_kernel's arithmetic; 8 lanes per token (4 channels each), 64 tokens per pass ----
  const int c4 = threadIdx.x & 7;
  const float4 m4 = reinterpret_cast<const float4*>(mrb[0])[c4];
  const float4 r4 = reinterpret_cast<const float4*>(mrb[1])[c4];
  const float4 be4 = reinterpret_cast<const float4*>(mrb[2])[c4];
  bf16* yb = y16 + ((long long)b * T) * C + cg * 32 + c4 * 4;
  for (int tok = threadIdx.x >> 3; tok < T; tok += INF_WARPS * 4) {
    const float4 v = reinterpret_cast<const float4*>(xs)[tok * 8 + c4];
    const float4 o = make_float4(fmaf(v.x - m4.x, r4.x, be4.x), fmaf(v.y - m4.y, r4.y, be4.y), fmaf(v.z - m4.z, r4.z, be4.z),
                                 fmaf(v.w - m4.w, r4.w, be4.w));
    __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
    *reinterpret_cast<uint2*>(yb + (long long)tok * C) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

typedef CUresult (*InEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static InEncodeTiledFn in_tma_encoder() {
  static InEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<InEncodeTiledFn>(q);
  }
  return fn;
}

// handled = false: shape / alignment / driver does not fit the fused kernel, nothing was launched
static int instnorm_fused_try(const float* x, float* mean, float* rstd, bf16* y16, int B, int T, int C, int twice, int n_pad, const float* pad_val,
                              float* pad_norm, const float* gamma, const float* beta, cudaStream_t st, bool& handled) {
  handled = false;
  static int allow = -1;
  if (allow < 0) { const char* e = getenv("MST_INSTNORM_FUSED"); allow = e ? atoi(e) : 1; }  // 0: the two-kernel sequence (experiments)
  if (!allow || T > INF_MAX_T || C % 32 != 0 || (long long)B * T > 0x7fffffffLL) return 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y16) & 7)) return 0;
  int box_rows = 0;
  for (int r = 256; r >= 8; --r)
    if (T % r == 0) { box_rows = r; break; }
  if (box_rows == 0) return 0;
  InEncodeTiledFn enc = in_tma_encoder();
  if (!enc) return 0;
  alignas(64) CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)B * T};
  const cuuint64_t gstride[1] = {(cuuint64_t)C * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 0;
  const int smem = T * 128 + 128;
  static int smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(instnorm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  handled = true;
  instnorm_fused_kernel<<<B * (C / 32), INF_WARPS * 32, smem, st>>>(tm, mean, rstd, y16, T, C, box_rows, twice, n_pad, pad_val, pad_norm, gamma, beta);
  return (int)cudaGetLastError();
}

}  // namespace mst

using namespace mst;

extern "C" int mst_instnorm(const float* x, float* mean, float* rstd, mst_bf16* y16, int B, int T, int C, int twice, int n_pad,
                            const float* pad_val, float* pad_norm, const float* gamma, const float* beta, void* stream) {
  if (!x || !mean || !rstd || !y16 || B <= 0 || T <= 0 || C <= 0 || C % 32 != 0 || n_pad < 0 || (n_pad > 0 && !pad_val)) return MST_ERR_BAD_ARG;
  if (twice && n_pad > 0) return MST_ERR_UNSUPPORTED;
  if (beta && (reinterpret_cast<uintptr_t>(beta) & 15)) return MST_ERR_BAD_ARG;
  bool handled = false;
  const int rc = instnorm_fused_try(x, mean, rstd, reinterpret_cast<bf16*>(y16), B, T, C, twice, n_pad, pad_val, pad_norm, gamma, beta,
                                    (cudaStream_t)stream, handled);
  if (handled) return rc;
  const int rc1 = mst_instnorm_stats_affine(x, mean, rstd, B, T, C, twice, n_pad, pad_val, pad_norm, gamma, beta, stream);
  if (rc1 != 0) return rc1;
  return mst_instnorm_apply_affine(x, mean, rstd, beta, y16, nullptr, B, T, C, stream);
}

extern "C" int mst_instnorm_fused_supported(int T, int C) {
  if (T <= 0 || T > INF_MAX_T || C <= 0 || C % 32 != 0) return 0;
  for (int r = 256; r >= 8; --r)
    if (T % r == 0) return 1;
  return 0;
}
