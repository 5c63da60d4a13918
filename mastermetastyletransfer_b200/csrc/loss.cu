// VGG-19 perceptual-loss side kernels (codes/loss.py): the Cin=3 first convolution, 2x2 max-pool, and the
// HBM-streaming reductions of the content / style loss.  The other twelve VGG convolutions run through the
// tensor-core implicit GEMM (gemm_tc.cu, zero padding + ReLU epilogue).  Activations are bf16 NHWC.
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

// ---------------------------------------------------------------- conv1_1: 3 -> 64, K = 27 (HBM-bound)
// One thread per output pixel, 64 accumulators in registers, weights [27][64] in shared memory.
__global__ void __launch_bounds__(128) conv3x3_first_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                            const float* __restrict__ bias, bf16* __restrict__ out, int B,
                                                            int H, int W, int relu) {
  __shared__ float ws[27 * 64];
  __shared__ float bs[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
    const int k = i / 64, n = i - k * 64;  // k = ci*9 + ky*3 + kx
    ws[i] = w[n * 27 + k];
  }
  if (threadIdx.x < 64) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W;
  if (pix >= total) return;
  const int b = (int)(pix / ((long long)H * W));
  const int rem = (int)(pix - (long long)b * H * W);
  const int y = rem / W, x = rem - y * W;
  float in[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        const bool ok = (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W;
        in[ci * 9 + ky * 3 + kx] = ok ? img[(((long long)b * 3 + ci) * H + yy) * W + xx] : 0.f;
      }
  uint4* o4 = reinterpret_cast<uint4*>(out + pix * 64);
#pragma unroll
  for (int n0 = 0; n0 < 64; n0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bs[n0 + j];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(ws + k * 64 + n0);
      const float4 w1 = *reinterpret_cast<const float4*>(ws + k * 64 + n0 + 4);
      acc[0] = fmaf(in[k], w0.x, acc[0]); acc[1] = fmaf(in[k], w0.y, acc[1]);
      acc[2] = fmaf(in[k], w0.z, acc[2]); acc[3] = fmaf(in[k], w0.w, acc[3]);
      acc[4] = fmaf(in[k], w1.x, acc[4]); acc[5] = fmaf(in[k], w1.y, acc[5]);
      acc[6] = fmaf(in[k], w1.z, acc[6]); acc[7] = fmaf(in[k], w1.w, acc[7]);
    }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = acc[2 * j], c = acc[2 * j + 1];
      if (relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
      __nv_bfloat162 h2 = __floats2bfloat162_rn(a, c);
      pk[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    o4[n0 / 8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ---------------------------------------------------------------- 2x2 max-pool, bf16 NHWC, 8 channels per thread
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8, int Ho,
                                                         int Wo, int C8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int c8 = (int)(i % C8);
  long long p = i / C8;
  const int xo = (int)(p % Wo); p /= Wo;
  const int yo = (int)(p % Ho);
  const long long b = p / Ho;
  const int W = 2 * Wo;
  const uint4* src = reinterpret_cast<const uint4*>(x);
  const long long base = ((b * 2 * Ho + 2 * yo) * W + 2 * xo) * C8 + c8;
  const uint4 v[4] = {src[base], src[base + C8], src[base + (long long)W * C8], src[base + (long long)W * C8 + C8]};
  uint4 r;
  uint32_t* rp = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 m = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const uint32_t*>(&v[0]) + j);
#pragma unroll
    for (int q = 1; q < 4; ++q) m = __hmax2(m, *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const uint32_t*>(&v[q]) + j));
    rp[j] = *reinterpret_cast<uint32_t*>(&m);
  }
  reinterpret_cast<uint4*>(y)[i] = r;
}

// ---------------------------------------------------------------- per-(image, channel) mean / biased variance of a bf16 tap
// x [B,T,C] bf16.  CTA = (T slab, image b), ALL channels: thread = (8-channel chunk c8 of C/8, row lane rl of 256/(C/8)), so a
// warp reads 512 contiguous bytes and consecutive passes walk consecutive rows.  Sums are taken relative to a per-channel shift
// (the channel's first element) so the sum of squares stays well conditioned; every slab writes its partial (sum, sum of
// squares) to the caller's scratch [slabs][B*C][2] and tap_stats_finalize adds the slabs in a fixed order (deterministic) into
// mean and biased variance.  The slab count is chosen by the host so that the whole grid is ONE resident wave of equal slabs
// (tap_stats_slabs): a grid of 2.3 waves ran the relu1_1 tap at 65 % of the HBM rate, the missing third being the tail wave.
// Arithmetic: sm_100's mixed-precision scalar ops (FHADD.BF16 takes a bf16 half of a 32-bit register straight into an fp32
// subtract), 3 instructions per element instead of 4.5 with an unpack.
__device__ __forceinline__ void bf16x2_minus_f32(uint32_t w, float klo, float khi, float& alo, float& ahi) {
  asm("{.reg .b16 lo, hi;\n mov.b32 {lo, hi}, %2;\n sub.rn.f32.bf16 %0, lo, %3;\n sub.rn.f32.bf16 %1, hi, %4;}"
      : "=f"(alo), "=f"(ahi) : "r"(w), "f"(klo), "f"(khi));
}

__global__ void __launch_bounds__(256, 5) tap_stats_kernel(const bf16* __restrict__ x, float* __restrict__ part, int T, int C,
                                                        int slabs, int BC) {
  extern __shared__ float red[];  // [2][RL][C + 1]
  const int C8 = C >> 3, RL = 256 / C8;
  const int c8 = threadIdx.x % C8, rl = threadIdx.x / C8;
  const int b = blockIdx.y;
  const int t0 = (int)((long long)blockIdx.x * T / slabs), t1 = (int)((long long)(blockIdx.x + 1) * T / slabs);
  const bf16* xb = x + (long long)b * T * C + c8 * 8;
  float k[8], s[8], q[8];
  {
    const uint4 f = *reinterpret_cast<const uint4*>(xb);
    const uint32_t w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { k[2 * e] = __uint_as_float(w[e] << 16); k[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u); }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  auto accumulate = [&](const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a, d;
      bf16x2_minus_f32(w[e], k[2 * e], k[2 * e + 1], a, d);
      s[2 * e] += a; q[2 * e] = fmaf(a, a, q[2 * e]);
      s[2 * e + 1] += d; q[2 * e + 1] = fmaf(d, d, q[2 * e + 1]);
    }
  };
  int t = t0 + rl;
  for (; t + 3 * RL < t1; t += 4 * RL) {  // four 16-byte loads in flight per thread: the kernel is a pure HBM stream
    const uint4 v0 = *reinterpret_cast<const uint4*>(xb + (long long)t * C);
    const uint4 v1 = *reinterpret_cast<const uint4*>(xb + (long long)(t + RL) * C);
    const uint4 v2 = *reinterpret_cast<const uint4*>(xb + (long long)(t + 2 * RL) * C);
    const uint4 v3 = *reinterpret_cast<const uint4*>(xb + (long long)(t + 3 * RL) * C);
    accumulate(v0); accumulate(v1); accumulate(v2); accumulate(v3);
  }
  for (; t < t1; t += RL) accumulate(*reinterpret_cast<const uint4*>(xb + (long long)t * C));
  const int pitch = C + 1;
#pragma unroll
  for (int e = 0; e < 8; ++e) { red[rl * pitch + c8 * 8 + e] = s[e]; red[(RL + rl) * pitch + c8 * 8 + e] = q[e]; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    const int which = i / C, cc = i - which * C;
    float a = 0.f;
    for (int r = 0; r < RL; ++r) a += red[(which * RL + r) * pitch + cc];
    part[((long long)blockIdx.x * BC + (long long)b * C + cc) * 2 + which] = a;
  }
}

// CTA = 32 consecutive (image, channel) entries x 8 slab lanes: lane l adds slabs l, l+8, ... (coalesced 256-byte loads, the
// loads of one lane independent), then the eight lane sums are added in lane order: a fixed order, so the result is
// deterministic, and a chain of slabs/8 loads instead of slabs.
__global__ void __launch_bounds__(256) tap_stats_finalize_kernel(const bf16* __restrict__ x, const float* __restrict__ part,
                                                                 float* __restrict__ mean, float* __restrict__ var, int B, int T,
                                                                 int C, int slabs) {
  __shared__ float2 acc[8][32];
  const int il = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + il;
  float s = 0.f, q = 0.f;
  if (i < B * C)
    for (int sl = lane; sl < slabs; sl += 8) {
      const float2 v = reinterpret_cast<const float2*>(part)[(long long)sl * B * C + i];
      s += v.x;
      q += v.y;
    }
  acc[lane][il] = make_float2(s, q);
  __syncthreads();
  if (lane != 0 || i >= B * C) return;
#pragma unroll
  for (int l = 1; l < 8; ++l) { s += acc[l][il].x; q += acc[l][il].y; }
  const int b = i / C, c = i - b * C;
  const float k = __bfloat162float(x[(long long)b * T * C + c]);
  const float m = s / (float)T;
  mean[i] = m + k;
  var[i] = fmaxf(q / (float)T - m * m, 0.f);
}

static size_t tap_stats_smem(int C) { return (size_t)2 * (256 / (C / 8)) * (C + 1) * sizeof(float); }

// Slabs per image: the grid (slabs x B) is one resident wave -- as many CTAs as the device holds at once (occupancy of
// tap_stats_kernel x SM count; 5 x 148 assumed when no device answers, as on a build host), rounded down to a multiple of B --
// of equal slabs, none shorter than 16 rows.
static int tap_stats_slabs(int B, int T, int C) {
  int per_sm = 0, sms = 0, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tap_stats_kernel, 256, tap_stats_smem(C)) != cudaSuccess ||
      per_sm < 1 || sms < 1) {
    (void)cudaGetLastError();
    per_sm = 5;
    sms = 148;
  }
  long long slabs = (long long)per_sm * sms / B;
  if (slabs > T / 16) slabs = T / 16;
  return (int)(slabs < 1 ? 1 : slabs);
}

// ---------------------------------------------------------------- content term: sum |IN(Fc) - IN(Fcs)| (or squared)
// fc, fo [B,T,C] bf16; stats [B,C].  Grid-stride over 8-channel vectors; deterministic per-CTA partial sums.
__global__ void __launch_bounds__(256) content_term_kernel(const bf16* __restrict__ fc, const bf16* __restrict__ fo,
                                                           const float* __restrict__ mean_c, const float* __restrict__ var_c,
                                                           const float* __restrict__ mean_o, const float* __restrict__ var_o,
                                                           int B, int T, int C8, int squared, float* __restrict__ partials) {
  // CTA = (image b, row slab); a thread keeps ONE 8-channel chunk for all of its rows, so the four statistics of its channels
  // are loaded (and the two rsqrt taken) once -- the row loop is then a pure 2 x 16-byte-per-thread HBM stream.
  __shared__ float red[8];
  const int nslab = gridDim.x / B;
  const int b = blockIdx.x % B, slab = blockIdx.x / B;
  float acc = 0.f;
  if (slab < nslab) {
    const int rpc = 256 / C8;  // rows covered by one pass of the CTA (C8 = 16 / 32 / 64 chunks per row)
    const int j = threadIdx.x % C8, r0 = slab * rpc + threadIdx.x / C8;
    const long long sb = (long long)b * C8 * 8 + j * 8;
    float mc[8], rc[8], mo[8], ro[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mc[e] = mean_c[sb + e]; rc[e] = rsqrtf(var_c[sb + e] + 1e-5f);
      mo[e] = mean_o[sb + e]; ro[e] = rsqrtf(var_o[sb + e] + 1e-5f);
    }
    const uint4* pc = reinterpret_cast<const uint4*>(fc) + (long long)b * T * C8 + j;
    const uint4* po = reinterpret_cast<const uint4*>(fo) + (long long)b * T * C8 + j;
    const int step = nslab * rpc;
    auto term = [&](const uint4& a, const uint4& o) {
      const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
      const uint32_t* op = reinterpret_cast<const uint32_t*>(&o);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a0 = __uint_as_float(ap[q] << 16), a1 = __uint_as_float(ap[q] & 0xFFFF0000u);
        const float o0 = __uint_as_float(op[q] << 16), o1 = __uint_as_float(op[q] & 0xFFFF0000u);
        const float d0 = (a0 - mc[2 * q]) * rc[2 * q] - (o0 - mo[2 * q]) * ro[2 * q];
        const float d1 = (a1 - mc[2 * q + 1]) * rc[2 * q + 1] - (o1 - mo[2 * q + 1]) * ro[2 * q + 1];
        acc += squared ? d0 * d0 + d1 * d1 : fabsf(d0) + fabsf(d1);
      }
    };
    int r = r0;
    for (; r + step < T; r += 2 * step) {  // two rows (four 16-byte loads) in flight per thread
      const uint4 a0 = pc[(long long)r * C8], o0 = po[(long long)r * C8];
      const uint4 a1 = pc[(long long)(r + step) * C8], o1 = po[(long long)(r + step) * C8];
      term(a0, o0);
      term(a1, o1);
    }
    for (; r < T; r += step) term(pc[(long long)r * C8], po[(long long)r * C8]);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    partials[blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------- finalize: content + lambda * style
// One CTA.  content = sum_taps partial_sum / (B*T*C);  style = sum_taps [mean|mu_s-mu_o| + mean|std_s-std_o|]
// with torch's unbiased std (codes/loss.py:122-130).  out = {total, content, style}.
__global__ void __launch_bounds__(256) loss_finalize_kernel(const MstLossTaps taps, float lambda, int squared_style, float* __restrict__ out) {
  __shared__ double red[2][8];
  double content = 0.0, style = 0.0;
  for (int t = 0; t < taps.n_taps; ++t) {
    const MstLossTap& tp = taps.tap[t];
    double cs = 0.0;
    for (int i = threadIdx.x; i < tp.n_partials; i += blockDim.x) cs += (double)tp.partials[i];
    content += cs / ((double)tp.B * tp.T * tp.C);
    const float ub = tp.T > 1 ? (float)tp.T / (float)(tp.T - 1) : 1.f;
    double ss = 0.0;
    for (int i = threadIdx.x; i < tp.B * tp.C; i += blockDim.x) {
      const float dm = tp.mean_s[i] - tp.mean_o[i];
      const float ds = sqrtf(tp.var_s[i] * ub) - sqrtf(tp.var_o[i] * ub);
      ss += squared_style ? (double)(dm * dm + ds * ds) : (double)(fabsf(dm) + fabsf(ds));
    }
    style += ss / ((double)tp.B * tp.C);
  }
  // block reduce (fixed order -> deterministic)
  for (int o = 16; o > 0; o >>= 1) {
    content += __shfl_xor_sync(0xffffffffu, content, o);
    style += __shfl_xor_sync(0xffffffffu, style, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = content; red[1][threadIdx.x >> 5] = style; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0.0, s = 0.0;
    for (int w = 0; w < 8; ++w) { c += red[0][w]; s += red[1][w]; }
    out[0] = (float)(c + (double)lambda * s);
    out[1] = (float)c;
    out[2] = (float)s;
  }
}

// ---------------------------------------------------------------- BatchNorm2d + ReLU of the VGG-19-BN loss variant
// y = max(0, (x - mean[c]) * gamma[c] * inv_std[c] + beta[c])  (codes/loss.py:41-63) for a token-major [M, C] activation; y is bf16.
// x is the convolution's fp32 output (x32 != NULL) -- a channel whose mean is large against its standard deviation would lose its
// signal if it were rounded to bf16 BEFORE the normalisation -- or, for the first layer, the bf16 tensor y itself (in place).
// Train mode (what the reference's scripts run): mean / inv_std are the statistics of the batch (mst_instnorm_stats with B = 1,
// or mst_tap_stats for bf16 input: then var_is_rstd = 0 and the kernel takes 1/sqrt(var + eps)); eval mode: the running statistics.
// Thread = eight channels of a row; the per-channel scale / shift live in registers.
__global__ void __launch_bounds__(256) bn_relu_kernel(const float* __restrict__ x32, bf16* __restrict__ y, const float* __restrict__ mean,
                                                      const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float eps, long long M, int C, int relu, int var_is_rstd) {
  const int C8 = C >> 3;
  const int c8 = threadIdx.x % C8, rl = threadIdx.x / C8, RL = 256 / C8;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c8 * 8 + e;
    sc[e] = gamma[c] * (var_is_rstd ? var[c] : 1.0f / sqrtf(var[c] + eps));
    sh[e] = beta[c] - mean[c] * sc[e];
  }
  for (long long r = (long long)blockIdx.x * RL + rl; r < M; r += (long long)gridDim.x * RL) {
    float f[8];
    uint4* py = reinterpret_cast<uint4*>(y + r * C + c8 * 8);
    if (x32) {
      const float4 a = *reinterpret_cast<const float4*>(x32 + r * C + c8 * 8), b = *reinterpret_cast<const float4*>(x32 + r * C + c8 * 8 + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
      const uint4 v = *py;
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { f[2 * e] = __uint_as_float(w[e] << 16); f[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u); }
    }
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a = fmaf(f[2 * e], sc[2 * e], sh[2 * e]), d = fmaf(f[2 * e + 1], sc[2 * e + 1], sh[2 * e + 1]);
      if (relu) { a = fmaxf(a, 0.f); d = fmaxf(d, 0.f); }
      __nv_bfloat162 h = __floats2bfloat162_rn(a, d);
      ow[e] = *reinterpret_cast<uint32_t*>(&h);
    }
    *py = o;
  }
}

}  // namespace mst

using namespace mst;

extern "C" int mst_conv3x3_first(const float* img, const float* w, const float* b, mst_bf16* out, int B, int H, int W, int relu,
                                 void* stream) {
  if (!img || !w || !b || !out || B <= 0 || H <= 0 || W <= 0) return MST_ERR_BAD_ARG;
  const long long total = (long long)B * H * W;
  conv3x3_first_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(img, w, b, reinterpret_cast<bf16*>(out), B, H, W, relu);
  return (int)cudaGetLastError();
}

extern "C" int mst_maxpool2x2(const mst_bf16* x, mst_bf16* y, int B, int H, int W, int C, void* stream) {
  if (!x || !y || B <= 0 || H <= 0 || W <= 0 || C <= 0) return MST_ERR_BAD_ARG;
  if ((H | W) & 1 || C % 8) return MST_ERR_UNSUPPORTED;
  const long long n8 = (long long)B * (H / 2) * (W / 2) * (C / 8);
  maxpool2x2_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y),
                                                                                 n8, H / 2, W / 2, C / 8);
  return (int)cudaGetLastError();
}

extern "C" int mst_bn_relu(const float* x32, mst_bf16* y, const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                           size_t M, int C, int relu, int var_is_rstd, void* stream) {
  if (!y || !mean || !var || !gamma || !beta || M == 0 || C <= 0 || C % 8 != 0 || 256 % (C / 8) != 0) return MST_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x32)) & 15) return MST_ERR_BAD_ARG;
  const int RL = 256 / (C / 8);
  long long blocks = ((long long)M + RL - 1) / RL;
  if (blocks > 148 * 16) blocks = 148 * 16;
  bn_relu_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x32, reinterpret_cast<bf16*>(y), mean, var, gamma, beta, eps, (long long)M, C, relu,
                                                                     var_is_rstd);
  return (int)cudaGetLastError();
}

extern "C" size_t mst_tap_stats_scratch_floats(int B, int T, int C) {
  if (B <= 0 || T <= 0 || C <= 0) return 0;
  return (size_t)tap_stats_slabs(B, T, C) * B * C * 2;
}

extern "C" int mst_tap_stats(const mst_bf16* x, float* mean, float* var, int B, int T, int C, float* scratch, size_t scratch_floats,
                             void* stream) {
  if (!x || !mean || !var || !scratch || B <= 0 || T <= 0 || C <= 0) return MST_ERR_BAD_ARG;
  if (C % 64 || 256 % (C / 8) != 0 || C > 2048) return MST_ERR_UNSUPPORTED;  // C/8 chunks per row must tile the 256 threads
  if (scratch_floats < mst_tap_stats_scratch_floats(B, T, C)) return MST_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int slabs = tap_stats_slabs(B, T, C);
  dim3 grid((unsigned)slabs, (unsigned)B, 1);
  tap_stats_kernel<<<grid, 256, tap_stats_smem(C), st>>>(reinterpret_cast<const bf16*>(x), scratch, T, C, slabs, B * C);
  tap_stats_finalize_kernel<<<(B * C + 31) / 32, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), scratch, mean, var, B, T, C, slabs);
  return (int)cudaGetLastError();
}

extern "C" int mst_content_term(const mst_bf16* fc, const mst_bf16* fo, const float* mean_c, const float* var_c, const float* mean_o,
                                const float* var_o, int B, int T, int C, int squared, float* partials, int n_partials, void* stream) {
  if (!fc || !fo || !mean_c || !var_c || !mean_o || !var_o || !partials || B <= 0 || T <= 0 || C <= 0 || n_partials <= 0) return MST_ERR_BAD_ARG;
  if (C % 8 || 256 % (C / 8) != 0) return MST_ERR_UNSUPPORTED;  // C in {8, ..., 2048} with C/8 a power of two <= 256
  if (n_partials < B) return MST_ERR_BAD_ARG;                 // one CTA per (image, row slab): at least one slab per image
  content_term_kernel<<<n_partials, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const bf16*>(fc), reinterpret_cast<const bf16*>(fo), mean_c, var_c,
                                                                    mean_o, var_o, B, T, C / 8, squared, partials);
  return (int)cudaGetLastError();
}

extern "C" int mst_loss_finalize(const MstLossTaps* taps, float lambda, int squared_style, float* out3, void* stream) {
  if (!taps || !out3 || taps->n_taps <= 0 || taps->n_taps > 4) return MST_ERR_BAD_ARG;
  for (int t = 0; t < taps->n_taps; ++t) {
    const MstLossTap& tp = taps->tap[t];
    if (!tp.partials || !tp.mean_s || !tp.var_s || !tp.mean_o || !tp.var_o || tp.B <= 0 || tp.T <= 0 || tp.C <= 0) return MST_ERR_BAD_ARG;
  }
  loss_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*taps, lambda, squared_style, out3);
  return (int)cudaGetLastError();
}
