// HBM-streaming kernels around the tensor-core path: LayerNorm (+ the Swin patch-merging gather),
// InstanceNorm statistics / apply, Swin patch embedding, weight packing, fp32 -> bf16 casts.
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

// ---------------------------------------------------------------- LayerNorm: one warp per row
// PER = C / 32 values per lane, lane-strided by float4 so global reads are 512 B coalesced per warp.
template <int C, bool MERGE>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ y, int rows,
                                                        int H, int W) {
  constexpr int V4 = C / 128;  // float4 per lane
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float4 v[V4];
  if (!MERGE) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)warp * C);
#pragma unroll
    for (int i = 0; i < V4; ++i) v[i] = xr[lane + 32 * i];
  } else {
    // tv _patch_merging_pad: out[b,y2,x2,:] = cat(x[2y2,2x2], x[2y2+1,2x2], x[2y2,2x2+1], x[2y2+1,2x2+1])
    constexpr int Cs = C / 4;  // source channels
    const int W2 = W >> 1, H2 = H >> 1;
    const int b = warp / (H2 * W2);
    const int rem = warp - b * (H2 * W2);
    const int y2 = rem / W2, x2 = rem - y2 * W2;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const int e = (lane + 32 * i) * 4;  // element index in the 4C row
      const int part = e / Cs, ch = e - part * Cs;
      const int dy = part & 1, dx = part >> 1;
      v[i] = *reinterpret_cast<const float4*>(x + (((long long)b * H + 2 * y2 + dy) * W + 2 * x2 + dx) * Cs + ch);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float a = v[i].x - mean, b2 = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b2 * b2 + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
  uint2* yr = reinterpret_cast<uint2*>(y + (long long)warp * C);
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
    const float4 bt = reinterpret_cast<const float4*>(beta)[lane + 32 * i];
    __nv_bfloat162 lo = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + bt.x, (v[i].y - mean) * rstd * g.y + bt.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + bt.z, (v[i].w - mean) * rstd * g.w + bt.w);
    yr[lane + 32 * i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// ---------------------------------------------------------------- InstanceNorm statistics
// x [B,T,C] fp32.  CTA = (b, 32-channel group); 16 warps stride over T, lane = channel (128 B rows), EIGHT rows in flight per
// warp (a CTA's slice is only T*128 B: the kernel is latency-bound, bytes in flight are what counts -- 16 KB per CTA).  ONE pass
// over HBM: shifted sums  sum(x - K), sum((x - K)^2)  with K = the channel's first row, so the cancellation of the textbook
// one-pass formula does not arise (|mean - K| is of the order of the standard deviation), fp32 accumulation over <= 8 partial
// sums per thread and a 16-way tree; biased variance as nn.InstanceNorm2d.
// Affine variant (decoder_use_instance_norm_with_affine, codes/style_transformer.py:982-984): gamma [C] folds into the scale the
// apply kernel multiplies with -- once: r*gamma; twice (the same affine module applied to its own output, :1056 then :468):
// gamma^2 * r / sqrt(gamma^2 * var * r^2 + eps) -- and beta [C] is added by the apply kernel (and to pad_norm).
constexpr int IN_WARPS = 16;
constexpr int IN_UNROLL = 8;
__global__ void __launch_bounds__(IN_WARPS * 32) instnorm_stats_kernel(const float* __restrict__ x, float* __restrict__ mean,
                                                                        float* __restrict__ rstd, int T, int C, int twice, int n_pad,
                                                                        const float* __restrict__ pad_val, float* __restrict__ pad_norm,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta) {
  __shared__ float red[2][IN_WARPS][33];
  const int groups = C / 32;
  const int b = blockIdx.x / groups;
  const int c = (blockIdx.x - b * groups) * 32 + (threadIdx.x & 31);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + (long long)b * T * C + c;
  const float K = xb[0];
  float s[IN_UNROLL], q[IN_UNROLL];
#pragma unroll
  for (int u = 0; u < IN_UNROLL; ++u) s[u] = q[u] = 0.f;
  int t = warp;
  for (; t + (IN_UNROLL - 1) * IN_WARPS < T; t += IN_UNROLL * IN_WARPS) {
    float v[IN_UNROLL];
#pragma unroll
    for (int u = 0; u < IN_UNROLL; ++u) v[u] = xb[(long long)(t + u * IN_WARPS) * C];
#pragma unroll
    for (int u = 0; u < IN_UNROLL; ++u) {
      const float d = v[u] - K;
      s[u] += d;
      q[u] = fmaf(d, d, q[u]);
    }
  }
  for (; t < T; t += IN_WARPS) {
    const float d = xb[(long long)t * C] - K;
    s[0] += d;
    q[0] = fmaf(d, d, q[0]);
  }
  red[0][warp][lane] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  red[1][warp][lane] = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
  __syncthreads();
  if (warp == 0) {
    float S = 0.f, Q = 0.f;
#pragma unroll
    for (int w = 0; w < IN_WARPS; ++w) { S += red[0][w][lane]; Q += red[1][w][lane]; }
    // n_pad extra tokens of value pad_val[c]: the zero-padded positions of a window-padded map after a Linear (= its bias)
    const float pv = n_pad > 0 ? pad_val[c] : 0.f;
    const float n_tot = (float)(T + n_pad);
    const float dp = pv - K;
    S += (float)n_pad * dp;
    Q += (float)n_pad * dp * dp;
    const float ms = S / n_tot;                       // mean - K
    const float var = fmaxf(Q / n_tot - ms * ms, 0.f);
    const float m = K + ms;
    float r = 1.0f / sqrtf(var + 1e-5f);
    const float g = gamma ? gamma[c] : 1.0f;
    if (twice) {
      // IN(IN(x)): the once-normalised tensor has mean beta and variance g^2*var*r^2, so the second pass multiplies by
      // g / sqrt(g^2*var*r^2 + eps) (codes/style_transformer.py:1056 then :468)
      r *= g * g / sqrtf(g * g * var * r * r + 1e-5f);
    } else {
      r *= g;
    }
    mean[(long long)b * C + c] = m;
    rstd[(long long)b * C + c] = r;
    if (pad_norm) pad_norm[(long long)b * C + c] = fmaf(pv - m, r, beta ? beta[c] : 0.f);  // (explicit: instnorm_fused_kernel must match)
  }
}

__global__ void __launch_bounds__(256) instnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ beta,
                                                             bf16* __restrict__ y16, float* __restrict__ y32, long long n4, int TC4, int C4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int b = (int)(i / TC4);
  const int c4 = (int)(i % C4);
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  const float4 m = reinterpret_cast<const float4*>(mean)[(long long)b * C4 + c4];
  const float4 r = reinterpret_cast<const float4*>(rstd)[(long long)b * C4 + c4];
  const float4 be = beta ? reinterpret_cast<const float4*>(beta)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 o = make_float4(fmaf(v.x - m.x, r.x, be.x), fmaf(v.y - m.y, r.y, be.y), fmaf(v.z - m.z, r.z, be.z), fmaf(v.w - m.w, r.w, be.w));
  if (y32) reinterpret_cast<float4*>(y32)[i] = o;
  if (y16) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
    reinterpret_cast<uint2*>(y16)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// Statistics of an image taken over (T, C) JOINTLY: the regular-MHA decoder variant hands [B, C, T] tensors to nn.InstanceNorm2d,
// which reads a 3-D input as ONE unbatched image and normalises over all of its elements (codes/style_transformer.py:1063-1119).
// One CTA per image; the scalar mean / rstd are written replicated to mean[b, :] / rstd[b, :] so that instnorm_apply applies them.
__global__ void __launch_bounds__(1024) jointnorm_stats_kernel(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ rstd,
                                                               long long n, int C) {
  __shared__ double red[2][32];
  const float* xb = x + (long long)blockIdx.x * n;
  const float K = xb[0];
  float s = 0.f, q = 0.f;
  double S = 0.0, Q = 0.0;
  int cnt = 0;
  for (long long i = (long long)threadIdx.x * 4; i < n; i += 4096) {  // n % 4 == 0 (C % 4 == 0)
    const float4 v = *reinterpret_cast<const float4*>(xb + i);
    const float d0 = v.x - K, d1 = v.y - K, d2 = v.z - K, d3 = v.w - K;
    s += (d0 + d1) + (d2 + d3);
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    if (++cnt == 64) { S += s; Q += q; s = q = 0.f; cnt = 0; }  // fp32 runs of 256 elements, fp64 across runs
  }
  S += s; Q += q;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { S += __shfl_xor_sync(0xffffffffu, S, o); Q += __shfl_xor_sync(0xffffffffu, Q, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = S; red[1][threadIdx.x >> 5] = Q; }
  __syncthreads();
  __shared__ float out2[2];
  if (threadIdx.x == 0) {
    double St = 0.0, Qt = 0.0;
    for (int w = 0; w < 32; ++w) { St += red[0][w]; Qt += red[1][w]; }
    const double ms = St / (double)n, var = Qt / (double)n - ms * ms;
    out2[0] = (float)((double)K + ms);
    out2[1] = (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + 1e-5));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    mean[(long long)blockIdx.x * C + c] = out2[0];
    rstd[(long long)blockIdx.x * C + c] = out2[1];
  }
}

// ---------------------------------------------------------------- row softmax (regular-MHA decoder variant)
// P[r, :] = softmax(scale * S[r, :]) for fp32 scores S [rows, n] -> bf16 probabilities (codes/style_transformer.py:1100-1106: one
// head over all T tokens of an image).  One warp per row, the row is read twice (max, then exp + sum) and the probabilities are
// written normalised; n % 4 == 0.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, bf16* __restrict__ P, int rows, int n, float scale) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* s4 = reinterpret_cast<const float4*>(S + (long long)row * n);
  const int n4 = n >> 2;
  const float sc = scale * 1.4426950408889634f;
  float mx = -INFINITY;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = s4[i];
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  mx = warp_max(mx) * sc;
  float sum = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = s4[i];
    sum += (exp2f(fmaf(v.x, sc, -mx)) + exp2f(fmaf(v.y, sc, -mx))) + (exp2f(fmaf(v.z, sc, -mx)) + exp2f(fmaf(v.w, sc, -mx)));
  }
  const float inv = 1.0f / warp_sum(sum);
  uint2* p2 = reinterpret_cast<uint2*>(P + (long long)row * n);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = s4[i];
    __nv_bfloat162 lo = __floats2bfloat162_rn(exp2f(fmaf(v.x, sc, -mx)) * inv, exp2f(fmaf(v.y, sc, -mx)) * inv);
    __nv_bfloat162 hi = __floats2bfloat162_rn(exp2f(fmaf(v.z, sc, -mx)) * inv, exp2f(fmaf(v.w, sc, -mx)) * inv);
    p2[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// ---------------------------------------------------------------- Swin patch embedding + LayerNorm(128)
// Persistent CTAs; each loops over groups of 64 consecutive tokens (flattened b, py, px).  The 12 (channel, row)
// image segments of a group are loaded coalesced into shared memory; then one warp per token.  Lane owns output
// channels lane + 32*o and keeps its 48 x 4 weights in REGISTERS (the kernel is otherwise bound by shared-memory
// weight reads: one LDS per FMA), so a token costs 12 broadcast LDS.128 + 192 FMA per lane.
constexpr int PE_TOK = 64;
__global__ void __launch_bounds__(256, 1) patch_embed_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ out, int B,
                                                             int S, int groups) {
  __shared__ float4 in_s[PE_TOK][13];  // [token][ci*4+ky] = 4 pixels of one patch row (+1 pad)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wr[48][4];
#pragma unroll
  for (int k = 0; k < 48; ++k)
#pragma unroll
    for (int o = 0; o < 4; ++o) wr[k][o] = w[(lane + 32 * o) * 48 + k];  // conv weight [128][3][4][4] -> k = ci*16 + ky*4 + kx
  float bs[4], gm[4], bt[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) { bs[o] = bias[lane + 32 * o]; gm[o] = gamma[lane + 32 * o]; bt[o] = beta[lane + 32 * o]; }
  const int P = S / 4;
  const long long total = (long long)B * P * P;
  // software pipeline: the next group's 3 float4 per thread are fetched into registers while this group computes
  auto fetch = [&](int grp, float4 (&v)[3]) {
    const long long first = (long long)grp * PE_TOK;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = threadIdx.x + j * 256;
      const int r = i / PE_TOK, tk = i - r * PE_TOK;  // consecutive threads -> consecutive tokens -> contiguous 16 B
      const long long tok = first + tk;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grp < groups && tok < total) {
        const int b = (int)(tok / (P * P));
        const int rem = (int)(tok - (long long)b * P * P);
        const int py = rem / P, px = rem - py * P;
        const int ci = r >> 2, ky = r & 3;
        v[j] = *reinterpret_cast<const float4*>(img + (((long long)b * 3 + ci) * S + py * 4 + ky) * S + px * 4);
      }
    }
  };
  float4 nxt[3];
  fetch(blockIdx.x, nxt);
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const long long first = (long long)grp * PE_TOK;
    __syncthreads();  // previous group's readers are done with in_s
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = threadIdx.x + j * 256;
      in_s[i % PE_TOK][i / PE_TOK] = nxt[j];
    }
    __syncthreads();
    fetch(grp + gridDim.x, nxt);
    for (int tk = warp; tk < PE_TOK; tk += 8) {
      const long long tok = first + tk;
      if (tok >= total) break;
      float acc[4] = {bs[0], bs[1], bs[2], bs[3]};
#pragma unroll
      for (int r = 0; r < 12; ++r) {
        const float4 iv = in_s[tk][r];  // broadcast read
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          acc[o] = fmaf(iv.x, wr[4 * r + 0][o], acc[o]);
          acc[o] = fmaf(iv.y, wr[4 * r + 1][o], acc[o]);
          acc[o] = fmaf(iv.z, wr[4 * r + 2][o], acc[o]);
          acc[o] = fmaf(iv.w, wr[4 * r + 3][o], acc[o]);
        }
      }
      const float mean = warp_sum(acc[0] + acc[1] + acc[2] + acc[3]) * (1.0f / 128.f);
      float q = 0.f;
#pragma unroll
      for (int o = 0; o < 4; ++o) q += (acc[o] - mean) * (acc[o] - mean);
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / 128.f) + 1e-5f);
#pragma unroll
      for (int o = 0; o < 4; ++o) out[tok * 128 + lane + 32 * o] = (acc[o] - mean) * rstd * gm[o] + bt[o];
    }
  }
}

// ---------------------------------------------------------------- Swin patch embedding on the tensor cores
// The 4x4/stride-4 conv is a [tokens x 48] . [48 x 128] GEMM: mma.sync m16n8k16 (bf16 in, fp32 accumulate), one warp per
// 16 consecutive tokens of a patch row.  k = ci*16 + ky*4 + kx, so one k-step is one input channel and a thread's A
// fragment is four float2 loads straight from the NCHW image (no staging).  The image is split hi + lo into two bf16
// operands (16 mantissa bits), the weights are rounded to bf16 like every other layer's and live in registers as B
// fragments.  LayerNorm(128) runs on the accumulator fragments (quad shuffles); with y16 != nullptr the first block's
// norm1 is applied as well and written as the bf16 operand of its QKV projection (saves one full LayerNorm pass).
MST_DEVINL uint32_t pe_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
MST_DEVINL void pe_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(256, 2) patch_embed_mma_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float* __restrict__ out,
                                                                 const float* __restrict__ gamma1, const float* __restrict__ beta1,
                                                                 bf16* __restrict__ y16, int S, int groups) {
  __shared__ float2 prm[5][64];  // bias, gamma, beta, gamma1, beta1 as column pairs
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    prm[0][i] = make_float2(bias[2 * i], bias[2 * i + 1]);
    prm[1][i] = make_float2(gamma[2 * i], gamma[2 * i + 1]);
    prm[2][i] = make_float2(beta[2 * i], beta[2 * i + 1]);
    prm[3][i] = y16 ? make_float2(gamma1[2 * i], gamma1[2 * i + 1]) : make_float2(0.f, 0.f);
    prm[4][i] = y16 ? make_float2(beta1[2 * i], beta1[2 * i + 1]) : make_float2(0.f, 0.f);
  }
  // B fragments (B[k][n] = W[n][ci*16 + k]) in shared memory, one uint2 per (ci, n-tile, lane): conflict-free 64-bit loads.
  // Keeping them in registers (96 per thread) capped the kernel at 8 warps per SM; it is latency-bound on the image loads.
  __shared__ uint2 bw_s[3 * 16][32];
  for (int i = threadIdx.x; i < 3 * 16 * 32; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) & 15, ci = i >> 9;
    const float* wp = w + (nt * 8 + (ln >> 2)) * 48 + ci * 16 + 2 * (ln & 3);
    bw_s[ci * 16 + nt][ln] = make_uint2(pe_pack(wp[0], wp[1]), pe_pack(wp[8], wp[9]));
  }
  __syncthreads();
  const int P = S >> 2;
  const int gpr = P >> 4;  // 16-token groups per patch row
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  // element offset (inside the image tensor) of this thread's first float2 of group gi: channel 0, patch row ky = tg>>1
  auto src_of = [&](int gi) -> long long {
    const int rowi = gi / gpr;          // b * P + py
    const int px0 = (gi - rowi * gpr) << 4;
    const int b = rowi / P, py = rowi - b * P;
    return ((long long)b * 3 * S + py * 4 + (tg >> 1)) * S + (px0 + g) * 4 + (tg & 1) * 2;
  };
  const long long plane = (long long)S * S;
  float2 nx[12];
  auto fetch = [&](int gi) {
    if (gi < groups) {
      const float* p0 = img + src_of(gi);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* pc = p0 + ci * plane;
        nx[ci * 4 + 0] = *reinterpret_cast<const float2*>(pc);                   // row g,   ky
        nx[ci * 4 + 1] = *reinterpret_cast<const float2*>(pc + 32);              // row g+8 (8 tokens = 32 pixels further)
        nx[ci * 4 + 2] = *reinterpret_cast<const float2*>(pc + 2 * S);           // row g,   ky + 2
        nx[ci * 4 + 3] = *reinterpret_cast<const float2*>(pc + 2 * S + 32);      // row g+8, ky + 2
      }
    }
  };
  fetch(wid);
  for (int gi = wid; gi < groups; gi += nwarps) {
    float2 cur[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) cur[i] = nx[i];
    fetch(gi + nwarps);
    float acc[16][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const float2 bb = prm[0][nt * 4 + tg];
      acc[nt][0] = bb.x; acc[nt][1] = bb.y; acc[nt][2] = bb.x; acc[nt][3] = bb.y;
    }
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = cur[ci * 4 + i];
        const __nv_bfloat16 hx = __float2bfloat16_rn(v.x), hy = __float2bfloat16_rn(v.y);
        ah[i] = pe_pack(__bfloat162float(hx), __bfloat162float(hy));
        al[i] = pe_pack(v.x - __bfloat162float(hx), v.y - __bfloat162float(hy));
      }
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const uint2 bq = bw_s[ci * 16 + nt][lane];
        pe_mma(acc[nt], ah, bq.x, bq.y);
        pe_mma(acc[nt], al, bq.x, bq.y);
      }
    }
    const long long tok0 = (long long)gi * 16 + g;  // rows g and g+8 of the group
    // ---- LayerNorm(128) of rows g / g+8 (fragment columns nt*8 + 2tg, +1; quad = the four tg lanes) ----
    auto layer_norm = [&](int gsel, int bsel) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) { s0 += acc[nt][0] + acc[nt][1]; s1 += acc[nt][2] + acc[nt][3]; }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float m0 = s0 * (1.f / 128.f), m1 = s1 * (1.f / 128.f);
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const float a = acc[nt][0] - m0, b = acc[nt][1] - m0, c = acc[nt][2] - m1, d = acc[nt][3] - m1;
        q0 += a * a + b * b; q1 += c * c + d * d;
      }
      q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
      const float r0 = rsqrtf(q0 * (1.f / 128.f) + 1e-5f), r1 = rsqrtf(q1 * (1.f / 128.f) + 1e-5f);
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const float2 gm = prm[gsel][nt * 4 + tg], bt = prm[bsel][nt * 4 + tg];
        acc[nt][0] = (acc[nt][0] - m0) * r0 * gm.x + bt.x; acc[nt][1] = (acc[nt][1] - m0) * r0 * gm.y + bt.y;
        acc[nt][2] = (acc[nt][2] - m1) * r1 * gm.x + bt.x; acc[nt][3] = (acc[nt][3] - m1) * r1 * gm.y + bt.y;
      }
    };
    layer_norm(1, 2);
    float* o0 = out + tok0 * 128 + 2 * tg;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      *reinterpret_cast<float2*>(o0 + nt * 8) = make_float2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<float2*>(o0 + 8 * 128 + nt * 8) = make_float2(acc[nt][2], acc[nt][3]);
    }
    if (y16) {
      layer_norm(3, 4);
      bf16* y0 = y16 + tok0 * 128 + 2 * tg;
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        *reinterpret_cast<uint32_t*>(y0 + nt * 8) = pe_pack(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<uint32_t*>(y0 + 8 * 128 + nt * 8) = pe_pack(acc[nt][2], acc[nt][3]);
      }
    }
  }
}

// ---------------------------------------------------------------- nearest x2 upsample, bf16 NHWC
// nn.Upsample(scale_factor=2, mode='nearest') (codes/decoder.py:27) for the one decoder conv whose input has >= 128 channels:
// the gathered implicit GEMM folds the upsample into its addresses, the TMA-fed one cannot (a tensor copy cannot repeat
// pixels), and materialising 33 MB is cheaper than the difference (105 -> 66 us).  One thread = 16 bytes of a source pixel.
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long n_chunks, int H, int W, int C8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_chunks) return;
  const int c = (int)(i % C8);
  const long long pix = i / C8;
  const int xs = (int)(pix % W);
  const long long t = pix / W;
  const int ys = (int)(t % H);
  const long long b = t / H;
  const uint4 v = x[i];
  uint4* o = y + (((b * 2 * H + 2 * ys) * (2LL * W)) + 2 * xs) * C8 + c;
  o[0] = v;
  o[C8] = v;
  o[2LL * W * C8] = v;
  o[2LL * W * C8 + C8] = v;
}

// ---------------------------------------------------------------- packing / casts
__global__ void cast_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  reinterpret_cast<uint2*>(y)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

// ---------------------------------------------------------------- uint8 image boundary (test_model.py:39-48,111-125,207)
// images_u8_to_nchw: uint8 [B,H,W,3] -> fp32 [B,3,H,W] = ((u8 / 255) - mean_c) / std_c, the exact fp32 operation order of
// torchvision's ToTensor (`.div(255)`) followed by Normalize (`.sub_(mean).div_(std)`); normalize == 0 stops after /255.
// images_nchw_to_u8: fp32 [B,3,H,W] -> uint8 [B,H,W,3] = trunc(clip(x * 255, 0, 255)) (`np.clip(x*255, 0, 255).astype(np.uint8)`).
// Four pixels per thread: 12 bytes of interleaved uint8 <-> one float4 per colour plane.  W % 4 == 0.
__global__ void __launch_bounds__(256) images_u8_to_nchw_kernel(const uint32_t* __restrict__ src, float* __restrict__ dst, long long quads,
                                                                long long plane, float m0, float m1, float m2, float s0, float s1,
                                                                float s2, int normalize) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= quads) return;
  const uint32_t w0 = src[3 * i], w1 = src[3 * i + 1], w2 = src[3 * i + 2];
  const uint32_t by[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24, w1 & 255u, (w1 >> 8) & 255u,
                           (w1 >> 16) & 255u, w1 >> 24, w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  const long long pix = 4 * i;              // first pixel of the quad, over [B, H*W]
  const long long b = pix / plane, r = pix - b * plane;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x = __fdiv_rn((float)by[3 * k + c], 255.0f);
      if (normalize) x = __fdiv_rn(__fsub_rn(x, mean[c]), sd[c]);
      v[k] = x;
    }
    *reinterpret_cast<float4*>(dst + (b * 3 + c) * plane + r) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

__global__ void __launch_bounds__(256) images_nchw_to_u8_kernel(const float* __restrict__ src, uint32_t* __restrict__ dst, long long quads,
                                                                long long plane) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= quads) return;
  const long long pix = 4 * i;
  const long long b = pix / plane, r = pix - b * plane;
  uint32_t by[12];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(src + (b * 3 + c) * plane + r);
    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) by[3 * k + c] = (uint32_t)fminf(fmaxf(__fmul_rn(x[k], 255.0f), 0.0f), 255.0f);  // NaN -> 0
  }
  dst[3 * i] = by[0] | (by[1] << 8) | (by[2] << 16) | (by[3] << 24);
  dst[3 * i + 1] = by[4] | (by[5] << 8) | (by[6] << 16) | (by[7] << 24);
  dst[3 * i + 2] = by[8] | (by[9] << 8) | (by[10] << 16) | (by[11] << 24);
}

}  // namespace mst

using namespace mst;

extern "C" int mst_layernorm(const float* x, const float* gamma, const float* beta, mst_bf16* y, int rows, int C, void* stream) {
  if (!x || !gamma || !beta || !y || rows <= 0) return MST_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (rows + 7) / 8;
  bf16* yy = reinterpret_cast<bf16*>(y);
  if (C == 128) layernorm_kernel<128, false><<<blocks, 256, 0, st>>>(x, gamma, beta, yy, rows, 0, 0);
  else if (C == 256) layernorm_kernel<256, false><<<blocks, 256, 0, st>>>(x, gamma, beta, yy, rows, 0, 0);
  else if (C == 512) layernorm_kernel<512, false><<<blocks, 256, 0, st>>>(x, gamma, beta, yy, rows, 0, 0);
  else return MST_ERR_UNSUPPORTED;
  return (int)cudaGetLastError();
}

extern "C" int mst_patch_merge_layernorm(const float* x, const float* gamma, const float* beta, mst_bf16* y, int B, int H,
                                         int W, int C, void* stream) {
  if (!x || !gamma || !beta || !y || B <= 0 || H <= 0 || W <= 0) return MST_ERR_BAD_ARG;
  if ((H | W) & 1) return MST_ERR_UNSUPPORTED;
  if (C != 128) return MST_ERR_UNSUPPORTED;
  const int rows = B * (H / 2) * (W / 2);
  layernorm_kernel<512, true><<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, reinterpret_cast<bf16*>(y), rows, H, W);
  return (int)cudaGetLastError();
}

extern "C" int mst_instnorm_stats(const float* x, float* mean, float* rstd, int B, int T, int C, int twice, void* stream) {
  if (!x || !mean || !rstd || B <= 0 || T <= 0 || C <= 0 || C % 32 != 0) return MST_ERR_BAD_ARG;
  instnorm_stats_kernel<<<B * (C / 32), IN_WARPS * 32, 0, (cudaStream_t)stream>>>(x, mean, rstd, T, C, twice, 0, nullptr, nullptr, nullptr, nullptr);
  return (int)cudaGetLastError();
}

extern "C" int mst_instnorm_stats_affine(const float* x, float* mean, float* rstd, int B, int T, int C, int twice, int n_pad,
                                         const float* pad_val, float* pad_norm, const float* gamma, const float* beta, void* stream) {
  if (!x || !mean || !rstd || B <= 0 || T <= 0 || C <= 0 || C % 32 != 0 || n_pad < 0 || (n_pad > 0 && !pad_val)) return MST_ERR_BAD_ARG;
  if (twice && n_pad > 0) return MST_ERR_UNSUPPORTED;
  instnorm_stats_kernel<<<B * (C / 32), IN_WARPS * 32, 0, (cudaStream_t)stream>>>(x, mean, rstd, T, C, twice, n_pad, pad_val, pad_norm, gamma, beta);
  return (int)cudaGetLastError();
}

extern "C" int mst_softmax_rows(const float* S, mst_bf16* P, int rows, int n, float scale, void* stream) {
  if (!S || !P || rows <= 0 || n <= 0 || n % 4 != 0 || (reinterpret_cast<uintptr_t>(S) & 15) || (reinterpret_cast<uintptr_t>(P) & 7)) return MST_ERR_BAD_ARG;
  if (scale <= 0.f) return MST_ERR_BAD_ARG;  // (the row maximum is scaled with the scores: a positive scale keeps it the maximum)
  softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(S, reinterpret_cast<bf16*>(P), rows, n, scale);
  return (int)cudaGetLastError();
}

extern "C" int mst_jointnorm_stats(const float* x, float* mean, float* rstd, int B, int T, int C, void* stream) {
  if (!x || !mean || !rstd || B <= 0 || T <= 0 || C <= 0 || C % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) return MST_ERR_BAD_ARG;
  jointnorm_stats_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(x, mean, rstd, (long long)T * C, C);
  return (int)cudaGetLastError();
}

extern "C" int mst_instnorm_stats_padded(const float* x, float* mean, float* rstd, int B, int T, int C, int n_pad, const float* pad_val,
                                         float* pad_norm, void* stream) {
  if (!x || !mean || !rstd || B <= 0 || T <= 0 || C <= 0 || C % 32 != 0 || n_pad < 0 || (n_pad > 0 && !pad_val)) return MST_ERR_BAD_ARG;
  instnorm_stats_kernel<<<B * (C / 32), IN_WARPS * 32, 0, (cudaStream_t)stream>>>(x, mean, rstd, T, C, 0, n_pad, pad_val, pad_norm, nullptr, nullptr);
  return (int)cudaGetLastError();
}

extern "C" int mst_instnorm_apply(const float* x, const float* mean, const float* rstd, mst_bf16* y16, float* y32, int B,
                                  int T, int C, void* stream) {
  if (!x || !mean || !rstd || (!y16 && !y32) || B <= 0 || T <= 0 || C <= 0 || C % 4 != 0) return MST_ERR_BAD_ARG;
  const long long n4 = (long long)B * T * C / 4;
  instnorm_apply_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, mean, rstd, nullptr, reinterpret_cast<bf16*>(y16), y32, n4, T * C / 4, C / 4);
  return (int)cudaGetLastError();
}

extern "C" int mst_instnorm_apply_affine(const float* x, const float* mean, const float* rstd, const float* beta, mst_bf16* y16, float* y32,
                                         int B, int T, int C, void* stream) {
  if (!x || !mean || !rstd || (!y16 && !y32) || B <= 0 || T <= 0 || C <= 0 || C % 4 != 0) return MST_ERR_BAD_ARG;
  if (beta && (reinterpret_cast<uintptr_t>(beta) & 15)) return MST_ERR_BAD_ARG;
  const long long n4 = (long long)B * T * C / 4;
  instnorm_apply_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, mean, rstd, beta, reinterpret_cast<bf16*>(y16), y32, n4, T * C / 4, C / 4);
  return (int)cudaGetLastError();
}

namespace mst {
int patch_embed_tc_try(const float* img, const float* w, const float* b, const float* gamma, const float* beta, float* x, const float* gamma1,
                       const float* beta1, bf16* y16, int B, int S, cudaStream_t st, bool& handled, const uint8_t* img8, const float* mean3,
                       const float* std3);  // patch_embed_tc.cu
}

extern "C" int mst_patch_embed(const float* img, const float* w, const float* b, const float* gamma, const float* beta,
                               float* x, int B, int S, void* stream) {
  return mst_patch_embed_ln(img, w, b, gamma, beta, x, nullptr, nullptr, nullptr, B, S, 0, stream);
}

extern "C" int mst_patch_embed_ln(const float* img, const float* w, const float* b, const float* gamma, const float* beta,
                                  float* x, const float* gamma1, const float* beta1, mst_bf16* y16, int B, int S, int exact,
                                  void* stream) {
  if (!img || !w || !b || !gamma || !beta || !x || B <= 0 || S <= 0 || S % 4 != 0) return MST_ERR_BAD_ARG;
  if (y16 && (!gamma1 || !beta1)) return MST_ERR_BAD_ARG;
  const long long total = (long long)B * (S / 4) * (S / 4);
  if (!exact) {  // tcgen05 kernel (patch_embed_tc.cu): 128-token tiles, both LayerNorms thread-local
    bool handled = false;
    const int rc = mst::patch_embed_tc_try(img, w, b, gamma, beta, x, gamma1, beta1, reinterpret_cast<bf16*>(y16), B, S, (cudaStream_t)stream, handled,
                                           nullptr, nullptr, nullptr);
    if (handled) return rc;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (S % 64 == 0 && !exact) {  // tensor-core path: 16-token groups inside a patch row
    const long long groups = total / 16;
    if (groups > 0x7fffffffLL) return MST_ERR_BAD_ARG;
    const long long want = (groups + 7) / 8;
    const unsigned grid = (unsigned)(want < 2LL * sms ? want : 2LL * sms);
    patch_embed_mma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, w, b, gamma, beta, x, gamma1, beta1,
                                                                   reinterpret_cast<bf16*>(y16), S, (int)groups);
    return (int)cudaGetLastError();
  }
  // fp32 SIMT kernel (any S % 4 == 0); a fused second LayerNorm is then a separate launch
  const long long groups = (total + PE_TOK - 1) / PE_TOK;
  const unsigned grid = (unsigned)(groups < 2LL * sms ? groups : 2LL * sms);
  patch_embed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, w, b, gamma, beta, x, B, S, (int)groups);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (y16) return mst_layernorm(x, gamma1, beta1, y16, (int)total, 128, stream);
  return 0;
}

extern "C" int mst_patch_embed_ln_u8_supported(int S) { return S > 0 && S % 16 == 0; }

extern "C" int mst_patch_embed_ln_u8(const uint8_t* img_u8, const float* mean3, const float* std3, const float* w, const float* b,
                                     const float* gamma, const float* beta, float* x, const float* gamma1, const float* beta1,
                                     mst_bf16* y16, int B, int S, void* stream) {
  if (!img_u8 || !w || !b || !gamma || !beta || !x || B <= 0 || S <= 0) return MST_ERR_BAD_ARG;
  if ((mean3 == nullptr) != (std3 == nullptr)) return MST_ERR_BAD_ARG;
  if (y16 && (!gamma1 || !beta1)) return MST_ERR_BAD_ARG;
  if (!mst_patch_embed_ln_u8_supported(S)) return MST_ERR_UNSUPPORTED;
  bool handled = false;
  const int rc = mst::patch_embed_tc_try(nullptr, w, b, gamma, beta, x, gamma1, beta1, reinterpret_cast<bf16*>(y16), B, S, (cudaStream_t)stream, handled,
                                         img_u8, mean3, std3);
  return handled ? rc : MST_ERR_BAD_ARG;  // not handled: misaligned pointers
}

extern "C" int mst_upsample2x_nhwc(const mst_bf16* x, mst_bf16* y, int B, int H, int W, int C, void* stream) {
  if (!x || !y || B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0) return MST_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return MST_ERR_BAD_ARG;
  const long long n = (long long)B * H * W * (C / 8);
  upsample2x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), n, H, W,
                                                                                  C / 8);
  return (int)cudaGetLastError();
}

extern "C" int mst_cast_bf16(const float* x, mst_bf16* y, size_t n, void* stream) {
  if (!x || !y || n == 0 || n % 4 != 0) return MST_ERR_BAD_ARG;
  const size_t n4 = n / 4;
  cast_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<bf16*>(y), n4);
  return (int)cudaGetLastError();
}

extern "C" int mst_images_u8_to_nchw(const uint8_t* src, float* dst, int B, int H, int W, const float* mean3, const float* std3,
                                     void* stream) {
  if (!src || !dst || B <= 0 || H <= 0 || W <= 0 || W % 4 != 0 || (mean3 == nullptr) != (std3 == nullptr)) return MST_ERR_BAD_ARG;
  if (((reinterpret_cast<uintptr_t>(src) & 3) | (reinterpret_cast<uintptr_t>(dst) & 15)) != 0) return MST_ERR_BAD_ARG;
  const long long plane = (long long)H * W, quads = (long long)B * plane / 4;
  const float m[3] = {mean3 ? mean3[0] : 0.f, mean3 ? mean3[1] : 0.f, mean3 ? mean3[2] : 0.f};
  const float sd[3] = {std3 ? std3[0] : 1.f, std3 ? std3[1] : 1.f, std3 ? std3[2] : 1.f};
  images_u8_to_nchw_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(src), dst, quads, plane, m[0], m[1], m[2], sd[0], sd[1], sd[2], mean3 != nullptr);
  return (int)cudaGetLastError();
}

extern "C" int mst_images_nchw_to_u8(const float* src, uint8_t* dst, int B, int H, int W, void* stream) {
  if (!src || !dst || B <= 0 || H <= 0 || W <= 0 || W % 4 != 0) return MST_ERR_BAD_ARG;
  if (((reinterpret_cast<uintptr_t>(dst) & 3) | (reinterpret_cast<uintptr_t>(src) & 15)) != 0) return MST_ERR_BAD_ARG;
  const long long plane = (long long)H * W, quads = (long long)B * plane / 4;
  images_nchw_to_u8_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<uint32_t*>(dst), quads, plane);
  return (int)cudaGetLastError();
}

namespace mst {
// ---------------------------------------------------------------- training-image transform (codes/get_dataloader.py:30-36)
// ToPILImage -> Resize((512, 512)) -> RandomCrop((256, 256)) -> ToTensor -> Normalize as ONE kernel on a decoded uint8 HWC image:
// only the cropped window of the resized image is ever computed.  The resize is Pillow's antialiased bilinear resample
// (libImaging/Resample.c, 8 bits per channel), restated exactly: a horizontal pass into an 8-bit intermediate, then a vertical
// pass, each  clip8((2^21 + sum_t pixel_t * k_t) >> 22)  with the fixed-point coefficients k = (int)(0.5 + w * 2^22) that the host
// precomputes the way precompute_coeffs / normalize_coeffs_8bpc do (data.pil_resize_coeffs); then ((v / 255) - mean) / std in the
// fp32 operation order of torchvision (bit-exact against the reference's transform pipeline, tests/test_gpu_kernels.py).
// Thread = one output pixel (three channels); the intermediate values a thread needs are recomputed, not stored.
__global__ void __launch_bounds__(256) resize_crop_normalize_kernel(const uint8_t* __restrict__ img, int H, int W, const int32_t* __restrict__ xmin,
                                                                    const int32_t* __restrict__ xcnt, const int32_t* __restrict__ xk, int ksx,
                                                                    const int32_t* __restrict__ ymin, const int32_t* __restrict__ ycnt,
                                                                    const int32_t* __restrict__ yk, int ksy, int top, int left, int ch, int cw,
                                                                    float m0, float m1, float m2, float s0, float s1, float s2, int normalize,
                                                                    float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cw || y >= ch) return;
  const int X = left + x, Y = top + y;           // position in the resized image
  const int x0 = xmin[X], nx = xcnt[X], y0 = ymin[Y], ny = ycnt[Y];
  const int32_t* kx = xk + (long long)X * ksx;
  const int32_t* ky = yk + (long long)Y * ksy;
  int acc[3] = {1 << 21, 1 << 21, 1 << 21};
  for (int v = 0; v < ny; ++v) {
    const uint8_t* row = img + ((long long)(y0 + v) * W + x0) * 3;
    int h[3] = {1 << 21, 1 << 21, 1 << 21};
    for (int u = 0; u < nx; ++u) {
      const int k = kx[u];
      h[0] += (int)row[3 * u] * k;
      h[1] += (int)row[3 * u + 1] * k;
      h[2] += (int)row[3 * u + 2] * k;
    }
    const int kv = ky[v];
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += min(max(h[c] >> 22, 0), 255) * kv;  // the 8-bit intermediate image of the horizontal pass
  }
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float f = __fdiv_rn((float)min(max(acc[c] >> 22, 0), 255), 255.0f);
    if (normalize) f = __fdiv_rn(__fsub_rn(f, mean[c]), sd[c]);
    out[((long long)c * ch + y) * cw + x] = f;
  }
}
}  // namespace mst

extern "C" int mst_resize_crop_normalize(const uint8_t* img, int H, int W, const int32_t* xmin, const int32_t* xcnt, const int32_t* xk, int ksx,
                                         const int32_t* ymin, const int32_t* ycnt, const int32_t* yk, int ksy, int top, int left, int ch, int cw,
                                         const float* mean3, const float* std3, float* out, void* stream) {
  if (!img || !xmin || !xcnt || !xk || !ymin || !ycnt || !yk || !out || H <= 0 || W <= 0 || ksx <= 0 || ksy <= 0) return MST_ERR_BAD_ARG;
  if (top < 0 || left < 0 || ch <= 0 || cw <= 0 || (mean3 == nullptr) != (std3 == nullptr)) return MST_ERR_BAD_ARG;
  const float m[3] = {mean3 ? mean3[0] : 0.f, mean3 ? mean3[1] : 0.f, mean3 ? mean3[2] : 0.f};
  const float sd[3] = {std3 ? std3[0] : 1.f, std3 ? std3[1] : 1.f, std3 ? std3[2] : 1.f};
  dim3 grid((cw + 255) / 256, ch);
  mst::resize_crop_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, H, W, xmin, xcnt, xk, ksx, ymin, ycnt, yk, ksy, top, left, ch, cw, m[0],
                                                                            m[1], m[2], sd[0], sd[1], sd[2], mean3 != nullptr, out);
  return (int)cudaGetLastError();
}

extern "C" int mst_version(void) { return 100; }
extern "C" int mst_sm_arch(void) { return 100; }
extern "C" const char* mst_error_string(int code) {
  if (code == 0) return "ok";
  if (code == MST_ERR_BAD_ARG) return "bad argument";
  if (code == MST_ERR_UNSUPPORTED) return "unsupported shape";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}
