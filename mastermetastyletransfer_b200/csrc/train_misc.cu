// HBM-streaming backward kernels of the training step: LayerNorm / InstanceNorm adjoints,
// the Query*sigma+mu blend adjoint, reflect-pad fold, max-pool adjoint, and the content/style loss gradient w.r.t.
// the VGG taps of the stylised image.  All are bandwidth-bound: 16-byte accesses, one pass where the maths allows.
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

MST_DEVINL void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}
MST_DEVINL uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------- LayerNorm backward: one warp per row, rows strided
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;   dgamma += dy * xhat,  dbeta += dy
template <int C>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const bf16* __restrict__ dy, float* __restrict__ dx_accum,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int rows) {
  constexpr int V4 = C / 128;
  __shared__ float red[2][8][C + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 gm[V4];
#pragma unroll
  for (int i = 0; i < V4; ++i) gm[i] = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
  float4 ag[V4], ab[V4];
#pragma unroll
  for (int i = 0; i < V4; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * C);
    const uint2* dr = reinterpret_cast<const uint2*>(dy + (long long)row * C);
    float4 v[V4], d[V4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      v[i] = xr[lane + 32 * i];
      const uint2 p = dr[lane + 32 * i];
      d[i] = make_float4(__uint_as_float(p.x << 16), __uint_as_float(p.x & 0xffff0000u), __uint_as_float(p.y << 16),
                         __uint_as_float(p.y & 0xffff0000u));
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
      ag[i].x += d[i].x * v[i].x; ag[i].y += d[i].y * v[i].y; ag[i].z += d[i].z * v[i].z; ag[i].w += d[i].w * v[i].w;
      ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
      d[i].x *= gm[i].x; d[i].y *= gm[i].y; d[i].z *= gm[i].z; d[i].w *= gm[i].w;  // g
      m1 += d[i].x + d[i].y + d[i].z + d[i].w;
      m2 += d[i].x * v[i].x + d[i].y * v[i].y + d[i].z * v[i].z + d[i].w * v[i].w;
    }
    m1 = warp_sum(m1) * (1.0f / C);
    m2 = warp_sum(m2) * (1.0f / C);
    float4* o = reinterpret_cast<float4*>(dx_accum + (long long)row * C);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      float4 acc = o[lane + 32 * i];
      acc.x += rstd * (d[i].x - m1 - v[i].x * m2);
      acc.y += rstd * (d[i].y - m1 - v[i].y * m2);
      acc.z += rstd * (d[i].z - m1 - v[i].z * m2);
      acc.w += rstd * (d[i].w - m1 - v[i].w * m2);
      o[lane + 32 * i] = acc;
    }
  }
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    *reinterpret_cast<float4*>(&red[0][warp][(lane + 32 * i) * 4]) = ag[i];
    *reinterpret_cast<float4*>(&red[1][warp][(lane + 32 * i) * 4]) = ab[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    const int which = c / C, cc = c - which * C;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[which][w][cc];
    atomicAdd((which ? dbeta : dgamma) + cc, s);
  }
}

// ---------------------------------------------------------------- InstanceNorm backward
// y = (x - mu) * g(v), v = biased variance over T.  dx_t = g*(dy_t - mean(dy)) + (2 g'(v)/T) * (x_t - mu) * sum_s dy_s (x_s - mu)
// once:  g = (v+eps)^-1/2, g' = -g^3/2.   twice (IN(IN(x))): g = r1 r2, r1 = (v+eps)^-1/2, r2 = (v r1^2 + eps)^-1/2,
//        g' = -r1^3 r2 / 2 - eps r1^5 r2^3 / 2.
template <bool F32>
__global__ void __launch_bounds__(256) instnorm_bwd_stats_kernel(const float* __restrict__ x, const void* __restrict__ dyv,
                                                                 float* __restrict__ coef, int T, int C, int twice) {
  __shared__ float red[3][8][33];
  const int groups = C / 32;
  const int b = blockIdx.x / groups;
  const int c = (blockIdx.x - b * groups) * 32 + (threadIdx.x & 31);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + (long long)b * T * C + c;
  float s = 0.f;
  for (int t = warp; t < T; t += 8) s += xb[(long long)t * C];
  red[0][warp][lane] = s;
  __syncthreads();
  float m = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) m += red[0][w][lane];
  m /= (float)T;
  __syncthreads();
  float q = 0.f, s1 = 0.f, S = 0.f;
  for (int t = warp; t < T; t += 8) {
    const long long off = (long long)b * T * C + (long long)t * C + c;
    const float d = xb[(long long)t * C] - m;
    const float g = F32 ? reinterpret_cast<const float*>(dyv)[off] : __bfloat162float(reinterpret_cast<const bf16*>(dyv)[off]);
    q += d * d;
    s1 += g;
    S += g * d;
  }
  red[0][warp][lane] = q; red[1][warp][lane] = s1; red[2][warp][lane] = S;
  __syncthreads();
  if (warp == 0) {
    q = s1 = S = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { q += red[0][w][lane]; s1 += red[1][w][lane]; S += red[2][w][lane]; }
    const float eps = 1e-5f;
    const float v = q / (float)T;
    const float r1 = 1.0f / sqrtf(v + eps);
    float g = r1, gp = -0.5f * r1 * r1 * r1;
    if (twice) {
      const float r2 = 1.0f / sqrtf(v * r1 * r1 + eps);
      g = r1 * r2;
      gp = -0.5f * r1 * r1 * r1 * r2 - 0.5f * eps * r1 * r1 * r1 * r1 * r1 * r2 * r2 * r2;
    }
    float4 o = make_float4(g, s1 / (float)T, 2.0f * gp * S / (float)T, m);
    reinterpret_cast<float4*>(coef)[(long long)b * C + c] = o;
  }
}

template <bool F32>
__global__ void __launch_bounds__(256) instnorm_bwd_apply_kernel(const float* __restrict__ x, const void* __restrict__ dyv,
                                                                 const float* __restrict__ coef, float* __restrict__ dx_accum,
                                                                 bf16* __restrict__ dx16, long long n4, int TC4, int C4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int b = (int)(i / TC4);
  const int c4 = (int)(i % C4);
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  float4 d;
  if (F32) {
    d = reinterpret_cast<const float4*>(dyv)[i];
  } else {
    const uint2 p = reinterpret_cast<const uint2*>(dyv)[i];
    d = make_float4(__uint_as_float(p.x << 16), __uint_as_float(p.x & 0xffff0000u), __uint_as_float(p.y << 16),
                    __uint_as_float(p.y & 0xffff0000u));
  }
  const float4* cf = reinterpret_cast<const float4*>(coef) + ((long long)b * C4 + c4) * 4;
  const float4 k0 = cf[0], k1 = cf[1], k2 = cf[2], k3 = cf[3];
  float4 o;
  o.x = k0.x * (d.x - k0.y) + k0.z * (v.x - k0.w);
  o.y = k1.x * (d.y - k1.y) + k1.z * (v.y - k1.w);
  o.z = k2.x * (d.z - k2.y) + k2.z * (v.z - k2.w);
  o.w = k3.x * (d.w - k3.y) + k3.z * (v.w - k3.w);
  if (dx_accum) {
    float4 a = reinterpret_cast<float4*>(dx_accum)[i];
    a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
    reinterpret_cast<float4*>(dx_accum)[i] = a;
  }
  if (dx16) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
    reinterpret_cast<uint2*>(dx16)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// ---------------------------------------------------------------- blend adjoint, add/cast
__global__ void __launch_bounds__(256) blend_bwd_kernel(const float4* __restrict__ gy, const float4* __restrict__ sigma,
                                                        const float4* __restrict__ query, float4* __restrict__ gquery,
                                                        uint2* __restrict__ gsigma16, uint2* __restrict__ gmu16, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 g = gy[i], s = sigma[i], q = query[i];
  gquery[i] = make_float4(g.x * s.x, g.y * s.y, g.z * s.z, g.w * s.w);
  __nv_bfloat162 a = __floats2bfloat162_rn(g.x * q.x, g.y * q.y), b = __floats2bfloat162_rn(g.z * q.z, g.w * q.w);
  gsigma16[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  a = __floats2bfloat162_rn(g.x, g.y);
  b = __floats2bfloat162_rn(g.z, g.w);
  gmu16[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

__global__ void __launch_bounds__(256) add_cast_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                       float4* __restrict__ out32, uint2* __restrict__ out16, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = a[i];
  if (b) {
    const float4 w = b[i];
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  if (out32) out32[i] = v;
  if (out16) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    out16[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// ---------------------------------------------------------------- token-map pad / crop (7x7 windows on padded maps, training path)
// src [B,Hs,Ws,*] -> dst [B,Hd,Wd,*] in 16-byte units (`u` per token): dst tokens outside the source map are zero (pad),
// source tokens outside the destination map are dropped (crop).  ACC: dst (fp32) += src instead of a copy.
template <bool ACC>
__global__ void __launch_bounds__(256) token_map_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int Hs, int Ws, int Hd,
                                                        int Wd, int u, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % u);
  long long t = i / u;
  const int x = (int)(t % Wd);
  t /= Wd;
  const int y = (int)(t % Hd);
  const long long b = t / Hd;
  const bool inside = y < Hs && x < Ws;
  if (ACC) {
    if (!inside) return;
    const uint4 sv = src[((b * Hs + y) * Ws + x) * u + c];
    float4 d = *reinterpret_cast<float4*>(dst + i);
    d.x += __uint_as_float(sv.x); d.y += __uint_as_float(sv.y); d.z += __uint_as_float(sv.z); d.w += __uint_as_float(sv.w);
    *reinterpret_cast<float4*>(dst + i) = d;
  } else {
    dst[i] = inside ? src[((b * Hs + y) * Ws + x) * u + c] : make_uint4(0u, 0u, 0u, 0u);
  }
}

// ---------------------------------------------------------------- reflect-pad fold (+ nearest-x2 upsample adjoint, + ReLU mask)
// dxp [B,H+2,W+2,C] on the padded grid; padded row 0 mirrors input row 1, padded row H+1 mirrors input row H-2.
__global__ void __launch_bounds__(256) reflect_fold_kernel(const bf16* __restrict__ dxp, const bf16* __restrict__ gate,
                                                           bf16* __restrict__ dx, int B, int H, int W, int C8, int upsample) {
  const int Ho = upsample ? H >> 1 : H, Wo = upsample ? W >> 1 : W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * Ho * Wo * C8;
  if (i >= total) return;
  const int c8 = (int)(i % C8);
  long long pix = i / C8;
  const int xo = (int)(pix % Wo); pix /= Wo;
  const int yo = (int)(pix % Ho);
  const int b = (int)(pix / Ho);
  const int Wp = W + 2;
  const uint4* src = reinterpret_cast<const uint4*>(dxp) + (long long)b * (H + 2) * Wp * C8 + c8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int reps = upsample ? 2 : 1;
  for (int dy = 0; dy < reps; ++dy) {
    const int y = upsample ? 2 * yo + dy : yo;
    int ys[2] = {y + 1, -1};
    if (y == 1) ys[1] = 0;
    else if (y == H - 2) ys[1] = H + 1;
    for (int dxx = 0; dxx < reps; ++dxx) {
      const int x = upsample ? 2 * xo + dxx : xo;
      int xs[2] = {x + 1, -1};
      if (x == 1) xs[1] = 0;
      else if (x == W - 2) xs[1] = W + 1;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (ys[a] < 0) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (xs[c] < 0) continue;
          float f[8];
          unpack8(src[((long long)ys[a] * Wp + xs[c]) * C8], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += f[e];
        }
      }
    }
  }
  if (gate) {
    float gt[8];
    unpack8(reinterpret_cast<const uint4*>(gate)[i], gt);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = gt[e] > 0.f ? acc[e] : 0.f;
  }
  reinterpret_cast<uint4*>(dx)[i] = pack8(acc);
}

// ---------------------------------------------------------------- MaxPool2d(2) adjoint fused with the preceding ReLU mask
__global__ void __launch_bounds__(256) maxpool2x2_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                             bf16* __restrict__ dx, int B, int H, int W, int C8) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * Ho * Wo * C8;
  if (i >= total) return;
  const int c8 = (int)(i % C8);
  long long pix = i / C8;
  const int xo = (int)(pix % Wo); pix /= Wo;
  const int yo = (int)(pix % Ho);
  const int b = (int)(pix / Ho);
  const long long base = (((long long)b * H + 2 * yo) * W + 2 * xo) * C8 + c8;
  const long long offs[4] = {0, C8, (long long)W * C8, (long long)W * C8 + C8};
  float v[4][8], g[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) unpack8(reinterpret_cast<const uint4*>(x)[base + offs[k]], v[k]);
  unpack8(reinterpret_cast<const uint4*>(dy)[i], g);
  float o[4][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    int best = 0;
    float mx = v[0][e];
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (v[k][e] > mx) { mx = v[k][e]; best = k; }
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k][e] = (k == best && mx > 0.f) ? g[e] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) reinterpret_cast<uint4*>(dx)[base + offs[k]] = pack8(o[k]);
}

__global__ void __launch_bounds__(256) nchw3_to_nhwc8_kernel(const float* __restrict__ g, bf16* __restrict__ out, int B, int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW) return;
  const int b = (int)(i / HW);
  const int p = (int)(i - (long long)b * HW);
  float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 3; ++c) f[c] = g[((long long)b * 3 + c) * HW + p];
  reinterpret_cast<uint4*>(out)[i] = pack8(f);
}

// ---------------------------------------------------------------- loss backward
// stats: s[b,c,0] = sum_t g, s[b,c,1] = sum_t g * n_o,  g = -phi'(n_c - n_o)  (phi = |.| or (.)^2)
// CTA = (T slab, b, 64-channel group); thread = (8-channel chunk, row lane)
__global__ void __launch_bounds__(256) loss_bwd_stats_kernel(const bf16* __restrict__ fc, const bf16* __restrict__ fo,
                                                             const float* __restrict__ mean_c, const float* __restrict__ var_c,
                                                             const float* __restrict__ mean_o, const float* __restrict__ var_o,
                                                             int T, int C, int squared, float* __restrict__ s, int rows_per_cta) {
  __shared__ float red[2][32][65];
  const int c8 = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int b = blockIdx.y;
  const int col = blockIdx.z * 64 + c8 * 8;
  const int t0 = blockIdx.x * rows_per_cta, t1 = min(T, t0 + rows_per_cta);
  float mc[8], rc[8], mo[8], ro[8], a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const long long k = (long long)b * C + col + e;
    mc[e] = mean_c[k]; rc[e] = rsqrtf(var_c[k] + 1e-5f);
    mo[e] = mean_o[k]; ro[e] = rsqrtf(var_o[k] + 1e-5f);
    a0[e] = a1[e] = 0.f;
  }
  for (int t = t0 + rl; t < t1; t += 32) {
    const long long off = ((long long)b * T + t) * C + col;
    float c[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(fc + off), c);
    unpack8(*reinterpret_cast<const uint4*>(fo + off), o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float no = (o[e] - mo[e]) * ro[e];
      const float d = (c[e] - mc[e]) * rc[e] - no;
      const float g = squared ? -2.0f * d : (d > 0.f ? -1.0f : (d < 0.f ? 1.0f : 0.f));
      a0[e] += g;
      a1[e] += g * no;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { red[0][rl][c8 * 8 + e] = a0[e]; red[1][rl][c8 * 8 + e] = a1[e]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, cc = threadIdx.x & 63;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += red[which][i][cc];
    atomicAdd(s + ((long long)b * C + blockIdx.z * 64 + cc) * 2 + which, acc);
  }
}

__global__ void __launch_bounds__(256) loss_bwd_apply_kernel(const bf16* __restrict__ fc, const bf16* __restrict__ fo,
                                                             const float* __restrict__ mean_c, const float* __restrict__ var_c,
                                                             const float* __restrict__ mean_o, const float* __restrict__ var_o,
                                                             const float* __restrict__ mean_s, const float* __restrict__ var_s,
                                                             const float* __restrict__ s, const float* __restrict__ w, int B, int T,
                                                             int C, int sq_c, int sq_s, bf16* __restrict__ dfo, int rows_per_cta) {
  const int c8 = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int b = blockIdx.y;
  const int col = blockIdx.z * 64 + c8 * 8;
  const int t0 = blockIdx.x * rows_per_cta, t1 = min(T, t0 + rows_per_cta);
  const float wc = w[0] / ((float)B * (float)T * (float)C), wsty = w[1] / ((float)B * (float)C);
  const float invT = 1.0f / (float)T, unb = (float)T / (float)(T - 1);
  float mc[8], rc[8], mo[8], ro[8], k0[8], k1[8], sa[8], sb[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const long long k = (long long)b * C + col + e;
    mc[e] = mean_c[k]; rc[e] = rsqrtf(var_c[k] + 1e-5f);
    mo[e] = mean_o[k]; ro[e] = rsqrtf(var_o[k] + 1e-5f);
    k0[e] = s[k * 2] * invT;
    k1[e] = s[k * 2 + 1] * invT;
    const float dmu = mean_s[k] - mo[e];
    const float so = sqrtf(var_o[k] * unb), ss = sqrtf(var_s[k] * unb);
    const float dsd = ss - so;
    const float pm = sq_s ? 2.0f * dmu : (dmu > 0.f ? 1.0f : (dmu < 0.f ? -1.0f : 0.f));
    const float ps = sq_s ? 2.0f * dsd : (dsd > 0.f ? 1.0f : (dsd < 0.f ? -1.0f : 0.f));
    sa[e] = -wsty * pm * invT;
    sb[e] = so > 1e-12f ? -wsty * ps / ((float)(T - 1) * so) : 0.f;
  }
  for (int t = t0 + rl; t < t1; t += 32) {
    const long long off = ((long long)b * T + t) * C + col;
    float c[8], o[8], out[8];
    unpack8(*reinterpret_cast<const uint4*>(fc + off), c);
    unpack8(*reinterpret_cast<const uint4*>(fo + off), o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float no = (o[e] - mo[e]) * ro[e];
      const float d = (c[e] - mc[e]) * rc[e] - no;
      const float g = sq_c ? -2.0f * d : (d > 0.f ? -1.0f : (d < 0.f ? 1.0f : 0.f));
      const float v = wc * ro[e] * (g - k0[e] - no * k1[e]) + sa[e] + sb[e] * (o[e] - mo[e]);
      out[e] = o[e] > 0.f ? v : 0.f;
    }
    *reinterpret_cast<uint4*>(dfo + off) = pack8(out);
  }
}

static inline unsigned blocks_for(long long n, int per = 256) { return (unsigned)((n + per - 1) / per); }

}  // namespace mst

using namespace mst;

extern "C" int mst_layernorm_bwd(const float* x, const float* gamma, const mst_bf16* dy, float* dx_accum, float* dgamma, float* dbeta,
                                 int rows, int C, void* stream) {
  if (!x || !gamma || !dy || !dx_accum || !dgamma || !dbeta || rows <= 0) return MST_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (rows + 7) / 8;
  if (grid > 296) grid = 296;
  const bf16* d = reinterpret_cast<const bf16*>(dy);
  if (C == 256) layernorm_bwd_kernel<256><<<grid, 256, 0, st>>>(x, gamma, d, dx_accum, dgamma, dbeta, rows);
  else if (C == 128) layernorm_bwd_kernel<128><<<grid, 256, 0, st>>>(x, gamma, d, dx_accum, dgamma, dbeta, rows);
  else return MST_ERR_UNSUPPORTED;
  return (int)cudaGetLastError();
}

extern "C" int mst_instnorm_bwd_stats(const float* x, const void* dy, int dy_is_f32, float* coef, int B, int T, int C, int twice,
                                      void* stream) {
  if (!x || !dy || !coef || B <= 0 || T <= 0 || C <= 0 || C % 32 != 0) return MST_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dy_is_f32) instnorm_bwd_stats_kernel<true><<<B * (C / 32), 256, 0, st>>>(x, dy, coef, T, C, twice);
  else instnorm_bwd_stats_kernel<false><<<B * (C / 32), 256, 0, st>>>(x, dy, coef, T, C, twice);
  return (int)cudaGetLastError();
}

extern "C" int mst_instnorm_bwd_apply(const float* x, const void* dy, int dy_is_f32, const float* coef, float* dx_accum,
                                      mst_bf16* dx16, int B, int T, int C, void* stream) {
  if (!x || !dy || !coef || (!dx_accum && !dx16) || B <= 0 || T <= 0 || C <= 0 || C % 4 != 0) return MST_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n4 = (long long)B * T * C / 4;
  if (dy_is_f32)
    instnorm_bwd_apply_kernel<true><<<blocks_for(n4), 256, 0, st>>>(x, dy, coef, dx_accum, reinterpret_cast<bf16*>(dx16), n4, T * C / 4, C / 4);
  else
    instnorm_bwd_apply_kernel<false><<<blocks_for(n4), 256, 0, st>>>(x, dy, coef, dx_accum, reinterpret_cast<bf16*>(dx16), n4, T * C / 4, C / 4);
  return (int)cudaGetLastError();
}

extern "C" int mst_blend_bwd(const float* gy, const float* sigma, const float* query, float* gquery, mst_bf16* gsigma16,
                             mst_bf16* gmu16, size_t n, void* stream) {
  if (!gy || !sigma || !query || !gquery || !gsigma16 || !gmu16 || n == 0 || n % 4 != 0) return MST_ERR_BAD_ARG;
  const long long n4 = (long long)(n / 4);
  blend_bwd_kernel<<<blocks_for(n4), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(gy), reinterpret_cast<const float4*>(sigma), reinterpret_cast<const float4*>(query),
      reinterpret_cast<float4*>(gquery), reinterpret_cast<uint2*>(gsigma16), reinterpret_cast<uint2*>(gmu16), n4);
  return (int)cudaGetLastError();
}

extern "C" int mst_add_cast(const float* a, const float* b, float* out32, mst_bf16* out16, size_t n, void* stream) {
  if (!a || (!out32 && !out16) || n == 0 || n % 4 != 0) return MST_ERR_BAD_ARG;
  const long long n4 = (long long)(n / 4);
  add_cast_kernel<<<blocks_for(n4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                                   reinterpret_cast<float4*>(out32), reinterpret_cast<uint2*>(out16), n4);
  return (int)cudaGetLastError();
}

extern "C" int mst_token_map_copy(const void* src, void* dst, int B, int Hs, int Ws, int Hd, int Wd, int token_bytes, int accumulate_f32,
                                  void* stream) {
  if (!src || !dst || B <= 0 || Hs <= 0 || Ws <= 0 || Hd <= 0 || Wd <= 0 || token_bytes <= 0 || token_bytes % 16 != 0) return MST_ERR_BAD_ARG;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) != 0) return MST_ERR_BAD_ARG;
  const int u = token_bytes / 16;
  const long long total = (long long)B * Hd * Wd * u;
  if (accumulate_f32)
    token_map_kernel<true><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst),
                                                                               Hs, Ws, Hd, Wd, u, total);
  else
    token_map_kernel<false><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst),
                                                                                Hs, Ws, Hd, Wd, u, total);
  return (int)cudaGetLastError();
}

extern "C" int mst_reflect_fold(const mst_bf16* dxp, const mst_bf16* gate, mst_bf16* dx, int B, int H, int W, int C, int upsample,
                                void* stream) {
  if (!dxp || !dx || B <= 0 || H < 4 || W < 4 || C <= 0 || C % 8 != 0) return MST_ERR_BAD_ARG;
  if (upsample && ((H | W) & 1)) return MST_ERR_BAD_ARG;
  const long long total = (long long)B * (upsample ? H / 2 : H) * (upsample ? W / 2 : W) * (C / 8);
  reflect_fold_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const bf16*>(dxp), reinterpret_cast<const bf16*>(gate),
                                                                          reinterpret_cast<bf16*>(dx), B, H, W, C / 8, upsample);
  return (int)cudaGetLastError();
}

extern "C" int mst_maxpool2x2_bwd(const mst_bf16* x, const mst_bf16* dy, mst_bf16* dx, int B, int H, int W, int C, void* stream) {
  if (!x || !dy || !dx || B <= 0 || H < 2 || W < 2 || ((H | W) & 1) || C <= 0 || C % 8 != 0) return MST_ERR_BAD_ARG;
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  maxpool2x2_bwd_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dy),
                                                                            reinterpret_cast<bf16*>(dx), B, H, W, C / 8);
  return (int)cudaGetLastError();
}

extern "C" int mst_nchw3_to_nhwc8(const float* g, mst_bf16* out, int B, int H, int W, void* stream) {
  if (!g || !out || B <= 0 || H <= 0 || W <= 0) return MST_ERR_BAD_ARG;
  const long long total = (long long)B * H * W;
  nchw3_to_nhwc8_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(g, reinterpret_cast<bf16*>(out), B, H * W);
  return (int)cudaGetLastError();
}

extern "C" int mst_loss_bwd_stats(const mst_bf16* fc, const mst_bf16* fo, const float* mean_c, const float* var_c, const float* mean_o,
                                  const float* var_o, int B, int T, int C, int squared, float* s, void* stream) {
  if (!fc || !fo || !mean_c || !var_c || !mean_o || !var_o || !s || B <= 0 || T <= 1 || C <= 0 || C % 64 != 0) return MST_ERR_BAD_ARG;
  const int rows_per_cta = 256;
  dim3 grid((unsigned)((T + rows_per_cta - 1) / rows_per_cta), (unsigned)B, (unsigned)(C / 64));
  loss_bwd_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const bf16*>(fc), reinterpret_cast<const bf16*>(fo), mean_c,
                                                                var_c, mean_o, var_o, T, C, squared, s, rows_per_cta);
  return (int)cudaGetLastError();
}

extern "C" int mst_loss_bwd_apply(const mst_bf16* fc, const mst_bf16* fo, const float* mean_c, const float* var_c, const float* mean_o,
                                  const float* var_o, const float* mean_s, const float* var_s, const float* s, const float* w, int B,
                                  int T, int C, int squared_content, int squared_style, mst_bf16* dfo, void* stream) {
  if (!fc || !fo || !mean_c || !var_c || !mean_o || !var_o || !mean_s || !var_s || !s || !w || !dfo) return MST_ERR_BAD_ARG;
  if (B <= 0 || T <= 1 || C <= 0 || C % 64 != 0) return MST_ERR_BAD_ARG;
  const int rows_per_cta = 256;
  dim3 grid((unsigned)((T + rows_per_cta - 1) / rows_per_cta), (unsigned)B, (unsigned)(C / 64));
  loss_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const bf16*>(fc), reinterpret_cast<const bf16*>(fo), mean_c,
                                                                var_c, mean_o, var_o, mean_s, var_s, s, w, B, T, C, squared_content,
                                                                squared_style, reinterpret_cast<bf16*>(dfo), rows_per_cta);
  return (int)cudaGetLastError();
}
