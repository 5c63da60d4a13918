// Fused shifted-window attention core (codes/style_transformer.py:83-111,127-168 and :544-607).
// One CTA per window, one warp per head (head_dim 32).  Roll + partition live in the load addresses,
// window reverse + roll back in the store addresses; scores, relative-position bias, 9-region mask,
// softmax and PV stay in registers / shared memory in fp32.  With v2/out2 set the softmax is shared by
// two value tensors (the sigma/mu attention of the style decoder).
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

template <int WS>
__global__ void __launch_bounds__(256) window_attn_kernel(const MstWindowAttn a, const WinGeom g) {
  constexpr int N = WS * WS;
  constexpr int D = 32;
  constexpr int NT = (2 * WS - 1) * (2 * WS - 1);
  extern __shared__ float smem[];
  // per warp: K [N][D], V [N][D], (V2 [N][D]); per CTA: bias table [NT*heads], labels [N], src [N]
  const int heads = a.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dual = a.v2 != nullptr;
  const int per_warp = N * D * (dual ? 3 : 2);
  float* table_s = smem + heads * per_warp;
  int* src_s = reinterpret_cast<int*>(table_s + NT * heads);
  int* lab_s = src_s + N;
  float* Ks = smem + warp * per_warp;
  float* Vs = Ks + N * D;
  float* V2s = Vs + N * D;

  const int win_global = blockIdx.x;
  const int b = win_global / g.nW;
  const int win = win_global - b * g.nW;
  const bool masked = (g.sy + g.sx) > 0;

  for (int i = threadIdx.x; i < NT * heads; i += blockDim.x) table_s[i] = a.bias_table[i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    int y, x;
    win_source(g, win, i, y, x);
    src_s[i] = (y < g.H && x < g.W) ? (b * g.H + y) * g.W + x : -1;  // token row in the [B*H*W, C] tensors
    lab_s[i] = win_label(g, win, i);
  }
  __syncthreads();

  const int h = warp;
  const int c0 = h * D;
  // ---- stage K, V (and V2) of this head: lane = channel, loop over tokens (64-byte coalesced rows) ----
  for (int j = 0; j < N; ++j) {
    const int s = src_s[j];
    float kv, vv, v2v = 0.f;
    if (s >= 0) {
      kv = __bfloat162float(reinterpret_cast<const bf16*>(a.k)[(long long)s * a.ldk + c0 + lane]);
      vv = __bfloat162float(reinterpret_cast<const bf16*>(a.v)[(long long)s * a.ldv + c0 + lane]);
      if (dual) v2v = __bfloat162float(reinterpret_cast<const bf16*>(a.v2)[(long long)s * a.ldv + c0 + lane]);
    } else {
      kv = a.pad_k ? a.pad_k[c0 + lane] : 0.f;
      vv = a.pad_v ? a.pad_v[c0 + lane] : 0.f;
      if (dual) v2v = a.pad_v2 ? a.pad_v2[c0 + lane] : 0.f;
    }
    Ks[j * D + lane] = kv;
    Vs[j * D + lane] = vv;
    if (dual) V2s[j * D + lane] = v2v;
  }
  __syncwarp();

  const float scale = 0.17677669529663687f;  // 32^-0.5
  for (int i = lane; i < N; i += 32) {
    // ---- q row (scaled as the reference does before the matmul) ----
    float q[D];
    const int si = src_s[i];
    if (si >= 0) {
      const uint4* q4 = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.q) + (long long)si * a.ldq + c0);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint4 u = q4[t];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          q[t * 8 + 2 * e] = __uint_as_float(w[e] << 16) * scale;
          q[t * 8 + 2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u) * scale;
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) q[d] = (a.pad_q ? a.pad_q[c0 + d] : 0.f) * scale;
    }
    const int li = lab_s[i];
    // ---- scores ----
    float sc[N];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float4* k4 = reinterpret_cast<const float4*>(Ks + j * D);
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float4 kk = k4[t];
        acc = fmaf(q[4 * t], kk.x, acc);
        acc = fmaf(q[4 * t + 1], kk.y, acc);
        acc = fmaf(q[4 * t + 2], kk.z, acc);
        acc = fmaf(q[4 * t + 3], kk.w, acc);
      }
      acc += table_s[rel_pos_index(i, j, WS) * heads + h];
      if (masked && lab_s[j] != li) acc += -100.0f;
      sc[j] = acc;
      mx = fmaxf(mx, acc);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      sc[j] = __expf(sc[j] - mx);
      sum += sc[j];
    }
    const float inv = 1.0f / sum;
    // ---- PV ----
    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float4* v4 = reinterpret_cast<const float4*>(Vs + j * D);
      const float pj = sc[j];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float4 vv = v4[t];
        o[4 * t] = fmaf(pj, vv.x, o[4 * t]);
        o[4 * t + 1] = fmaf(pj, vv.y, o[4 * t + 1]);
        o[4 * t + 2] = fmaf(pj, vv.z, o[4 * t + 2]);
        o[4 * t + 3] = fmaf(pj, vv.w, o[4 * t + 3]);
      }
    }
    if (si >= 0) {
      uint32_t pk[16];
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        __nv_bfloat162 hh = __floats2bfloat162_rn(o[2 * d] * inv, o[2 * d + 1] * inv);
        pk[d] = *reinterpret_cast<uint32_t*>(&hh);
      }
      uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(a.out) + (long long)si * a.ldo + c0);
#pragma unroll
      for (int t = 0; t < 4; ++t) o4[t] = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
    }
    if (dual) {
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float4* v4 = reinterpret_cast<const float4*>(V2s + j * D);
        const float pj = sc[j];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float4 vv = v4[t];
          o[4 * t] = fmaf(pj, vv.x, o[4 * t]);
          o[4 * t + 1] = fmaf(pj, vv.y, o[4 * t + 1]);
          o[4 * t + 2] = fmaf(pj, vv.z, o[4 * t + 2]);
          o[4 * t + 3] = fmaf(pj, vv.w, o[4 * t + 3]);
        }
      }
      if (si >= 0) {
        uint32_t pk[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) {
          __nv_bfloat162 hh = __floats2bfloat162_rn(o[2 * d] * inv, o[2 * d + 1] * inv);
          pk[d] = *reinterpret_cast<uint32_t*>(&hh);
        }
        uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(a.out2) + (long long)si * a.ldo + c0);
#pragma unroll
        for (int t = 0; t < 4; ++t) o4[t] = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
      }
    }
  }
}

__global__ void window_maps_kernel(WinGeom g, int32_t* gather, int32_t* labels, int32_t* relidx) {
  const int N = g.ws * g.ws;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < g.nW * N) {
    const int win = tid / N, tok = tid - win * N;
    int y, x;
    win_source(g, win, tok, y, x);
    if (gather) gather[tid] = (y < g.H && x < g.W) ? y * g.W + x : -1;
    if (labels) labels[tid] = win_label(g, win, tok);
  }
  if (relidx && tid < N * N) relidx[tid] = rel_pos_index(tid / N, tid % N, g.ws);
}

template <int WS>
static int launch_attn(const MstWindowAttn& a, const WinGeom& g, cudaStream_t st) {
  constexpr int N = WS * WS, NT = (2 * WS - 1) * (2 * WS - 1);
  const size_t smem = sizeof(float) * ((size_t)a.heads * N * 32 * (a.v2 ? 3 : 2) + (size_t)NT * a.heads) + sizeof(int) * 2 * N;
  cudaError_t e = cudaFuncSetAttribute(window_attn_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return (int)e;
  if (smem > 200 * 1024) return MST_ERR_UNSUPPORTED;
  window_attn_kernel<WS><<<a.B * g.nW, a.heads * 32, smem, st>>>(a, g);
  return (int)cudaGetLastError();
}

}  // namespace mst

extern "C" int mst_window_attention(const MstWindowAttn* a, void* stream) {
  using namespace mst;
  if (!a || !a->q || !a->k || !a->v || !a->out || !a->bias_table) return MST_ERR_BAD_ARG;
  if ((a->v2 == nullptr) != (a->out2 == nullptr)) return MST_ERR_BAD_ARG;
  if (a->B <= 0 || a->H <= 0 || a->W <= 0 || a->heads <= 0 || a->heads > 8) return MST_ERR_BAD_ARG;
  if ((a->ldq | a->ldk | a->ldv | a->ldo) % 8 != 0) return MST_ERR_BAD_ARG;
  if (a->shift < 0 || a->shift >= a->ws) return MST_ERR_BAD_ARG;
  const WinGeom g = make_geom(a->H, a->W, a->ws, a->shift);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->ws == 8) return launch_attn<8>(*a, g, st);
  if (a->ws == 7) return launch_attn<7>(*a, g, st);
  return MST_ERR_UNSUPPORTED;
}

extern "C" int mst_window_maps(int H, int W, int ws, int shift, int32_t* gather, int32_t* labels, int32_t* relidx,
                               void* stream) {
  using namespace mst;
  if (H <= 0 || W <= 0 || ws <= 0 || shift < 0) return MST_ERR_BAD_ARG;
  const WinGeom g = make_geom(H, W, ws, shift);
  const int N = ws * ws;
  const int total = max(g.nW * N, N * N);
  window_maps_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, gather, labels, relidx);
  return (int)cudaGetLastError();
}
