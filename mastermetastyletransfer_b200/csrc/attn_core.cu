// Shifted-window attention CORE on tcgen05 for the two passes of a style-transformer layer whose operands are projections of
// DIFFERENT tensors (so the fused projection + attention kernel of attn_fused.cu does not apply):
//   * StyleEncoder Scale / Shift pass: q, k from Key', v = Wv Scale, v2 = Wv Shift  (codes/style_transformer.py:127-155, :866-882)
//   * StyleDecoder sigma / mu attention: q = IN(Query), k = IN(Wk Key), v = Wvs Scale, v2 = Wvh Shift  (:544-594)
// Both share ONE softmax between two value tensors ("dual"): out = softmax(q k^T * d^-1/2 + bias + mask) v, out2 = (same P) v2.
// It replaces the mma.sync kernel of window_attn.cu on this path (55 TFLOP/s, 58 us per launch at batch 32) whenever the shape
// allows: dual, head_dim 32, an even number of heads, windows of 7 or 8 that tile the map (no zero-padded tokens).
//
// Same tile as attn_fused.cu: two windows = 128 slot rows, a CTA owns one head pair and walks tiles with a static stride.
//   warps 8-9  producers: the pair's 64-channel slice (128 B per token) of q, k, v, v2 for the tile's two windows: one
//              cp.async.bulk.tensor.4d per tensor and window (box [64 ch x ws x ws] at the window's rolled coordinates:
//              partition and cyclic shift are the load's coordinates), row-gathered cp.async for windows on the wrap-around edge;
//              two tile buffers, so tile t+1 loads under the softmax of tile t
//   warps 0-7  four per head: S_p = Q_p K_p^T (K = 32: both operands are the same [128 x 128 B] swizzled tiles at a 64 p byte
//              offset), softmax of one query row per thread out of TMEM (relative-position bias, closed-form shift mask), P back
//              into TMEM as the A operand of  O_p = P V_p  and  O2_p = P V2_p  (V tiles token-major, read MN-major), 1/rowsum,
//              bf16, stored to the row's source token (window reverse + roll back).  The head's leader warp issues the MMAs.
// TMEM: S_p / P_p / O_p in columns 256 + 128 p .. (P: 0..31, O: 64..127 of the region), O2_p in columns 128 p .. 128 p + 63.
#include "../../include/mst_b200.h"
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace mst {

constexpr int AC_EPI_WARPS = 8;
constexpr int AC_PROD_WARP0 = 8;
constexpr int AC_PROD_THREADS = 64;
constexpr int AC_THREADS = (AC_EPI_WARPS + 2) * 32;  // 320
constexpr int AC_TILE_BYTES = 128 * 128;             // one tensor's [128 rows x 128 B] tile
constexpr int AC_BUF_BYTES = 4 * AC_TILE_BYTES;      // q | k | v | v2
constexpr int AC_SMEM_BYTES = 1024 + 2 * AC_BUF_BYTES;
constexpr int AC_COL_S = 256;

MST_DEVINL void ac_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
MST_DEVINL void ac_tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
               : "memory");
}
MST_DEVINL float ac_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
MST_DEVINL uint32_t ac_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
MST_DEVINL void ac_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// MN-major SWIZZLE_128B descriptor (token-major value tile: a key's 64 dims are contiguous); see attn_fused.cu / wgrad_tc.cu
MST_DEVINL uint64_t ac_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16384 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct AcArgs {
  const bf16* src[4];  // q, k, v, v2
  int ld[4];
  bf16* out; bf16* out2;
  const float* bias_table;
  int B, H, W, heads, ldo;
};
struct AcMaps { CUtensorMap m[4]; };

template <int WS>
__global__ void __launch_bounds__(AC_THREADS, 1) attn_core_kernel(const AcArgs p, const WinGeom g, const __grid_constant__ AcMaps tm,
                                                                  const int use_tma, const int n_tiles, const int total_windows) {
  constexpr int N = WS * WS;
  constexpr int NTAB = 2 * WS - 1;
  constexpr int NT = NTAB * NTAB;
  constexpr float LOG2E = 1.4426950408889634f;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t x_full[2], x_empty[2], s_full[2], o_full[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float table_s[2 * NT];

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform for ptxas
  const uint32_t x_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nhp = p.heads >> 1;
  const int hp = blockIdx.x % nhp;     // this CTA's head pair
  const int first = blockIdx.x / nhp;  // its first tile
  const int stride = gridDim.x / nhp;
  const int n_my = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;
  const int nWy = g.Hp / WS;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&x_full[b]), AC_PROD_THREADS + 1);
      mbar_init(smem_u32(&x_empty[b]), 2);  // both heads' leaders, after their PV MMAs
      mbar_init(smem_u32(&s_full[b]), 1);
      mbar_init(smem_u32(&o_full[b]), 1);
    }
    mbar_fence_init();
  }
  // bias table of the pair's two heads, [head][entry], pre-multiplied by log2(e) (the softmax runs in base 2)
  for (int i = threadIdx.x; i < 2 * NT; i += AC_THREADS) {
    const int pp = i / NT, idx = i - pp * NT;
    table_s[i] = p.bias_table[idx * p.heads + hp * 2 + pp] * LOG2E;
  }
  // tiles start as zeros: slot rows past a 7x7 window are never written (TMA lands 49 rows), and they must stay finite
  for (uint32_t i = threadIdx.x; i < (uint32_t)(2 * AC_BUF_BYTES) / 16u; i += AC_THREADS) ac_sts128(x_base + i * 16u, 0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == AC_PROD_WARP0) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= AC_PROD_WARP0) {
    // =========================== producers ===========================
    const int pt = threadIdx.x - AC_PROD_WARP0 * 32;  // 0..63
    const int c = pt & 7, r0 = pt >> 3;
    int siy[8], six[8];  // this thread's eight slot rows of a gathered window: in-window position, -1 past a 7x7 window
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int slot = r0 + 8 * i;
      siy[i] = slot < N ? slot / WS : -1;
      six[i] = slot < N ? slot - (slot / WS) * WS : 0;
    }
    const unsigned long long magic_nwx_p = (1ull << 32) / (unsigned)g.nwx + 1ull;  // win / nwx without a run-time division
    const unsigned long long magic_nw_p = (1ull << 32) / (unsigned)g.nW + 1ull;
    for (int lt = 0; lt < n_my; ++lt) {
      const int tile = first + lt * stride;
      const int buf = lt & 1, u = lt >> 1;
      if (u >= 1) mbar_wait(smem_u32(&x_empty[buf]), (u - 1) & 1);
      const uint32_t xb = x_base + buf * AC_BUF_BYTES;
      const uint32_t bar = smem_u32(&x_full[buf]);
      int bw[2], wyv[2], wxv[2];
      bool tma_w[2], ok_w[2];
      uint32_t tx = 0;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int wg = tile * 2 + w;
        ok_w[w] = wg < total_windows;
        bw[w] = ok_w[w] ? (g.nW == 1 ? wg : (int)(((unsigned long long)(unsigned)wg * magic_nw_p) >> 32)) : 0;
        if (ok_w[w] && (bw[w] + 1) * g.nW <= wg) ++bw[w];
        const int win = ok_w[w] ? wg - bw[w] * g.nW : 0;
        wyv[w] = (int)(((unsigned long long)(unsigned)win * magic_nwx_p) >> 32);
        wxv[w] = win - wyv[w] * g.nwx;
        const bool wrapped = (g.sy > 0 && wyv[w] == nWy - 1) || (g.sx > 0 && wxv[w] == g.nwx - 1);
        tma_w[w] = ok_w[w] && use_tma && !wrapped;
        if (tma_w[w]) tx += (uint32_t)(4 * N * 128);
      }
      if (pt == 0) {
        if (tx) ac_arrive_expect_tx(bar, tx); else mbar_arrive(bar);
      }
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        if (!ok_w[w]) continue;
        if (tma_w[w]) {
          if (pt == 0) {
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
              ac_tma_load_4d(xb + kb * AC_TILE_BYTES + w * 8192, &tm.m[kb], hp * 64, wxv[w] * WS + g.sx, wyv[w] * WS + g.sy, bw[w], bar);
          }
        } else {
          const int y0 = wyv[w] * WS + g.sy, x0 = wxv[w] * WS + g.sx;
          const long long img = (long long)bw[w] * g.H;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int slot = r0 + 8 * i;
            long long src = -1;
            if (siy[i] >= 0) {
              int y = y0 + siy[i], x = x0 + six[i];
              if (y >= g.Hp) y -= g.Hp;
              if (x >= g.Wp) x -= g.Wp;
              if (y < g.H && x < g.W) src = (img + y) * g.W + x;
            }
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
              const bf16* sp = src >= 0 ? p.src[kb] + src * p.ld[kb] + hp * 64 + c * 8 : p.src[kb];
              cp_async16(xb + kb * AC_TILE_BYTES + sw128_offset(w * 64 + slot, c), sp, src >= 0);
            }
          }
        }
      }
      cp_async_mbar_arrive_noinc(bar);
    }
    cp_async_wait_all();
  } else {
    // =========================== attention warps 0-7 ===========================
    const int quad = warp & 3, part = warp >> 2;
    const int r = quad * 32 + lane;  // row of the 128-row tile
    const int w = quad >> 1;         // its window (warp-uniform)
    const int slot = r & 63;         // its slot inside the window
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int head = hp * 2 + part;
    const int iy = slot < N ? slot / WS : 0, ix = slot < N ? slot - (slot / WS) * WS : 0;
    constexpr int RP_PAD = (WS - 1) * NTAB + WS - 1;
    const int rp = slot < N ? (iy + WS - 1) * NTAB + ix + WS - 1 : RP_PAD;
    const float* tab = table_s + part * NT + rp;
    const float scale2 = 0.17677669529663687f * LOG2E;  // head_dim^-0.5 (the reference scales q, :127), base-2 softmax
    const float MASKV = -100.0f * LOG2E;
    const uint32_t s_addr = lane_addr + AC_COL_S + part * 128;
    const uint32_t s_addr_mma = tmem_base + AC_COL_S + part * 128;
    const uint32_t o2_addr = lane_addr + part * 128, o2_addr_mma = tmem_base + part * 128;
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 32) | (1u << 16);  // B (a value tile) is MN-major
    const int step_w = 2 * stride;
    const int db = step_w / g.nW, dwin = step_w - db * g.nW;
    int wg = first * 2 + w;
    int b = wg / g.nW, win = wg - b * g.nW;
    const unsigned long long magic_nwx = (1ull << 32) / (unsigned)g.nwx + 1ull;

    for (int lt = 0; lt < n_my; ++lt) {
      // ---- this row's place in the feature map and the window's shift mask (integer, bit-exact; as attn_fused.cu) ----
      const bool valid = wg < total_windows;
      const int wy = (int)(((unsigned long long)(unsigned)win * magic_nwx) >> 32), wx = win - wy * g.nwx;
      long long src = -2;
      if (valid && slot < N) {
        int y = wy * WS + iy + g.sy, x = wx * WS + ix + g.sx;
        if (y >= g.Hp) y -= g.Hp;
        if (x >= g.Wp) x -= g.Wp;
        src = (y < g.H && x < g.W) ? ((long long)b * g.H + y) * g.W + x : -1;
      }
      const bool ywrap = g.sy > 0 && wy == nWy - 1, xwrap = g.sx > 0 && wx == g.nwx - 1;
      const bool masked = valid && (ywrap || xwrap);
      uint32_t m_lo = 0, m_hi = 0;
      if (masked) {
        constexpr unsigned long long ALL = N == 64 ? ~0ull : ((1ull << (N & 63)) - 1ull);
        unsigned long long rep = 0;
#pragma unroll
        for (int rr = 0; rr < WS; ++rr) rep |= 1ull << (rr * WS);
        const int ty = ywrap ? WS - g.sy : WS, tx = xwrap ? WS - g.sx : WS;
        const unsigned long long my = ty >= WS ? 0ull : (ALL >> (ty * WS)) << (ty * WS);
        const unsigned long long mxm = (unsigned long long)(((1u << WS) - 1u) & ~((1u << tx) - 1u)) * rep;
        const unsigned long long m = ((iy >= ty ? ~my : my) | (ix >= tx ? ~mxm : mxm)) & ALL;
        m_lo = (uint32_t)m;
        m_hi = (uint32_t)(m >> 32);
      }
      wg += step_w; b += db; win += dwin;
      if (win >= g.nW) { win -= g.nW; ++b; }

      // ---- (1) the tile's q | k | v | v2 have landed; the head's leader issues S_p = Q_p K_p^T (two windows stacked) ----
      const int buf = lt & 1, u = lt >> 1;
      const uint32_t xb = x_base + buf * AC_BUF_BYTES;
      if (lane == 0) mbar_wait(smem_u32(&x_full[buf]), u & 1);
      __syncwarp();
      fence_proxy_async_smem();  // (gathered windows were written by cp.async)
      // all four warps of the head have finished reading O / O2 of the previous tile before S overwrites the region
      asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
      if (quad == 0) {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16_pred(s_addr_mma, umma_desc_sw128(xb + part * 64 + k * 32), umma_desc_sw128(xb + AC_TILE_BYTES + part * 64 + k * 32), idesc_s,
                         k != 0);
        umma_commit_pred(smem_u32(&s_full[part]));
      }

      // ---- (2) softmax of this thread's query row; P back into tensor memory ----
      if (lane == 0) mbar_wait(smem_u32(&s_full[part]), lt & 1);
      __syncwarp();
      tc_fence_after();
      float inv;
      {
        uint32_t sa[32], sb[32];
        tmem_ld32(s_addr + w * 64, sa);
        tmem_ld32(s_addr + w * 64 + 32, sb);
        tmem_wait_ld();
        float s[64];
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (masked) {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            if (j < N) {
              const int cp = (j / WS) * NTAB + (j % WS);
              float t = fmaf(__uint_as_float(j < 32 ? sa[j & 31] : sb[j & 31]), scale2, tab[-cp]);
              const uint32_t mbit = j < 32 ? (m_lo >> (j & 31)) : (m_hi >> (j & 31));
              if (mbit & 1u) t += MASKV;
              s[j] = t;
              mxa[j & 3] = fmaxf(mxa[j & 3], t);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            if (j < N) {
              const int cp = (j / WS) * NTAB + (j % WS);
              const float t = fmaf(__uint_as_float(j < 32 ? sa[j & 31] : sb[j & 31]), scale2, tab[-cp]);
              s[j] = t;
              mxa[j & 3] = fmaxf(mxa[j & 3], t);
            }
          }
        }
        const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
        float suma[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int h16 = 0; h16 < 2; ++h16) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j0 = h16 * 32 + e * 2;
            const float p0 = j0 < N ? ac_ex2(s[j0 < N ? j0 : 0] - mx) : 0.f;  // keys past a 7x7 window: probability 0
            const float p1 = j0 + 1 < N ? ac_ex2(s[j0 + 1 < N ? j0 + 1 : 0] - mx) : 0.f;
            suma[e & 3] += p0 + p1;
            pk[e] = ac_pack(p0, p1);
          }
          tmem_st16(s_addr + h16 * 16, pk);
        }
        inv = 1.0f / ((suma[0] + suma[1]) + (suma[2] + suma[3]));
      }
      tmem_wait_st();
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
      if (quad == 0) {  // O_p[:, 32w..] = P_p V_{p,w} and the same with V2: 16 keys per step, P from TMEM, values MN-major
        tc_fence_after();
#pragma unroll
        for (int ww = 0; ww < 2; ++ww)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_ts_pred(s_addr_mma + 64 + ww * 32, s_addr_mma + k * 8, ac_desc_mn_sw128(xb + 2 * AC_TILE_BYTES + ww * 8192 + part * 64 + k * 2048), idesc_pv,
                         k != 0);
            umma_ts_pred(o2_addr_mma + ww * 32, s_addr_mma + k * 8, ac_desc_mn_sw128(xb + 3 * AC_TILE_BYTES + ww * 8192 + part * 64 + k * 2048), idesc_pv,
                         k != 0);
          }
        umma_commit_pred(smem_u32(&o_full[part]));
        umma_commit_pred(smem_u32(&x_empty[buf]));  // this head is done with the tile's operands
      }

      // ---- (3) O, O2 -> * 1/rowsum -> bf16 -> the row's source token ----
      if (lane == 0) mbar_wait(smem_u32(&o_full[part]), lt & 1);
      __syncwarp();
      tc_fence_after();
      {
        uint32_t ov[32], ov2[32];
        tmem_ld32(s_addr + 64 + w * 32, ov);
        tmem_ld32(o2_addr + w * 32, ov2);
        tmem_wait_ld();
        tc_fence_before();
        if (src >= 0) {
          bf16* op = p.out + src * p.ldo + head * 32;
          bf16* op2 = p.out2 + src * p.ldo + head * 32;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t pk[8], pk2[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              pk[e] = ac_pack(__uint_as_float(ov[h2 * 16 + 2 * e]) * inv, __uint_as_float(ov[h2 * 16 + 2 * e + 1]) * inv);
              pk2[e] = ac_pack(__uint_as_float(ov2[h2 * 16 + 2 * e]) * inv, __uint_as_float(ov2[h2 * 16 + 2 * e + 1]) * inv);
            }
            st_global_256(op + h2 * 16, pk);
            st_global_256(op2 + h2 * 16, pk2);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == AC_PROD_WARP0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*AcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static AcEncodeTiledFn ac_tma_encoder() {
  static AcEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<AcEncodeTiledFn>(q);
  }
  return fn;
}

static int ac_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

template <int WS>
static int launch_core(const MstWindowAttn& a, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_core_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, AC_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const WinGeom g = make_geom(a.H, a.W, a.ws, a.shift);
  const long long total = (long long)a.B * g.nW;
  if (total <= 0 || total > 0x3fffffffLL) return MST_ERR_BAD_ARG;
  const int n_tiles = (int)((total + 1) / 2);
  const int nhp = a.heads / 2;
  long long grid = (long long)n_tiles * nhp;
  const long long cap = (long long)(ac_num_sms() / nhp) * nhp;
  if (grid > cap) grid = cap;
  AcArgs args;
  const mst_bf16* srcs[4] = {a.q, a.k, a.v, a.v2};
  const int lds[4] = {a.ldq, a.ldk, a.ldv, a.ldv};
  alignas(64) AcMaps maps;
  memset(&maps, 0, sizeof(maps));
  int use_tma = 0;
  AcEncodeTiledFn enc = ac_tma_encoder();
  if (enc) use_tma = 1;
  const int C = a.heads * 32;
  for (int i = 0; i < 4; ++i) {
    args.src[i] = reinterpret_cast<const bf16*>(srcs[i]);
    args.ld[i] = lds[i];
    if (use_tma) {  // the tensor as [C, W, H, B]: a window that does not wrap is one [64 x ws x ws] box per head pair
      const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
      const cuuint64_t gstride[3] = {(cuuint64_t)lds[i] * 2, (cuuint64_t)a.W * lds[i] * 2, (cuuint64_t)a.H * a.W * lds[i] * 2};
      const cuuint32_t box[4] = {64, (cuuint32_t)WS, (cuuint32_t)WS, 1};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      if (enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(reinterpret_cast<const void*>(srcs[i])), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        use_tma = 0;
    }
  }
  args.out = reinterpret_cast<bf16*>(a.out);
  args.out2 = reinterpret_cast<bf16*>(a.out2);
  args.bias_table = a.bias_table;
  args.B = a.B; args.H = a.H; args.W = a.W; args.heads = a.heads; args.ldo = a.ldo;
  attn_core_kernel<WS><<<(unsigned)grid, AC_THREADS, AC_SMEM_BYTES, st>>>(args, g, maps, use_tma, n_tiles, (int)total);
  return (int)cudaGetLastError();
}

// Called first by mst_window_attention (window_attn.cu): handled = false leaves the call to the general kernel.
int attn_core_try(const MstWindowAttn& a, cudaStream_t st, bool& handled) {
  handled = false;
  static int allow = -1;
  if (allow < 0) { const char* e = getenv("MST_ATTN_CORE"); allow = e ? atoi(e) : 1; }  // 0: always the general kernel (experiments)
  if (!allow) return 0;
  if (!a.v2 || !a.out2) return 0;                                   // dual passes only
  if (a.ws != 7 && a.ws != 8) return 0;
  if (a.H % a.ws != 0 || a.W % a.ws != 0) return 0;                 // zero-padded tokens (pad_q ...): the general kernel
  if (a.heads < 2 || (a.heads & 1) || a.heads > 32) return 0;
  if ((a.ldq | a.ldk | a.ldv | a.ldo) % 16 != 0) return 0;          // 32-byte row segments (256-bit stores), 16-byte TMA strides
  const uintptr_t al = reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v) |
                       reinterpret_cast<uintptr_t>(a.v2) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.out2);
  if (al & 31) return 0;
  handled = true;
  return a.ws == 8 ? launch_core<8>(a, st) : launch_core<7>(a, st);
}

}  // namespace mst
