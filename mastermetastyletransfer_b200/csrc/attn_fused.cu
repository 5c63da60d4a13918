// Fused self-attention half of a (shifted-)window transformer block for sm_100a:
//
//     out = window_attention(x Wq^T + bq, x Wk^T + bk, x Wv^T + bv)          (before the output projection)
//
// i.e. codes/style_transformer.py:77-155 (pad -> roll -> partition -> three linears -> QK^T + relative-position bias + shift
// mask -> softmax -> PV -> window reverse -> roll back -> unpad) and the torchvision Swin block's attention (tv
// swin_transformer.py:116-220) in ONE kernel: q, k, v never exist in HBM.  The only global traffic is one read of the bf16
// token tile (x, already LayerNorm'd where the block has a norm1) and one write of the bf16 attention output, which the fused
// projection + LayerNorm + MLP kernel (mlp_fused.cu, PRE) consumes.
//
// Work decomposition.  A tile = TWO windows = 128 slot rows (7x7 windows use 49 of their 64 slots; rows past the window and
// zero-padded tokens are zero rows of x, so their q/k/v are the biases exactly as the reference's pad-then-linear order makes
// them).  A CTA owns ONE HEAD PAIR for its whole life: the [192 x C] slice of Wq|Wk|Wv of those two heads stays resident in
// shared memory (48 / 96 KB) and the CTA walks tiles with a static stride; the heads/2 CTAs working on the same tile run side
// by side, so x is read from HBM once and from L2 otherwise.
//
//   warps 0-7 : "epilogue" warps, warp = (TMEM lane quadrant q = w & 3 -> rows 32q..32q+31, head p = w >> 2 of the pair).
//               The two heads are INDEPENDENT chains (own accumulators, tiles and mbarriers): while one head's warps wait for
//               a tensor-core hand-off, the other head's warps (same schedulers) issue.
//               (1) drain the head's QKV accumulator: + bias -> bf16 -> the head's [128 x (32 q | 32 k)] 128B-swizzled K-major
//                   tile and its 32 columns of the token-major V tile (read by the tensor core as an MN-major operand);
//               (2) softmax of one query row per thread straight out of TMEM: scale, relative-position bias (gathered from the
//                   225-entry table in shared memory), 9-region shift mask (bit-exact labels, common.cuh), base-2 exponentials,
//                   un-normalised probabilities -> bf16 -> swizzled P tile (it overwrites Q / K, which are dead by then);
//               (3) O = P V out of TMEM, * 1/rowsum -> bf16 -> global, window reverse + roll back in the store address.
//   warp 8    : TMEM allocation + tcgen05.mma issue (warp-uniform loop, one elected lane):
//                 QKV_p[128 x 96] = X[128 x C] . W_p^T        (N = 96 per head, K = C)
//                 S_p[128 x 128]  = Q_p . K_p^T               (two windows stacked: the diagonal 64x64 blocks are used)
//                 O_p[:, 32w..]  = P_p[128 x 64] . V_{p,w}    (N = 32, K = 64, per window w; rows of the other window ignored)
//               and QKV of the NEXT tile is issued right after S, so it runs under the softmax of this one.
//   warps 9-10: x-tile producers.  Windows that do not wrap around the rolled map are ONE cp.async.bulk.tensor (TMA) per
//               64-channel k-block: box [64 ch x ws x ws] of the [C, W, H, B] tensor map at (x0 + shift, y0 + shift) -- the
//               partition and the cyclic shift are the box coordinates, the zero padding of 7x7 windows is the hardware's
//               out-of-bounds fill.  Windows on the wrap-around edge (last window row / column of a shifted map) are gathered
//               row by row with 16-byte cp.async.
//
// TMEM (512 columns): QKV_0 0..95, QKV_1 128..223, S_0 256..383, S_1 384..511; O_p overwrites the first 64 columns of S_p.
// Shared memory: weights KB x 24 KB | x tile (double buffered for C = 128) | QK_0 (later P_0) | QK_1 (later P_1) | V.
#include "../../include/mst_b200.h"
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace mst {

constexpr int AF_EPI_WARPS = 8;
constexpr int AF_MMA_WARP = 8;
constexpr int AF_PROD_WARP0 = 9;
constexpr int AF_PROD_WARPS = 2;
constexpr int AF_PROD_THREADS = AF_PROD_WARPS * 32;
constexpr int AF_THREADS = (AF_EPI_WARPS + 1 + AF_PROD_WARPS) * 32;  // 352
constexpr int AF_WROWS = 192;                  // (q | k | v) x 32 dims x 2 heads
constexpr int AF_WKB_BYTES = AF_WROWS * 128;   // one 64-channel k-block of the pair's weights
constexpr int AF_XKB_BYTES = 128 * 128;        // one 64-channel k-block of a 128-row tile
constexpr int AF_COL_ACC = 128;                // QKV accumulator of head p at TMEM column 128 p (96 used)
constexpr int AF_COL_S = 256;                  // S_p at TMEM column 256 + 128 p

MST_DEVINL void af_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
MST_DEVINL void af_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
MST_DEVINL void af_tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
               : "memory");
}
MST_DEVINL float af_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
MST_DEVINL uint32_t af_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
MST_DEVINL void af_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// MN-major SWIZZLE_128B shared-memory descriptor (the V tile is token-major: a key's 64 dims are contiguous): SBO = 1024 B
// between 8-key groups, LBO = distance between 64-element groups along N (not reached: N = 32).  See wgrad_tc.cu.
MST_DEVINL uint64_t af_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16384 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct AttnCore {
  const bf16* x; const uint8_t* wqkv; const float* bqkv; const float* bias_table; bf16* out; bf16* dbg_qkv;
  int B, H, W, heads, ldx, ldo;
};

template <int C>
struct AttnCfg {
  static constexpr int KB = C / 64;
  static constexpr int XB = C == 128 ? 2 : 1;            // x-tile buffers (C = 256: the weights take 96 KB)
  static constexpr int W_BYTES = KB * AF_WKB_BYTES;
  static constexpr int X_BYTES = KB * AF_XKB_BYTES;
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + XB * X_BYTES + 3 * 16384;
};

template <int C, int WS>
__global__ void __launch_bounds__(AF_THREADS, 1) attn_fused_kernel(const AttnCore p, const WinGeom g, const __grid_constant__ CUtensorMap tm,
                                                                   const int use_tma, const int n_tiles, const int total_windows, const int stagger) {
  using Cfg = AttnCfg<C>;
  constexpr int KB = Cfg::KB, XB = Cfg::XB;
  constexpr int N = WS * WS;
  constexpr int NTAB = 2 * WS - 1;
  constexpr int NT = NTAB * NTAB;
  constexpr float LOG2E = 1.4426950408889634f;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full, x_full[2], x_empty[2], acc_full[2], qkv_ready[2], s_full[2], p_ready[2], o_full[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[AF_WROWS];
  __shared__ float table_s[2 * NT];

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform for ptxas
  const uint32_t w_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_base = w_base + Cfg::W_BYTES;
  const uint32_t qk_base = x_base + XB * Cfg::X_BYTES;  // head p: [128 rows x (32 q dims | 32 k dims)] at qk_base + 16 KB p; later P_p
  const uint32_t v_base = qk_base + 2 * 16384;          // V tile [128 keys x 64 dims] token-major (head p = columns 32p..)

  const int nhp = p.heads >> 1;
  const int hp = blockIdx.x % nhp;             // this CTA's head pair
  const int first = blockIdx.x / nhp;          // its first tile
  const int stride = gridDim.x / nhp;
  const int n_my = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;
  const int nWy = g.Hp / WS;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&w_full), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&x_full[b]), AF_PROD_THREADS + 1);
      mbar_init(smem_u32(&x_empty[b]), 1);
      mbar_init(smem_u32(&p_ready[b]), AF_EPI_WARPS / 2);
      mbar_init(smem_u32(&o_full[b]), 1);
      mbar_init(smem_u32(&acc_full[b]), 1);
      mbar_init(smem_u32(&qkv_ready[b]), AF_EPI_WARPS / 2);
      mbar_init(smem_u32(&s_full[b]), 1);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < AF_WROWS; i += AF_THREADS) bias_s[i] = p.bqkv[hp * AF_WROWS + i];
  // bias table of the pair's two heads, [head][entry], pre-multiplied by log2(e) (the softmax runs in base 2)
  for (int i = threadIdx.x; i < 2 * NT; i += AF_THREADS) {
    const int pp = i / NT, idx = i - pp * NT;
    table_s[i] = p.bias_table[idx * p.heads + hp * 2 + pp] * LOG2E;
  }
  // x tiles start as zeros: slot rows past a 7x7 window are never written (TMA lands 49 rows), and they must stay finite
  for (uint32_t i = threadIdx.x; i < (uint32_t)(XB * Cfg::X_BYTES) / 16u; i += AF_THREADS) af_sts128(x_base + i * 16u, 0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == AF_MMA_WARP) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= AF_PROD_WARP0) {
    // =========================== x-tile producers ===========================
    const int pt = threadIdx.x - AF_PROD_WARP0 * 32;  // 0..63
    if (pt == 0) {  // the head pair's weights: resident for the whole kernel
      af_arrive_expect_tx(smem_u32(&w_full), Cfg::W_BYTES);
      const uint8_t* wsrc = p.wqkv + (size_t)hp * Cfg::W_BYTES;
      for (int kb = 0; kb < KB; ++kb) af_bulk_g2s(w_base + kb * AF_WKB_BYTES, wsrc + (size_t)kb * AF_WKB_BYTES, AF_WKB_BYTES, smem_u32(&w_full));
    }
    const int c = pt & 7, r0 = pt >> 3;
    // this thread's eight slot rows of a gathered window (r0 + 8 i): in-window position, or -1 for slots past a 7x7 window
    int siy[8], six[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int slot = r0 + 8 * i;
      siy[i] = slot < N ? slot / WS : -1;
      six[i] = slot < N ? slot - (slot / WS) * WS : 0;
    }
    const unsigned long long magic_nwx_p = (1ull << 32) / (unsigned)g.nwx + 1ull;  // win / nwx without a run-time division
    const unsigned long long magic_nw_p = (1ull << 32) / (unsigned)g.nW + 1ull;    // wg / nW likewise (wg < 2^16 * nW is far away)
    for (int lt = 0; lt < n_my; ++lt) {
      const int tile = first + lt * stride;
      const int buf = lt % XB, u = lt / XB;
      if (u >= 1) mbar_wait(smem_u32(&x_empty[buf]), (u - 1) & 1);
      const uint32_t xb = x_base + buf * Cfg::X_BYTES;
      const uint32_t bar = smem_u32(&x_full[buf]);
      // pass 1: which windows go through TMA (thread 0 announces the bytes before issuing anything)
      int bw[2], wyv[2], wxv[2];
      bool tma_w[2], ok_w[2];
      uint32_t tx = 0;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int wg = tile * 2 + w;
        ok_w[w] = wg < total_windows;
        bw[w] = ok_w[w] ? (g.nW == 1 ? wg : (int)(((unsigned long long)(unsigned)wg * magic_nw_p) >> 32)) : 0;
        if (ok_w[w] && (bw[w] + 1) * g.nW <= wg) ++bw[w];  // (the magic quotient can be one short for very large wg)
        const int win = ok_w[w] ? wg - bw[w] * g.nW : 0;
        wyv[w] = (int)(((unsigned long long)(unsigned)win * magic_nwx_p) >> 32);
        wxv[w] = win - wyv[w] * g.nwx;
        const bool wrapped = (g.sy > 0 && wyv[w] == nWy - 1) || (g.sx > 0 && wxv[w] == g.nwx - 1);
        tma_w[w] = ok_w[w] && use_tma && !wrapped;
        if (tma_w[w]) tx += (uint32_t)(KB * N * 128);
      }
      if (pt == 0) {
        if (tx) af_arrive_expect_tx(bar, tx); else mbar_arrive(bar);
      }
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        if (!ok_w[w]) continue;  // (odd window count: the tile's second half keeps whatever finite rows it holds)
        if (tma_w[w]) {
          if (pt == 0) {
            for (int kb = 0; kb < KB; ++kb)
              af_tma_load_4d(xb + kb * AF_XKB_BYTES + w * 8192, &tm, kb * 64, wxv[w] * WS + g.sx, wyv[w] * WS + g.sy, bw[w], bar);
          }
        } else {
          // wrap-around window (or no TMA): gathered row by row, 16 bytes per cp.async; win_source() of common.cuh with the
          // window side as a compile-time constant
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
            const bf16* sp = src >= 0 ? p.x + src * p.ldx + c * 8 : p.x;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
              cp_async16(xb + kb * AF_XKB_BYTES + sw128_offset(w * 64 + slot, c), src >= 0 ? sp + kb * 64 : sp, src >= 0);
          }
        }
      }
      cp_async_mbar_arrive_noinc(bar);
    }
    cp_async_wait_all();
  } else if (warp == AF_MMA_WARP) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc_qkv = umma_idesc_bf16(128, 96);
    mbar_wait(smem_u32(&w_full), 0);
    tc_fence_after();
    // This warp issues only the projections: QKV_p(t) as soon as the x tile has landed and head p's accumulator has been
    // drained (qkv_ready[p] of tile t-1).  S_p and PV_p are issued by the head's OWN epilogue warps (its leader warp, right after
    // a named barrier of the head's four warps): a hand-off through this warp costs an mbarrier round trip and a poll, twice per
    // tile and head, on a chain that is latency-bound.
    for (int t = 0; t < n_my; ++t) {
      const int buf = t % XB, u = t / XB;
      mbar_wait(smem_u32(&x_full[buf]), u & 1);
#pragma unroll 1
      for (int pp = 0; pp < 2; ++pp) {
        if (t > 0) mbar_wait(smem_u32(&qkv_ready[pp]), (t - 1) & 1);
        tc_fence_after();
        const uint32_t xb = x_base + buf * Cfg::X_BYTES;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)  // head pp: weight rows 96 pp .. 96 pp + 95 of every k-block
            umma_bf16_pred(tmem_base + pp * AF_COL_ACC, umma_desc_sw128(xb + kb * AF_XKB_BYTES + k * 32),
                           umma_desc_sw128(w_base + kb * AF_WKB_BYTES + pp * 12288 + k * 32), idesc_qkv, (kb | k) != 0);
        if (pp == 1) umma_commit_pred(smem_u32(&x_empty[buf]));  // both heads have read this x tile
        umma_commit_pred(smem_u32(&acc_full[pp]));
      }
    }
    tc_fence_before();
  } else {
    // =========================== epilogue warps 0-7 ===========================
    const int quad = warp & 3, part = warp >> 2;
    const int r = quad * 32 + lane;   // row of the 128-row tile
    const int w = quad >> 1;          // its window (warp-uniform)
    const int slot = r & 63;          // its slot inside the window
    const uint32_t xrow = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int head = hp * 2 + part;
    const int iy = slot < N ? slot / WS : 0, ix = slot < N ? slot - (slot / WS) * WS : 0;
    constexpr int RP_PAD = (WS - 1) * NTAB + WS - 1;  // rows past the window read the table's centre entry (discarded)
    const int rp = slot < N ? (iy + WS - 1) * NTAB + ix + WS - 1 : RP_PAD;
    const float* tab = table_s + part * NT + rp;
    const float scale2 = 0.17677669529663687f * LOG2E;  // head_dim^-0.5 (the reference scales q, :127), base-2 softmax
    const float MASKV = -100.0f * LOG2E;
    const uint32_t qk_tile = qk_base + part * 16384;   // this head's q | k tile, later its P tile
    const uint32_t acc_addr = lane_addr + part * AF_COL_ACC;
    const uint32_t s_addr = lane_addr + AF_COL_S + part * 128;
    const uint32_t s_addr_mma = tmem_base + AF_COL_S + part * 128;  // the accumulator address the head's leader warp issues to (lane 0)
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 32) | (1u << 16);  // B (the V tile) is MN-major
    // window index of this warp's half-tile, advanced incrementally (no per-tile divisions by run-time values)
    const int step_w = 2 * stride;
    const int db = step_w / g.nW, dwin = step_w - db * g.nW;
    int wg = first * 2 + w;
    int b = wg / g.nW, win = wg - b * g.nW;
    const unsigned long long magic_nwx = (1ull << 32) / (unsigned)g.nwx + 1ull;  // win / nwx = (win * magic) >> 32 for win < 2^16

#ifdef MST_AF_PROF
    long long tB = 0, tAw = 0, tD = 0, tSw = 0, tS = 0, tOw = 0, tO = 0, t_all = clock64(), tm = clock64();
#define AF_MARK(acc) { long long n_ = clock64(); acc += n_ - tm; tm = n_; }
#else
#define AF_MARK(acc)
#endif
    for (int lt = 0; lt < n_my; ++lt) {
      // ---- this row's place in the feature map, and the window's shift mask (integer, bit-exact: common.cuh) ----
      const bool valid = wg < total_windows;
      const int wy = (int)(((unsigned long long)(unsigned)win * magic_nwx) >> 32), wx = win - wy * g.nwx;
      long long src = -2;  // -2: no token (slot past the window / window past the end); -1: zero-padded token
      if (valid && slot < N) {  // win_source() of common.cuh with the window side as a compile-time constant
        int y = wy * WS + iy + g.sy, x = wx * WS + ix + g.sx;
        if (y >= g.Hp) y -= g.Hp;
        if (x >= g.Wp) x -= g.Wp;
        src = (y < g.H && x < g.W) ? ((long long)b * g.H + y) * g.W + x : -1;
      }
      // Shift mask (style_transformer.py:134-150).  win_band() of common.cuh labels a row p of the rolled map 0 / 1 / 2 for
      // p < Hp-ws / < Hp-s / else; windows are ws-aligned and Hp is a multiple of ws, so only the LAST window row (column) of a
      // shifted map holds two labels, split at in-window row (column) ws - s: key j is masked for this query row when they
      // lie on different sides of the row split or of the column split.  Closed form of the same integer arithmetic
      // (checked against the reference's own masks through the parity tests on every geometry).
      const bool ywrap = g.sy > 0 && wy == nWy - 1, xwrap = g.sx > 0 && wx == g.nwx - 1;
      const bool masked = valid && (ywrap || xwrap);
      uint32_t m_lo = 0, m_hi = 0;  // bit j: key j carries another region label than this row -> -100
      if (masked) {
        constexpr unsigned long long ALL = N == 64 ? ~0ull : ((1ull << (N & 63)) - 1ull);
        unsigned long long rep = 0;  // bit 0 of every in-window row
#pragma unroll
        for (int rr = 0; rr < WS; ++rr) rep |= 1ull << (rr * WS);
        const int ty = ywrap ? WS - g.sy : WS, tx = xwrap ? WS - g.sx : WS;
        const unsigned long long my = ty >= WS ? 0ull : (ALL >> (ty * WS)) << (ty * WS);                   // keys with yj >= ty
        const unsigned long long mxm = (unsigned long long)(((1u << WS) - 1u) & ~((1u << tx) - 1u)) * rep;  // keys with xj >= tx
        const unsigned long long m = ((iy >= ty ? ~my : my) | (ix >= tx ? ~mxm : mxm)) & ALL;
        m_lo = (uint32_t)m;
        m_hi = (uint32_t)(m >> 32);
      }
      wg += step_w; b += db; win += dwin;
      if (win >= g.nW) { win -= g.nW; ++b; }
      AF_MARK(tB)

      // ---- (1) this head's QKV accumulator -> + bias -> bf16 -> q | k tile and its V columns ----
      // (the tile doubles as P_p and the V columns are read by PV_p: both were released by o_full[part] of the previous tile,
      //  which this warp waited for in step (3))
      if (lane == 0) mbar_wait(smem_u32(&acc_full[part]), lt & 1);
      __syncwarp();
      AF_MARK(tAw)
      tc_fence_after();
      {
        uint32_t vq[32], vk[32];
        tmem_ld32(acc_addr, vq);
        tmem_ld32(acc_addr + 32, vk);
        tmem_wait_ld();
        const float4* b4 = reinterpret_cast<const float4*>(bias_s + part * 96);
        const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const float4 q0 = b4[2 * qd], q1 = b4[2 * qd + 1], k0 = b4[8 + 2 * qd], k1 = b4[8 + 2 * qd + 1];
          const int d = qd * 8;
          af_sts128(qk_tile + xrow + (((uint32_t)qd ^ sw) << 4),
                    af_pack(__uint_as_float(vq[d]) + q0.x, __uint_as_float(vq[d + 1]) + q0.y),
                    af_pack(__uint_as_float(vq[d + 2]) + q0.z, __uint_as_float(vq[d + 3]) + q0.w),
                    af_pack(__uint_as_float(vq[d + 4]) + q1.x, __uint_as_float(vq[d + 5]) + q1.y),
                    af_pack(__uint_as_float(vq[d + 6]) + q1.z, __uint_as_float(vq[d + 7]) + q1.w));
          af_sts128(qk_tile + xrow + (((uint32_t)(4 + qd) ^ sw) << 4),
                    af_pack(__uint_as_float(vk[d]) + k0.x, __uint_as_float(vk[d + 1]) + k0.y),
                    af_pack(__uint_as_float(vk[d + 2]) + k0.z, __uint_as_float(vk[d + 3]) + k0.w),
                    af_pack(__uint_as_float(vk[d + 4]) + k1.x, __uint_as_float(vk[d + 5]) + k1.y),
                    af_pack(__uint_as_float(vk[d + 6]) + k1.z, __uint_as_float(vk[d + 7]) + k1.w));
        }
        if (p.dbg_qkv && src >= 0) {  // test hook: the projected q | k | v rows in the layout of a fused-QKV GEMM output [T, 3C]
          bf16* dq = p.dbg_qkv + src * (3 * C) + head * 32;
          const float* bq = bias_s + part * 96;
#pragma unroll
          for (int d = 0; d < 32; ++d) {
            dq[d] = __float2bfloat16(__uint_as_float(vq[d]) + bq[d]);
            dq[C + d] = __float2bfloat16(__uint_as_float(vk[d]) + bq[32 + d]);
          }
        }
      }
      {
        uint32_t vv[32];
        tmem_ld32(acc_addr + 64, vv);
        tmem_wait_ld();
        tc_fence_before();  // this warp is done with the QKV accumulator
        const float4* b4 = reinterpret_cast<const float4*>(bias_s + part * 96 + 64);
        const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const float4 v0 = b4[2 * qd], v1 = b4[2 * qd + 1];
          const int d = qd * 8;
          af_sts128(v_base + xrow + (((uint32_t)(part * 4 + qd) ^ sw) << 4),
                    af_pack(__uint_as_float(vv[d]) + v0.x, __uint_as_float(vv[d + 1]) + v0.y),
                    af_pack(__uint_as_float(vv[d + 2]) + v0.z, __uint_as_float(vv[d + 3]) + v0.w),
                    af_pack(__uint_as_float(vv[d + 4]) + v1.x, __uint_as_float(vv[d + 5]) + v1.y),
                    af_pack(__uint_as_float(vv[d + 6]) + v1.z, __uint_as_float(vv[d + 7]) + v1.w));
        }
        if (p.dbg_qkv && src >= 0) {
          bf16* dv = p.dbg_qkv + src * (3 * C) + 2 * C + head * 32;
          const float* bv = bias_s + part * 96 + 64;
#pragma unroll
          for (int d = 0; d < 32; ++d) dv[d] = __float2bfloat16(__uint_as_float(vv[d]) + bv[d]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&qkv_ready[part]));  // (tells the projection issuer that this head's accumulator is free)
      // the head's four warps meet; its leader issues S_p = Q_p K_p^T (two windows stacked) itself
      asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
      if (quad == 0) {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16_pred(s_addr_mma, umma_desc_sw128(qk_tile + k * 32), umma_desc_sw128(qk_tile + 64 + k * 32), idesc_s, k != 0);
        umma_commit_pred(smem_u32(&s_full[part]));
      }
      AF_MARK(tD)

      // ---- (2) softmax of this thread's query row ----
      if (lane == 0) mbar_wait(smem_u32(&s_full[part]), lt & 1);
      __syncwarp();
      AF_MARK(tSw)
      tc_fence_after();
      float inv;
      {
        uint32_t sa[32], sb[32];
        tmem_ld32(s_addr + w * 64, sa);
        tmem_ld32(s_addr + w * 64 + 32, sb);
        tmem_wait_ld();
        float s[64];
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains (two warps per scheduler: ILP hides latency)
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
        // P goes back into TENSOR MEMORY, bf16 pairs in columns 0..31 of the head's S region, and is the TMEM A operand of PV
        // (an N = 32 MMA takes 16 clk with A in TMEM against 40 from shared memory, tools/micro/umma_rate.cu; no shared-memory
        // stores, no proxy fence).  Columns 0..31 are free in every lane: rows of window 0 have just read them (their own S),
        // rows of window 1 hold window-1-query x window-0-key products there that nobody reads.
#pragma unroll
        for (int h16 = 0; h16 < 2; ++h16) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j0 = h16 * 32 + e * 2;
            const float p0 = j0 < N ? af_ex2(s[j0 < N ? j0 : 0] - mx) : 0.f;          // keys past a 7x7 window: probability 0
            const float p1 = j0 + 1 < N ? af_ex2(s[j0 + 1 < N ? j0 + 1 : 0] - mx) : 0.f;
            suma[e & 3] += p0 + p1;
            pk[e] = af_pack(p0, p1);
          }
          tmem_st16(s_addr + h16 * 16, pk);
        }
        const float sum = (suma[0] + suma[1]) + (suma[2] + suma[3]);
        inv = 1.0f / sum;
      }
      tmem_wait_st();
      tc_fence_before();  // S_p has been read and P_p written: PV_p may run (it writes O into columns 64..127 of the region)
      asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
      if (quad == 0) {  // O_p[:, 32w..] = P_p . V_{p,w}: 16 keys per step = 32 B along K in the P tile, two 8-key groups (2 KB) in the V tile
        tc_fence_after();
#pragma unroll
        for (int ww = 0; ww < 2; ++ww)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ts_pred(s_addr_mma + 64 + ww * 32, s_addr_mma + k * 8, af_desc_mn_sw128(v_base + ww * 8192 + part * 64 + k * 2048), idesc_pv, k != 0);
        umma_commit_pred(smem_u32(&o_full[part]));
      }
      AF_MARK(tS)

      // ---- (3) O = P V -> * 1/rowsum -> bf16 -> global (window reverse + roll back = the row's source token) ----
      if (lane == 0) mbar_wait(smem_u32(&o_full[part]), lt & 1);
      __syncwarp();
      AF_MARK(tOw)
      tc_fence_after();
      {
        uint32_t ov[32];
        tmem_ld32(s_addr + 64 + w * 32, ov);
        tmem_wait_ld();
        tc_fence_before();
        if (src >= 0) {
          bf16* op = p.out + src * p.ldo + head * 32;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              pk[e] = af_pack(__uint_as_float(ov[h2 * 16 + 2 * e]) * inv, __uint_as_float(ov[h2 * 16 + 2 * e + 1]) * inv);
            st_global_256(op + h2 * 16, pk);
          }
        }
      }
      AF_MARK(tO)
    }
#ifdef MST_AF_PROF
    if (blockIdx.x == 2 && lane == 0 && (warp == 0 || warp == 5))
      printf("af prof C=%d WS=%d warp %d tiles %d total %lld | book %lld | acc wait %lld drain %lld | s wait %lld softmax %lld | o wait %lld out %lld\n", C, WS,
             warp, n_my, clock64() - t_all, tB, tAw, tD, tSw, tS, tOw, tO);
#endif
    tc_fence_before();
  }
  __syncthreads();
  if (warp == AF_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- weight packing
// dst holds, per head pair hp, the exact shared-memory image of its [192 x C] weight slice: KB k-blocks of [192 rows x 64]
// bf16, 128B-swizzled K-major.  Row n of the slice = head 2hp + n/96, matrix (n%96)/32 (q, k, v), dim n%32.
__global__ void pack_attn_qkv_kernel(const float* __restrict__ wq, const float* __restrict__ wk, const float* __restrict__ wv,
                                     const float* __restrict__ bq, const float* __restrict__ bk, const float* __restrict__ bv,
                                     bf16* __restrict__ dst_w, float* __restrict__ dst_b, int C, int heads) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nhp = heads / 2;
  const long long total = (long long)nhp * AF_WROWS * C;
  if (i < total) {
    const int kk = (int)(i % C);
    const int n = (int)((i / C) % AF_WROWS);
    const int hp = (int)(i / ((long long)C * AF_WROWS));
    const int part = n / 96, rem = n - part * 96, which = rem / 32, d = rem - which * 32;
    const int row = (hp * 2 + part) * 32 + d;
    const float* src = which == 0 ? wq : (which == 1 ? wk : wv);
    const int kb = kk >> 6, k64 = kk & 63, c = k64 >> 3, e = k64 & 7;
    const long long off = ((long long)hp * (C / 64) + kb) * (AF_WKB_BYTES / 2) + ((n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4)) / 2 + e;
    dst_w[off] = __float2bfloat16(src[(long long)row * C + kk]);
  }
  if (i < (long long)nhp * AF_WROWS) {
    const int n = (int)(i % AF_WROWS), hp = (int)(i / AF_WROWS);
    const int part = n / 96, rem = n - part * 96, which = rem / 32, d = rem - which * 32;
    const int row = (hp * 2 + part) * 32 + d;
    const float* src = which == 0 ? bq : (which == 1 ? bk : bv);
    dst_b[i] = src ? src[row] : 0.f;
  }
}

typedef CUresult (*AfEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static AfEncodeTiledFn af_tma_encoder() {
  static AfEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* e = getenv("MST_ATTN_TMA");
    if (e && e[0] == '0') return nullptr;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<AfEncodeTiledFn>(q);
  }
  return fn;
}

static int af_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

template <int C, int WS>
static int launch_attn_fused(const MstAttnBlock& a, cudaStream_t st) {
  using Cfg = AttnCfg<C>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fused_kernel<C, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const WinGeom g = make_geom(a.H, a.W, a.ws, a.shift);
  const long long total = (long long)a.B * g.nW;
  if (total <= 0 || total > 0x3fffffffLL) return MST_ERR_BAD_ARG;
  const int n_tiles = (int)((total + 1) / 2);
  const int nhp = a.heads / 2;
  long long grid = (long long)n_tiles * nhp;
  const long long cap = (long long)(af_num_sms() / nhp) * nhp;
  if (grid > cap) grid = cap;
  // x as a [C, W, H, B] tensor: a window that does not wrap is one [64 x ws x ws] box per k-block
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  int use_tma = 0;
  if (AfEncodeTiledFn enc = af_tma_encoder()) {
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    const cuuint64_t gstride[3] = {(cuuint64_t)a.ldx * 2, (cuuint64_t)a.W * a.ldx * 2, (cuuint64_t)a.H * a.W * a.ldx * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)WS, (cuuint32_t)WS, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    use_tma = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(reinterpret_cast<const void*>(a.x)), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  AttnCore core;
  core.x = reinterpret_cast<const bf16*>(a.x);
  core.wqkv = reinterpret_cast<const uint8_t*>(a.wqkv);
  core.bqkv = a.bqkv;
  core.bias_table = a.bias_table;
  core.out = reinterpret_cast<bf16*>(a.out);
  core.dbg_qkv = reinterpret_cast<bf16*>(a.dbg_qkv);
  core.B = a.B; core.H = a.H; core.W = a.W; core.heads = a.heads; core.ldx = a.ldx; core.ldo = a.ldo;
  static int stagger = -1;
  if (stagger < 0) {
    const char* e = getenv("MST_ATTN_STAGGER");  // 0: heads start together; 1 (default): head 1 starts when head 0's S is issued; 2: when its PV is
    stagger = e ? atoi(e) : 1;
  }
  attn_fused_kernel<C, WS><<<(unsigned)grid, AF_THREADS, Cfg::SMEM_BYTES, st>>>(core, g, tmap, use_tma, n_tiles, (int)total, stagger);
  return (int)cudaGetLastError();
}

}  // namespace mst

using namespace mst;

extern "C" size_t mst_attn_qkv_packed_bytes(int C, int heads) {
  return ((C == 128 || C == 256) && heads * 32 == C) ? (size_t)3 * C * C * 2 : 0;
}

extern "C" int mst_pack_attn_qkv(const float* wq, const float* wk, const float* wv, const float* bq, const float* bk, const float* bv,
                                 mst_bf16* dst_w, float* dst_b, int C, int heads, void* stream) {
  if (!wq || !wk || !wv || !dst_w || !dst_b || (C != 128 && C != 256) || heads * 32 != C) return MST_ERR_BAD_ARG;
  const long long n = 3LL * C * C;
  pack_attn_qkv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wq, wk, wv, bq, bk, bv, reinterpret_cast<bf16*>(dst_w), dst_b, C,
                                                                                     heads);
  return (int)cudaGetLastError();
}

extern "C" int mst_attn_block(const MstAttnBlock* a, void* stream) {
  if (!a || !a->x || !a->wqkv || !a->bqkv || !a->bias_table || !a->out) return MST_ERR_BAD_ARG;
  if (a->B <= 0 || a->H <= 0 || a->W <= 0 || a->shift < 0 || a->shift >= a->ws) return MST_ERR_BAD_ARG;
  if (a->heads * 32 != a->C || (a->heads & 1)) return MST_ERR_UNSUPPORTED;
  if (a->ldx % 8 != 0 || a->ldx < a->C || a->ldo % 16 != 0 || a->ldo < a->C) return MST_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(a->x) & 15) || (reinterpret_cast<uintptr_t>(a->out) & 31) || (reinterpret_cast<uintptr_t>(a->wqkv) & 15))
    return MST_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->C == 128 && a->ws == 7) return launch_attn_fused<128, 7>(*a, st);
  if (a->C == 128 && a->ws == 8) return launch_attn_fused<128, 8>(*a, st);
  if (a->C == 256 && a->ws == 7) return launch_attn_fused<256, 7>(*a, st);
  if (a->C == 256 && a->ws == 8) return launch_attn_fused<256, 8>(*a, st);
  return MST_ERR_UNSUPPORTED;
}
