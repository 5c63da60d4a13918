// Tensor-core GEMM for sm_100a: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), persistent and warp-specialised.
//
//   warps 0-7  : epilogue.  warp w reads TMEM lane quadrant w%4 (rows) and column half w/4 of the finished
//                accumulator, applies bias / activation / residual / blend and stores fp32 and/or bf16.
//   warps 8-11 : A-operand producers.  128 threads gather 16-byte chunks with cp.async into 128B-swizzled
//                K-major stages (F.linear rows, or 3x3-conv taps with zero / reflect padding and an optional
//                nearest x2 upsample folded into the address).  Thread 0 of warp 8 also issues ONE bulk copy
//                (cp.async.bulk, UBLKCP) per stage for the weight tile, which is stored pre-swizzled in HBM.
//   warp 12    : TMEM allocation; one elected lane issues tcgen05.mma and commits to mbarriers.
//
// Each CTA loops over output tiles (128 x BN) with a static stride schedule; the accumulator is double
// buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the loads and MMAs of tile i+1.
#include "../../include/mst_b200.h"
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace mst {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_PROD_WARPS = 7;  // 8 + 7 + 1 = 16 warps: register files are allotted per 4 warps, a 17th warp costs 32 regs/thread
constexpr int GEMM_THREADS = (NUM_EPI_WARPS + NUM_PROD_WARPS + 1) * 32;
constexpr int MAX_BIAS = 1024;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 5 : 6);
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + slack for manual 1024 B alignment
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;     // two accumulators
  static constexpr int EPI_SPLIT = BN >= 64 ? 2 : 1;               // column halves handled by warps 0-3 / 4-7
  static constexpr int COLS_PER_WARP = BN / EPI_SPLIT;
};

MST_DEVINL void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
MST_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// res_stages == 0: streaming mode -- every stage carries an A k-block and the matching weight k-block.
// res_stages  > 0: weight-resident mode (the CTA's [BN x k_pad] weight slab fits in shared memory): the slab is
//                  fetched once, the CTA keeps one n-tile and walks m-tiles, and the ring (res_stages deep) carries only
//                  A k-blocks.  Cuts the L2 traffic of the K<=256 projections by the weight re-reads, which
//                  otherwise exceed the activation traffic (ncu: lts__t_bytes ~3.6x the algorithmic bytes).
constexpr int MAX_STAGES = 8;
// EXT = false: the inference epilogue (bias / activation / residual / blend).  EXT = true additionally compiles the
// training-step epilogue (pre-activation copy, ReLU / GELU' gates, bf16 addend, per-sample row scale) and the
// conv_full gather; kept out of the inference instantiation so its register allocation and code size are untouched.
template <bool EXT> struct ExtSel { typedef GemmNoExt type; };
template <> struct ExtSel<true> { typedef GemmExt type; };
// TMA = true: the A operand of a plain (F.linear) GEMM is fetched with ONE cp.async.bulk.tensor per stage (2-D tensor map over
// A[M, lda], box 64 x 128, SWIZZLE_128B -- the hardware writes exactly the swizzled K-major UMMA image, zero-fills the M / K
// tails) instead of 1024 16-byte cp.async: ncu showed the LSU data pipe as the busiest unit of the K <= 256 projections
// (l1tex__data_pipe_lsu_wavefronts 64 %), and the 16-byte LDGSTS path tops out near 11 B/clk/SM against L2.
MST_DEVINL void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
               : "memory");
}
struct TmaNone {};
template <bool TMA> struct TmaSel { typedef TmaNone type; };
template <> struct TmaSel<true> { typedef CUtensorMap type; };
MST_DEVINL void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

MST_DEVINL void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
constexpr int TMO_SLAB_BYTES = 4096;  // one output staging slab per epilogue warp

// TMO = true (its own instantiation: plain epilogues with exactly one output, no residual): the output leaves through TMA tensor
// stores.  A lane owns a row, so direct stores touch 32 different rows per instruction (the pattern that bounded the patch embedding
// and the fused MLP before their stores went through TMA); here each epilogue warp stages its [32 rows x 32 columns] block in a 4 KB
// slab at slab_off (fp32: 128-byte rows, 128B swizzle; bf16: 64-byte rows, 64B swizzle) and one lane issues the store.  Rows past M
// are clipped by the tensor map tmo_ ([N columns, M rows] of the output, row stride ld_out).
template <int BN, bool EXT, bool TMA, bool TMO = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const GemmCore p, const typename ExtSel<EXT>::type x_,
                                                                  const __grid_constant__ typename TmaSel<TMA>::type tm_, const int num_tiles,
                                                                  const int res_stages, const int wsplit,
                                                                  const __grid_constant__ typename TmaSel<TMO>::type tmo_, const int slab_off) {
  using Cfg = GemmCfg<BN>;
  static_assert(!TMO || (!EXT && Cfg::COLS_PER_WARP >= 32), "TMA output stores: inference epilogue, 32-column chunks");
  const bool resident = res_stages > 0;
  const int STAGES = resident ? res_stages : Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t b_full_bar;
  __shared__ uint64_t tma_bar[MAX_STAGES];  // conv TMA mode: tensor copies land here, the producer warp patches reflect borders, then arrives on full_bar
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float bias_s[MAX_BIAS];

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for ptxas (uniform-register MMA issue loop)
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  const int n_tiles = p.N / BN;
  const int nkb = p.k_pad / BK;
  // tile schedule (identical in every role), advanced without divisions inside the loops
  const int m_tiles = num_tiles / n_tiles;
  const int g_q = gridDim.x / n_tiles, g_r = gridDim.x % n_tiles;
  struct TileIter {
    int n_tile, m_tile;
  };
  auto tile_begin = [&]() { return TileIter{(int)(blockIdx.x % n_tiles), (int)(blockIdx.x / n_tiles)}; };
  auto tile_next = [&](TileIter& t) {  // streaming: tile index += gridDim.x ; resident: same n-tile, m += gridDim.x / n_tiles
    t.m_tile += g_q;
    if (!resident) {
      t.n_tile += g_r;
      if (t.n_tile >= n_tiles) { t.n_tile -= n_tiles; ++t.m_tile; }
    }
  };
  const int res_nt = blockIdx.x % n_tiles;
  // shared-memory carve-up: [resident weight slab (nkb k-blocks)] [ring of stages]
  const uint32_t ring_base = smem_base + (resident ? nkb * Cfg::B_STAGE_BYTES : 0);
  const uint32_t stage_bytes = resident ? A_STAGE_BYTES : Cfg::STAGE_BYTES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), TMA ? 1 : NUM_PROD_WARPS * 32 + (resident ? 0 : 1));
      mbar_init(smem_u32(&empty_bar[s]), 1);
      mbar_init(smem_u32(&tma_bar[s]), 1);
    }
    mbar_init(smem_u32(&b_full_bar), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full_bar[b]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[b]), NUM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < p.N; i += GEMM_THREADS) bias_s[i] = p.bias ? p.bias[i] : 0.f;
  if (warp == NUM_EPI_WARPS + NUM_PROD_WARPS) {
    tmem_alloc(smem_u32(&tmem_base_slot), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= NUM_EPI_WARPS && warp < NUM_EPI_WARPS + NUM_PROD_WARPS) {
    // =========================== producers ===========================
    constexpr int RSTEP = NUM_PROD_WARPS * 4;            // row distance between a thread's rows (28)
    constexpr int RPT = (BM + RSTEP - 1) / RSTEP;        // rows per thread (5; the fifth only for r0 < BM - 4*RSTEP)
    const int t = threadIdx.x - NUM_EPI_WARPS * 32;
    const int c = t & 7;    // 16-byte chunk column inside the 128-byte k-slab
    const int r0 = t >> 3;  // first of this thread's rows; rows r0 + RSTEP*i
    bool full = false;
    if constexpr (EXT) full = x_.conv_full != 0;
    const int Hs = full ? p.H - 2 : (p.upsample ? (p.H >> 1) : p.H);  // stored input grid
    const int Ws = full ? p.W - 2 : (p.upsample ? (p.W >> 1) : p.W);
    const int tap_off = full ? 2 : 1;
    uint32_t a_dst[RPT];  // swizzled byte offset of (row r0 + RSTEP*i, chunk c) inside a stage
    bool row_in[RPT];     // the row exists in the 128-row tile (false only for the last i of the higher producer warps)
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      row_in[i] = r0 + RSTEP * i < BM;
      a_dst[i] = sw128_offset(r0 + RSTEP * i, c);
    }
    const bf16* Abase = reinterpret_cast<const bf16*>(p.A);
    const uint8_t* Wbase = reinterpret_cast<const uint8_t*>(p.Wt);
    // wsplit = 1: the weights were packed in 2 BN-wide tiles (mst_gemm_tile_n) and this launch runs them as two BN-wide halves
    // (small M: twice the CTAs): half h of packed block (n_tile >> 1, kb) starts B_STAGE_BYTES * h into the block
    const size_t w_kb_stride = (size_t)Cfg::B_STAGE_BYTES << wsplit;
    auto wtile_of = [&](int nt) { return Wbase + (size_t)(nt >> wsplit) * nkb * w_kb_stride + (size_t)(nt & wsplit) * Cfg::B_STAGE_BYTES; };
    const bool conv = p.a_mode != MST_A_PLAIN;
    if (resident && t == 0) {  // the CTA's whole weight slab, once
      mbar_arrive_expect_tx(smem_u32(&b_full_bar), nkb * Cfg::B_STAGE_BYTES);
      const uint8_t* wslab = wtile_of(res_nt);
      for (int kb = 0; kb < nkb; ++kb)
        bulk_g2s(smem_base + kb * Cfg::B_STAGE_BYTES, wslab + (size_t)kb * w_kb_stride, Cfg::B_STAGE_BYTES, smem_u32(&b_full_bar));
    }
    int stage = 0;
    uint32_t pphase = 1;  // producer's view of empty_bar: a fresh barrier passes a wait on parity 1
    if constexpr (TMA) {
      if (conv) {
        // 3x3 conv as implicit GEMM fed by TMA: k-block kb = 64 channels (c0 ..) of tap (ky, kx); the 128 tile rows are
        // 128 / W whole image rows (or a 128-pixel row segment), each fetched with one 4-D tensor copy at pixel offset
        // (kx - 1, ky - 1): out-of-image coordinates are zero-filled by the hardware (= zero padding).  Reflect padding:
        // the row coordinate is reflected before the copy, and the one border pixel per row a kx != 1 tap reads out of the
        // image is patched (its 128 bytes are requested before the copy lands, written after) by the issuing warp.
        // Each ring stage is owned by one producer warp, so several stages are being issued / patched at any time.
        const int pw = warp - NUM_EPI_WARPS;
        const int Wb = p.W < BM ? p.W : BM;         // pixels per tensor copy
        const int nrow = BM / Wb;                    // image rows (or 1 segment) per tile
        const int hw = p.H * p.W;
        const int kb_per_tap = p.Cin >> 6;
        const uint32_t row_bytes = (uint32_t)Wb * 128u;
        const uint32_t stage_tx = A_STAGE_BYTES + (resident ? 0 : Cfg::B_STAGE_BYTES);
        long long it = 0;  // running (tile, k-block) count: ring position and the warp that owns it
        for (TileIter tt = tile_begin(); tt.m_tile < m_tiles; tile_next(tt)) {
          const int m0 = tt.m_tile * BM;
          const int b = m0 / hw;
          const int rem = m0 - b * hw;
          const int y0 = rem / p.W, x0 = rem - y0 * p.W;
          const uint8_t* wtile = wtile_of(tt.n_tile);
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = (int)(it % STAGES);
            // a ring stage is owned by ONE warp, which therefore sees its uses in order: a parity wait cannot tell "phase
            // u-1 complete" from "phase u-3 complete", so the waiter must itself have produced use u-1
            if (s % NUM_PROD_WARPS != pw) continue;
            const uint32_t ph = (uint32_t)((it / STAGES) & 1);
            const int tap = kb / kb_per_tap, c0c = (kb - tap * kb_per_tap) << 6;
            const int ky = tap / 3, kx = tap - ky * 3;
            // reflect patch: lane = (image row j of the tile, 16-byte chunk); source pixel x = 1 (kx = 0) or W - 2 (kx = 2)
            const int j = lane >> 3, chunk = lane & 7;
            const bool left = kx == 0 && x0 == 0, right = kx == 2 && x0 + Wb == p.W;
            const bool patch = p.pad_mode == 1 && (left || right) && j < nrow;
            uint4 pv = make_uint4(0, 0, 0, 0);
            if (patch) {
              int yy = y0 + j + ky - 1;
              yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
              const int xs = left ? 1 : p.W - 2;
              pv = *reinterpret_cast<const uint4*>(Abase + ((long long)(b * p.H + yy) * p.W + xs) * p.Cin + c0c + chunk * 8);
            }
            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
            const uint32_t a_stage = ring_base + s * stage_bytes;
            if (lane == 0) {
              mbar_arrive_expect_tx(smem_u32(&tma_bar[s]), stage_tx);
              for (int r = 0; r < nrow; ++r) {
                int yy = y0 + r + ky - 1;
                if (p.pad_mode == 1) yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
                tma_load_4d(a_stage + r * row_bytes, &tm_, c0c, x0 + kx - 1, yy, b, smem_u32(&tma_bar[s]));
              }
              if (!resident) bulk_g2s(a_stage + A_STAGE_BYTES, wtile + (size_t)kb * w_kb_stride, Cfg::B_STAGE_BYTES, smem_u32(&tma_bar[s]));
            }
            mbar_wait(smem_u32(&tma_bar[s]), ph);
            if (patch) {
              const int row = j * Wb + (left ? 0 : Wb - 1);
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_stage + sw128_offset(row, chunk)), "r"(pv.x), "r"(pv.y), "r"(pv.z), "r"(pv.w) : "memory");
              fence_proxy_async_smem();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
          }
        }
      } else if (t == 0) {  // one thread feeds the whole ring: a tensor copy for A (+ a bulk copy for the weight tile) per stage
        for (TileIter tt = tile_begin(); tt.m_tile < m_tiles; tile_next(tt)) {
          const int m0 = tt.m_tile * BM;
          const uint8_t* wtile = wtile_of(tt.n_tile);
          for (int kb = 0; kb < nkb; ++kb) {
            const int s = stage;
            mbar_wait(smem_u32(&empty_bar[s]), pphase);
            if (++stage == STAGES) { stage = 0; pphase ^= 1; }
            const uint32_t a_stage = ring_base + s * stage_bytes;
            mbar_arrive_expect_tx(smem_u32(&full_bar[s]), A_STAGE_BYTES + (resident ? 0 : Cfg::B_STAGE_BYTES));
            tma_load_2d(a_stage, &tm_, kb * BK, m0, smem_u32(&full_bar[s]));
            if (!resident) bulk_g2s(a_stage + A_STAGE_BYTES, wtile + (size_t)kb * w_kb_stride, Cfg::B_STAGE_BYTES, smem_u32(&full_bar[s]));
          }
        }
      }
    } else
    for (TileIter tt = tile_begin(); tt.m_tile < m_tiles; tile_next(tt)) {
      const int n_tile = tt.n_tile, m0 = tt.m_tile * BM;
      // Per-row source bookkeeping, computed once per tile.
      //   PLAIN: rowp = &A[m*lda].   CONV: rowp = image base; yo/xo = element offsets of the three source rows /
      //   columns a 3x3 tap can touch (reflect or zero padding and the nearest-x2 upsample already applied),
      //   vmask bit ky / bit 3+kx = that source row / column exists.
      const bf16* rowp[RPT];
      int yo[RPT][3], xo[RPT][3];
      uint32_t vmask[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int m = m0 + r0 + RSTEP * i;
        vmask[i] = 0;
        rowp[i] = Abase;
#pragma unroll
        for (int j = 0; j < 3; ++j) { yo[i][j] = 0; xo[i][j] = 0; }
        if (row_in[i] && m < p.M) {
          if (!conv) {
            rowp[i] = Abase + (long long)m * p.lda;
            vmask[i] = 0x3F;
          } else {
            const int hw = p.H * p.W;
            const int b = m / hw;
            const int rem = m - b * hw;
            const int y = rem / p.W;
            const int x = rem - y * p.W;
            rowp[i] = Abase + (long long)b * Hs * Ws * p.Cin;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              int yy = y + j - tap_off, xx = x + j - tap_off;
              bool vy = true, vx = true;
              if (p.pad_mode == 1) {  // reflect (no edge repeat): -1 -> 1, H -> H-2
                yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
                xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
              } else {  // zeros; conv_full: the input grid is (H-2) x (W-2)
                vy = (unsigned)yy < (unsigned)(full ? Hs : p.H);
                vx = (unsigned)xx < (unsigned)(full ? Ws : p.W);
              }
              if (p.upsample) { yy >>= 1; xx >>= 1; }
              yo[i][j] = vy ? yy * Ws * p.Cin : 0;
              xo[i][j] = vx ? xx * p.Cin : 0;
              vmask[i] |= (vy ? 1u : 0u) << j | (vx ? 1u : 0u) << (3 + j);
            }
          }
        }
      }
      const uint8_t* wtile = wtile_of(n_tile);
      int tap = 0, ch = c * 8;  // conv: this thread's (tap, channel) for k0 = kb*64 + c*8, advanced without divisions
      while (conv && ch >= p.Cin) { ch -= p.Cin; ++tap; }
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = stage;
        mbar_wait(smem_u32(&empty_bar[s]), pphase);
        if (++stage == STAGES) { stage = 0; pphase ^= 1; }
        const uint32_t a_stage = ring_base + s * stage_bytes;
        if (!resident && t == 0) {  // weight tile: one bulk copy of the pre-swizzled [BN x 64] block
          mbar_arrive_expect_tx(smem_u32(&full_bar[s]), Cfg::B_STAGE_BYTES);
          bulk_g2s(a_stage + A_STAGE_BYTES, wtile + (size_t)kb * w_kb_stride, Cfg::B_STAGE_BYTES, smem_u32(&full_bar[s]));
        }
        if (!conv) {
          const int k0 = kb * BK + c * 8;
          const bool kvalid = k0 < p.K;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const bool valid = kvalid && vmask[i] != 0;
            if (row_in[i]) cp_async16(a_stage + a_dst[i], valid ? rowp[i] + k0 : Abase, valid);
          }
        } else {
          const int ky = tap / 3, kx = tap - ky * 3;  // tap < 16: cheap constant division
          const bool kvalid = tap < 9;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int yoff = ky == 0 ? yo[i][0] : (ky == 1 ? yo[i][1] : yo[i][2]);
            const int xoff = kx == 0 ? xo[i][0] : (kx == 1 ? xo[i][1] : xo[i][2]);
            const bool valid = kvalid && ((vmask[i] >> ky) & (vmask[i] >> (3 + kx)) & 1u);
            if (row_in[i]) cp_async16(a_stage + a_dst[i], valid ? rowp[i] + (yoff + xoff + ch) : Abase, valid);
          }
          ch += BK;
          while (ch >= p.Cin) { ch -= p.Cin; ++tap; }
        }
        // asynchronous arrival: counts as this thread's arrival once all of its cp.async above have landed,
        // so the thread never blocks and every stage of the ring can be in flight
        cp_async_mbar_arrive_noinc(smem_u32(&full_bar[s]));
      }
    }
    cp_async_wait_all();  // nothing may still be writing shared memory when the CTA exits
  } else if (warp == NUM_EPI_WARPS + NUM_PROD_WARPS) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    int tcount = 0, stage = 0;
    uint32_t cphase = 0;
    if (resident) mbar_wait(smem_u32(&b_full_bar), 0);
    for (TileIter tt = tile_begin(); tt.m_tile < m_tiles; tile_next(tt), ++tcount) {
      const int buf = tcount & 1;
      mbar_wait(smem_u32(&tmem_empty_bar[buf]), ((tcount >> 1) & 1) ^ 1);  // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = stage;
        mbar_wait(smem_u32(&full_bar[s]), cphase);
        if (++stage == STAGES) { stage = 0; cphase ^= 1; }
        tc_fence_after();
        {
          // The whole warp runs this convergently on warp-uniform values; one lane is elected inside the asm.  ptxas
          // then forms the descriptors in uniform registers and emits back-to-back UTCHMMA (issuing from an
          // `if (lane == 0)` branch costs ~15 instructions of register -> uniform-register broadcasts per MMA).
          const uint32_t a_stage = ring_base + s * stage_bytes;
          const uint32_t b_stage = resident ? smem_base + kb * Cfg::B_STAGE_BYTES : a_stage + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr>>4) field
            umma_bf16_pred(d_tmem, umma_desc_sw128(a_stage + k * 32), umma_desc_sw128(b_stage + k * 32), idesc, (kb | k) != 0);
          }
          umma_commit_pred(smem_u32(&empty_bar[s]));                          // stage reusable once these MMAs retire
          if (kb == nkb - 1) umma_commit_pred(smem_u32(&tmem_full_bar[buf]));  // accumulator complete
        }
      }
    }
    tc_fence_before();
  } else {
    // =========================== epilogue (warps 0-7) ===========================
    constexpr int CPW = Cfg::COLS_PER_WARP;
    constexpr int CH = CPW >= 32 ? 32 : 16;  // columns per tcgen05.ld
    const int quad = warp & 3;
    const int half = warp >> 2;
    const bool active = half < Cfg::EPI_SPLIT;
    // 256-bit accesses when every row segment starts on a 32-byte boundary (CH >= 16 columns per chunk)
    const bool wide16 = CH >= 16 && p.out_bf16 && ((reinterpret_cast<uintptr_t>(p.out_bf16) | (uintptr_t)(p.ld_out16 * 2)) & 31) == 0;
    const bool wide32 = p.out_f32 && !p.out_nchw && ((reinterpret_cast<uintptr_t>(p.out_f32) | (uintptr_t)(p.ld_out32 * 4)) & 31) == 0;
    const bool wideres = p.res && ((reinterpret_cast<uintptr_t>(p.res) | (uintptr_t)(p.ld_res * 4) |
                                    (p.mul ? reinterpret_cast<uintptr_t>(p.mul) : 0)) & 31) == 0;
    int tcount = 0;
    for (TileIter tt = tile_begin(); tt.m_tile < m_tiles; tile_next(tt), ++tcount) {
      const int n_tile = tt.n_tile, m0 = tt.m_tile * BM;
      const int buf = tcount & 1;
      if (lane == 0) mbar_wait(smem_u32(&tmem_full_bar[buf]), (tcount >> 1) & 1);  // one poller per warp
      __syncwarp();
      tc_fence_after();
      if (active) {
        const int row = m0 + quad * 32 + lane;
        const bool row_ok = row < p.M;
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN + half * CPW;
        const int nbase = n_tile * BN + half * CPW;
        long long nchw_base = 0;
        int hw = 0;
        if (p.out_nchw && row_ok) {
          hw = p.H * p.W;
          const int b = row / hw;
          nchw_base = (long long)b * p.n_real * hw + (row - b * hw);
        }
#pragma unroll 1
        for (int col0 = 0; col0 < CPW; col0 += CH) {
          uint32_t v[CH];
          if constexpr (CH == 32) tmem_ld32(t_row + col0, v);
          else tmem_ld16(t_row + col0, v);
          tmem_wait_ld();
          if (col0 + CH >= CPW) {  // last read of this accumulator: hand it back to the MMA warp early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
          }
          if (!TMO && !row_ok) continue;  // (TMO: the whole warp takes part in the tensor store; rows past M are clipped by the map)
          const int n = nbase + col0;
          float x[CH];
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + n);  // n % 16 == 0: LDS.128 broadcasts (4x fewer LSU wavefronts)
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) {
            const float4 bb = b4[j];
            x[4 * j] = __uint_as_float(v[4 * j]) + bb.x; x[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
            x[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z; x[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
          }
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            if constexpr (!EXT) {
              if (p.act == MST_ACT_RELU) x[j] = fmaxf(x[j], 0.0f);
              else if (p.act == MST_ACT_GELU) x[j] = gelu_erf(x[j]);
            }
          }
          if constexpr (EXT) if (x_.out_pre16) {  // pre-activation copy for the backward pass
            bf16* op = reinterpret_cast<bf16*>(x_.out_pre16) + (long long)row * p.ld_out16 + n;
            if (CH >= 16 && ((reinterpret_cast<uintptr_t>(x_.out_pre16) | (uintptr_t)(p.ld_out16 * 2)) & 31) == 0) {
#pragma unroll
              for (int j = 0; j < CH / 16; ++j) {
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(x[16 * j + 2 * e], x[16 * j + 2 * e + 1]);
                  pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
                st_global_256(op + 16 * j, pk);
              }
            } else {
              uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
              for (int j = 0; j < CH / 8; ++j) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(x[8 * j + 2 * e], x[8 * j + 2 * e + 1]);
                  pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
                o4[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
          if constexpr (EXT) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              if (p.act == MST_ACT_RELU) x[j] = fmaxf(x[j], 0.0f);
              else if (p.act == MST_ACT_GELU) x[j] = gelu_erf(x[j]);
            }
          }
          if constexpr (EXT) {
            const bool wide_g = CH >= 16 && ((uintptr_t)(x_.ld_gate * 2) & 31) == 0;
            auto load_row = [&](const void* base, uint32_t (&w)[CH / 2]) {  // CH bf16 of this row, as packed pairs
              const bf16* gp = reinterpret_cast<const bf16*>(base) + (long long)row * x_.ld_gate + n;
              if (wide_g && (reinterpret_cast<uintptr_t>(base) & 31) == 0) {
#pragma unroll
                for (int j = 0; j < CH / 16; ++j) {
                  uint32_t t8[8];
                  ld_global_256(gp + 16 * j, t8);
#pragma unroll
                  for (int e = 0; e < 8; ++e) w[8 * j + e] = t8[e];
                }
              } else {
#pragma unroll
                for (int j = 0; j < CH / 8; ++j) {
                  const uint4 t4 = reinterpret_cast<const uint4*>(gp)[j];
                  w[4 * j] = t4.x; w[4 * j + 1] = t4.y; w[4 * j + 2] = t4.z; w[4 * j + 3] = t4.w;
                }
              }
            };
            if (x_.gate_mode != MST_GATE_NONE) {
              uint32_t gw[CH / 2];
              load_row(x_.gate, gw);
#pragma unroll
              for (int e = 0; e < CH / 2; ++e) {
                const float g0 = __uint_as_float(gw[e] << 16), g1 = __uint_as_float(gw[e] & 0xffff0000u);
                if (x_.gate_mode == MST_GATE_RELU) {
                  x[2 * e] = g0 > 0.f ? x[2 * e] : 0.f;
                  x[2 * e + 1] = g1 > 0.f ? x[2 * e + 1] : 0.f;
                } else {
                  x[2 * e] *= gelu_erf_grad(g0);
                  x[2 * e + 1] *= gelu_erf_grad(g1);
                }
              }
            }
            if (x_.add16) {
              uint32_t aw[CH / 2];
              load_row(x_.add16, aw);
#pragma unroll
              for (int e = 0; e < CH / 2; ++e) {
                x[2 * e] += __uint_as_float(aw[e] << 16);
                x[2 * e + 1] += __uint_as_float(aw[e] & 0xffff0000u);
              }
            }
          }
          if constexpr (EXT) if (x_.row_scale) {
            const float rs = x_.row_scale[row / x_.rows_per_scale];
#pragma unroll
            for (int j = 0; j < CH; ++j) x[j] *= rs;
          }
          if (p.res && wideres) {
            const float* rp = p.res + (long long)row * p.ld_res + n;
            const float* mp = p.mul ? p.mul + (long long)row * p.ld_res + n : nullptr;
#pragma unroll
            for (int j = 0; j < CH / 8; ++j) {
              float r[8];
              ld_global_256f(rp + 8 * j, r);
              if (mp) {
                float m[8];
                ld_global_256f(mp + 8 * j, m);
#pragma unroll
                for (int e = 0; e < 8; ++e) x[8 * j + e] += r[e] * m[e];
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[8 * j + e] += r[e];
              }
            }
          } else if (p.res) {
            const float4* r4 = reinterpret_cast<const float4*>(p.res + (long long)row * p.ld_res + n);
            if (p.mul) {
              const float4* m4 = reinterpret_cast<const float4*>(p.mul + (long long)row * p.ld_res + n);
#pragma unroll
              for (int j = 0; j < CH / 4; ++j) {
                const float4 r = r4[j], m = m4[j];
                x[4 * j + 0] += r.x * m.x; x[4 * j + 1] += r.y * m.y; x[4 * j + 2] += r.z * m.z; x[4 * j + 3] += r.w * m.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < CH / 4; ++j) {
                const float4 r = r4[j];
                x[4 * j + 0] += r.x; x[4 * j + 1] += r.y; x[4 * j + 2] += r.z; x[4 * j + 3] += r.w;
              }
            }
          }
          if constexpr (TMO) {
            const uint32_t slab = smem_base + (uint32_t)slab_off + (uint32_t)warp * TMO_SLAB_BYTES;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the slab's previous store has been read
            __syncwarp();
            if (p.out_f32) {
              const uint32_t srow = slab + (uint32_t)lane * 128u;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(srow + ((uint32_t)(q ^ (lane & 7)) << 4)), "f"(x[4 * q]),
                             "f"(x[4 * q + 1]), "f"(x[4 * q + 2]), "f"(x[4 * q + 3]) : "memory");
            } else {
              const uint32_t srow = slab + (uint32_t)lane * 64u;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(x[8 * q + 2 * e], x[8 * q + 2 * e + 1]);
                  pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(srow + ((uint32_t)(q ^ ((lane >> 1) & 3)) << 4)), "r"(pk[0]),
                             "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmo_, slab, n, m0 + quad * 32);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else if (p.out_nchw) {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (n + j < p.n_real) p.out_f32[nchw_base + (long long)(n + j) * hw] = x[j];
          } else {
            if (p.out_f32 && wide32) {
              float* op = p.out_f32 + (long long)row * p.ld_out32 + n;
#pragma unroll
              for (int j = 0; j < CH / 8; ++j) st_global_256f(op + 8 * j, x + 8 * j);
            } else if (p.out_f32) {
              float4* o4 = reinterpret_cast<float4*>(p.out_f32 + (long long)row * p.ld_out32 + n);
#pragma unroll
              for (int j = 0; j < CH / 4; ++j) o4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            }
            if (p.out_bf16 && wide16) {
              bf16* op = reinterpret_cast<bf16*>(p.out_bf16) + (long long)row * p.ld_out16 + n;
#pragma unroll
              for (int j = 0; j < CH / 16; ++j) {
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(x[16 * j + 2 * e], x[16 * j + 2 * e + 1]);
                  pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
                st_global_256(op + 16 * j, pk);
              }
            } else if (p.out_bf16) {
              uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out_bf16) + (long long)row * p.ld_out16 + n);
#pragma unroll
              for (int j = 0; j < CH / 8; ++j) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(x[8 * j + 2 * e], x[8 * j + 2 * e + 1]);
                  pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
                o4[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
      }
    }
    if (TMO && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA exits
    tc_fence_before();
  }
  __syncthreads();
  if (warp == NUM_EPI_WARPS + NUM_PROD_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static int g_num_sms = 0;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda); NULL = unavailable
static EncodeTiledFn tma_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* e = getenv("MST_GEMM_TMA");
    if (e && e[0] == '0') return nullptr;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int BN, bool EXT, bool TMA>
static int launch_gemm_ext(const MstGemm& g, cudaStream_t st, const typename TmaSel<TMA>::type& tmap, int wsplit) {
  using Cfg = GemmCfg<BN>;
  constexpr int MAX_SMEM = 222 * 1024;  // + ~4.3 KB static (bias, barriers) <= 227 KB
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, EXT, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const int n_tiles = g.N / BN;
  const int m_tiles = (g.M + BM - 1) / BM;
  const long long tiles = (long long)m_tiles * n_tiles;
  if (tiles <= 0 || tiles > 0x7fffffffLL) return MST_ERR_BAD_ARG;
  // weight-resident mode when the [BN x k_pad] slab leaves room for >= 3 A stages and there are enough m-tiles to amortise it
  const long long slab = (long long)(g.k_pad / BK) * Cfg::B_STAGE_BYTES;
  int res_stages = 0;
  if (slab + 3 * A_STAGE_BYTES + 1024 <= MAX_SMEM && n_tiles <= g_num_sms && m_tiles >= 2 * (g_num_sms / n_tiles)) {
    res_stages = (int)((MAX_SMEM - 1024 - slab) / A_STAGE_BYTES);
    if (res_stages > MAX_STAGES) res_stages = MAX_STAGES;
  }
  unsigned grid;
  size_t smem;
  if (res_stages > 0) {
    long long gsz = (long long)(g_num_sms / n_tiles) * n_tiles;
    if (gsz > tiles) gsz = tiles;
    grid = (unsigned)gsz;
    smem = 1024 + (size_t)slab + (size_t)res_stages * A_STAGE_BYTES;
  } else {
    grid = (unsigned)(tiles < g_num_sms ? tiles : g_num_sms);
    smem = Cfg::SMEM_BYTES;
  }
  GemmCore core;
  static_assert(sizeof(GemmCore) <= sizeof(MstGemm), "GemmCore must be a prefix of MstGemm");
  memcpy(&core, &g, sizeof(GemmCore));
  if constexpr (BN == 256 && !EXT && TMA) {
    // output through TMA tensor stores (see the kernel): plain linear layers with exactly one output and no residual
    static int tmo_allow = -1;
    if (tmo_allow < 0) { const char* e = getenv("MST_GEMM_TMA_OUT"); tmo_allow = e ? atoi(e) : 1; }
    const bool one_out = (g.out_f32 != nullptr) != (g.out_bf16 != nullptr);
    const int slabs = NUM_EPI_WARPS * TMO_SLAB_BYTES;
    int rs = res_stages;
    size_t smem_o = 0;
    if (res_stages > 0) {
      const int fit = (int)((MAX_SMEM - 1024 - slab - slabs) / A_STAGE_BYTES);
      if (rs > fit) rs = fit;
      smem_o = 1024 + (size_t)slab + (size_t)rs * A_STAGE_BYTES + slabs;
    } else {
      smem_o = (size_t)Cfg::SMEM_BYTES + slabs;
    }
    if (tmo_allow && wsplit == 0 && g.a_mode == MST_A_PLAIN && one_out && !g.out_nchw && !g.res && !g.mul && (res_stages == 0 || rs >= 3) &&
        smem_o <= (size_t)MAX_SMEM) {
      if (EncodeTiledFn enc = tma_encoder()) {
        alignas(64) CUtensorMap tmo;
        memset(&tmo, 0, sizeof(tmo));
        const bool f32 = g.out_f32 != nullptr;
        void* optr = f32 ? (void*)g.out_f32 : (void*)g.out_bf16;
        const int ld = f32 ? g.ld_out32 : g.ld_out16;
        const cuuint64_t gdim[2] = {(cuuint64_t)g.N, (cuuint64_t)g.M};
        const cuuint64_t gstride[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
        const cuuint32_t box[2] = {32, 32};
        const cuuint32_t estr[2] = {1, 1};
        if ((reinterpret_cast<uintptr_t>(optr) & 15) == 0 && gstride[0] % 16 == 0 &&
            enc(&tmo, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, optr, gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
          static bool attr_o = false;
          if (!attr_o) {
            cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM);
            if (e != cudaSuccess) return (int)e;
            attr_o = true;
          }
          GemmNoExt none;
          gemm_tc_kernel<BN, false, true, true><<<grid, GEMM_THREADS, smem_o, st>>>(core, none, tmap, (int)tiles, rs, wsplit, tmo,
                                                                                     (int)(smem_o - 1024 - slabs));
          return (int)cudaGetLastError();
        }
      }
    }
  }
  typename ExtSel<EXT>::type ext;
  if constexpr (EXT) {
    ext.gate = g.gate; ext.add16 = g.add16; ext.out_pre16 = g.out_pre16; ext.row_scale = g.row_scale;
    ext.gate_mode = g.gate_mode; ext.ld_gate = g.ld_gate; ext.rows_per_scale = g.rows_per_scale; ext.conv_full = g.conv_full;
  }
  gemm_tc_kernel<BN, EXT, TMA><<<grid, GEMM_THREADS, smem, st>>>(core, ext, tmap, (int)tiles, res_stages, wsplit, TmaNone{}, 0);
  return (int)cudaGetLastError();
}

// Tensor map for the A operand, or false when the launch is not eligible (then the gathered cp.async producers are used).
static bool make_a_tensor_map(const MstGemm& g, CUtensorMap* tmap) {
  EncodeTiledFn enc = tma_encoder();
  if (!enc || (reinterpret_cast<uintptr_t>(g.A) & 15) != 0) return false;
  if (g.a_mode == MST_A_PLAIN) {
    if ((g.lda * 2) % 16 != 0) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)g.K, (cuuint64_t)g.M};
    const cuuint64_t gstride[1] = {(cuuint64_t)g.lda * 2};
    const cuuint32_t box[2] = {BK, BM};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(reinterpret_cast<const void*>(g.A)), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  // 3x3 conv through TMA: 64-channel k-blocks inside one tap, tiles = whole image rows (or 128-pixel row segments)
  if (g.a_mode != MST_A_CONV3X3 || g.upsample || g.conv_full || g.Cin % 64 != 0 || g.k_pad != g.K || g.W < 8 || g.M % BM != 0) return false;
  if (!((g.W <= BM && BM % g.W == 0 && g.H % (BM / g.W) == 0 && (g.pad_mode == 0 || BM / g.W <= 4)) || g.W % BM == 0)) return false;
  const int B = g.M / (g.H * g.W);
  const cuuint64_t gdim[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.W * g.Cin * 2, (cuuint64_t)g.H * g.W * g.Cin * 2};
  const cuuint32_t box[4] = {BK, (cuuint32_t)(g.W < BM ? g.W : BM), 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(reinterpret_cast<const void*>(g.A)), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
static int launch_gemm(const MstGemm& g, cudaStream_t st, int wsplit = 0) {
  const bool ext = g.out_pre16 || g.gate || g.add16 || g.row_scale || g.conv_full;
  alignas(64) CUtensorMap tmap;
  const bool tma = make_a_tensor_map(g, &tmap);
  if (ext) return tma ? launch_gemm_ext<BN, true, true>(g, st, tmap, wsplit) : launch_gemm_ext<BN, true, false>(g, st, TmaNone{}, wsplit);
  return tma ? launch_gemm_ext<BN, false, true>(g, st, tmap, wsplit) : launch_gemm_ext<BN, false, false>(g, st, TmaNone{}, wsplit);
}

// ---------------------------------------------------------------- weight packing (tile-blocked, pre-swizzled)
// dst is the exact shared-memory image of every [bn x 64] weight tile: block (n_tile, kb) is contiguous
// (bn*128 bytes) and inside it row r, 16-byte chunk c sits at sw128_offset(r, c) -- so one cp.async.bulk
// lands a ready-to-use SWIZZLE_128B K-major UMMA operand.
template <bool CONV>
__global__ void pack_weight_kernel(const float* __restrict__ w, int N, int K, int Cin, bf16* __restrict__ dst, int n_pad,
                                   int k_pad, int bn) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_pad * k_pad) return;
  const int n = (int)(i / k_pad), k = (int)(i % k_pad);
  float v = 0.f;
  if (n < N && k < K) {
    if (CONV) {
      const int tap = k / Cin, ci = k - tap * Cin;
      v = w[((long long)n * Cin + ci) * 9 + tap];  // [N][Cin][3][3] -> k = (ky*3+kx)*Cin + ci
    } else {
      v = w[(long long)n * K + k];
    }
  }
  const int nkb = k_pad / 64;
  const int n_tile = n / bn, r = n - n_tile * bn;
  const int kb = k / 64, kk = k - kb * 64;
  const int c = kk >> 3, e = kk & 7;
  const long long block = ((long long)n_tile * nkb + kb) * bn * 64;  // elements
  const long long off = block + ((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) / 2 + e;
  dst[off] = __float2bfloat16(v);
}

// The same tile-blocked image from a bf16 matrix that already lives on the device (an ACTIVATION used as the B operand: the keys
// and the values of the regular-MHA decoder variant's global attention).  trans = 0: W[n][k] = src[n * ld + k];
// trans = 1: W[n][k] = src[k * ld + n] (the [T, 2C] value tensor read as the [2C, T] operand of P.V).
__global__ void pack_bf16_kernel(const bf16* __restrict__ src, int N, int K, int ld, int trans, bf16* __restrict__ dst, int n_pad,
                                 int k_pad, int bn) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_pad * k_pad) return;
  // trans: consecutive threads walk n (the contiguous direction of src), otherwise k
  const int n = trans ? (int)(i % n_pad) : (int)(i / k_pad);
  const int k = trans ? (int)(i / n_pad) : (int)(i % k_pad);
  bf16 v = __float2bfloat16(0.f);
  if (n < N && k < K) v = trans ? src[(long long)k * ld + n] : src[(long long)n * ld + k];
  const int nkb = k_pad / 64;
  const int n_tile = n / bn, r = n - n_tile * bn;
  const int kb = k / 64, kk = k - kb * 64;
  const int c = kk >> 3, e = kk & 7;
  const long long block = ((long long)n_tile * nkb + kb) * bn * 64;
  dst[block + ((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) / 2 + e] = v;
}

static int tile_n_for(int n_pad) {
  if (n_pad % 256 == 0) return 256;
  if (n_pad % 128 == 0) return 128;
  if (n_pad % 64 == 0) return 64;
  if (n_pad % 32 == 0) return 32;
  return 16;
}

}  // namespace mst

using namespace mst;

extern "C" int mst_gemm_tile_n(int n_pad) { return n_pad > 0 && n_pad % 16 == 0 ? tile_n_for(n_pad) : MST_ERR_BAD_ARG; }

extern "C" int mst_pack_linear_weight(const float* w, int N, int K, mst_bf16* dst, int n_pad, int k_pad, void* stream) {
  if (!w || !dst || N <= 0 || K <= 0 || n_pad < N || k_pad < K || n_pad % 16 || k_pad % 64) return MST_ERR_BAD_ARG;
  const long long n = (long long)n_pad * k_pad;
  pack_weight_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, N, K, 0, reinterpret_cast<bf16*>(dst),
                                                                                         n_pad, k_pad, tile_n_for(n_pad));
  return (int)cudaGetLastError();
}

extern "C" int mst_pack_bf16_matrix(const mst_bf16* src, int N, int K, int ld, int trans, mst_bf16* dst, int n_pad, int k_pad, void* stream) {
  if (!src || !dst || N <= 0 || K <= 0 || n_pad < N || k_pad < K || n_pad % 16 || k_pad % 64 || ld < (trans ? N : K)) return MST_ERR_BAD_ARG;
  const long long n = (long long)n_pad * k_pad;
  pack_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const bf16*>(src), N, K, ld, trans,
                                                                                  reinterpret_cast<bf16*>(dst), n_pad, k_pad, tile_n_for(n_pad));
  return (int)cudaGetLastError();
}

extern "C" int mst_pack_conv3x3_weight(const float* w, int N, int Cin, mst_bf16* dst, int n_pad, int k_pad, void* stream) {
  if (!w || !dst || N <= 0 || Cin <= 0 || n_pad < N || k_pad < 9 * Cin || n_pad % 16 || k_pad % 64) return MST_ERR_BAD_ARG;
  const long long n = (long long)n_pad * k_pad;
  pack_weight_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, N, 9 * Cin, Cin, reinterpret_cast<bf16*>(dst),
                                                                                        n_pad, k_pad, tile_n_for(n_pad));
  return (int)cudaGetLastError();
}

extern "C" int mst_gemm(const MstGemm* g, void* stream) {
  if (!g || !g->A || !g->Wt) return MST_ERR_BAD_ARG;
  if (g->M <= 0 || g->N <= 0 || g->K <= 0 || g->k_pad % BK != 0 || g->k_pad < g->K) return MST_ERR_BAD_ARG;
  if (g->N % 16 != 0 || g->K % 8 != 0 || g->N > MAX_BIAS) return MST_ERR_UNSUPPORTED;
  if (!g->out_f32 && !g->out_bf16 && !g->out_pre16) return MST_ERR_BAD_ARG;
  if (g->mul && !g->res) return MST_ERR_BAD_ARG;
  if ((g->gate_mode != MST_GATE_NONE) != (g->gate != nullptr) || g->gate_mode < 0 || g->gate_mode > MST_GATE_GELU) return MST_ERR_BAD_ARG;
  if ((g->gate || g->add16) && g->ld_gate % 8 != 0) return MST_ERR_BAD_ARG;
  if (g->row_scale && g->rows_per_scale <= 0) return MST_ERR_BAD_ARG;
  if (g->out_pre16 && (g->ld_out16 % 8 != 0 || g->out_nchw)) return MST_ERR_BAD_ARG;
  if (g->out_nchw && (g->gate || g->add16 || g->row_scale)) return MST_ERR_BAD_ARG;
  if (g->conv_full && (g->a_mode != MST_A_CONV3X3 || g->pad_mode != 0 || g->upsample || g->H < 3 || g->W < 3)) return MST_ERR_BAD_ARG;
  if (g->a_mode == MST_A_PLAIN) {
    if (g->lda % 8 != 0 || g->lda < g->K) return MST_ERR_BAD_ARG;
  } else if (g->a_mode == MST_A_CONV3X3) {
    if (g->Cin % 8 != 0 || g->K != 9 * g->Cin || g->H < 2 || g->W < 2 || g->H > 32767 || g->W > 32767) return MST_ERR_BAD_ARG;
    if (g->M % (g->H * g->W) != 0) return MST_ERR_BAD_ARG;
    if (g->upsample && ((g->H | g->W) & 1)) return MST_ERR_BAD_ARG;
  } else {
    return MST_ERR_BAD_ARG;
  }
  if (g->out_nchw == MST_OUT_IMAGE_U8) return MST_ERR_UNSUPPORTED;  // the row-streaming kernel's epilogue only (mst_conv3x3_rows)
  if (g->out_nchw) {
    if (g->out_nchw != MST_OUT_NCHW_F32 || !g->out_f32 || g->a_mode != MST_A_CONV3X3 || g->n_real <= 0 || g->n_real > g->N || g->res) return MST_ERR_BAD_ARG;
  } else {
    if (g->out_f32 && g->ld_out32 % 4 != 0) return MST_ERR_BAD_ARG;
    if (g->out_bf16 && g->ld_out16 % 8 != 0) return MST_ERR_BAD_ARG;
    if (g->res && g->ld_res % 4 != 0) return MST_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // the tile width is a property of the packed weight (N here must be the n_pad it was packed with)
  switch (tile_n_for(g->N)) {
    case 256: {
      // Small M (the training step's batch-8 GEMMs: 64 row tiles): 256-wide tiles would occupy fewer than half of the SMs.  Run
      // the same packed weights as 128-wide halves (twice the CTAs, half the work each): the kernel addresses half h of a packed
      // 256-row block directly, no re-packing.
      static int sms = 0;
      if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
      }
      const long long tiles256 = (long long)((g->M + BM - 1) / BM) * (g->N / 256);
      if (tiles256 * 2 <= sms) return launch_gemm<128>(*g, st, 1);
      return launch_gemm<256>(*g, st);
    }
    case 128: return launch_gemm<128>(*g, st);
    case 64: return launch_gemm<64>(*g, st);
    case 32: return launch_gemm<32>(*g, st);
    default: return launch_gemm<16>(*g, st);
  }
}
