// 3x3 convolution for the WIDE layers (Cout a multiple of 128 -- or 64: half-filled operand --, Cin a multiple of 64, W in {32, 64, 128}) on sm_100a:
// the CNN decoder's 256->128 @32^2 and 128->128 @64^2 layers (codes/decoder.py:25-37) and VGG-19 conv2_x .. conv4_x
// (codes/loss.py:23-37).  bf16 NHWC in, bf16 NHWC out, fp32 accumulation, zero or reflect padding, optional ReLU, optional nearest-x2
// upsample of the input folded into the row fetch (decoder.py:27).
//
// Why another conv kernel: the gathered implicit GEMM (gemm_tc.cu) fetches every input pixel nine times (once per tap) and the
// whole [Cout x 9 Cin] weight matrix once per 128-pixel tile.  For 128 -> 128 channels that is 576 KB of L2 -> shared-memory
// traffic per tile; over 148 SMs the kernel moved 8.9 TB/s, three quarters of the L2's throughput cap, with the tensor pipe at
// 26 % -- it was L2-bound.  Here
//   * the OUTPUT CHANNELS are the MMA's M dimension (A operand = a [128 x 64] weight k-block, K-major, 128B-swizzled) and the
//     PIXELS OF ONE IMAGE ROW are its N dimension (B operand = that row staged once in shared memory, pixel-major, swizzled):
//     tcgen05 runs at full rate for any N with M = 128, so a 32 / 64 / 128-pixel row is a full-rate MMA and no operand ever
//     straddles two image rows;
//   * tap (ky, kx) is the staged row y + ky - 1 read through a descriptor whose start is moved by kx pixels (the swizzle is a
//     function of the absolute shared-memory address: tools/micro/umma_rowshift.cu) -- no im2col, every input row of a unit is
//     fetched once and used by three output rows x three taps;
//   * a unit = R = 256 / W output rows x 128 output channels = 256 accumulator columns in TMEM; one weight k-block (16 KB)
//     streamed from L2 feeds 4 R MMAs, so the weights are fetched once per 256 pixels instead of once per 128;
//   * the weight k-block is the MMA's A operand and is reused by R instructions per K step.  Read from shared memory it costs
//     4 KB of shared-memory bandwidth per instruction (6 KB per N = 64 MMA = 48 clk for 32 clk of math: measured 54).  So each
//     k-block is copied ONCE into tensor memory (tcgen05.cp 128x256b, 32 columns per k-block, eight k-blocks in flight) and
//     the MMAs take A from TMEM; the shared-memory stage returns to the streamer as soon as the copy has read it;
//   * with A out of the way a single issuing warp is the limit (~7 uniform-datapath instructions per MMA at ~7 clk each for
//     a lone warp): TWO warps issue, one the even and one the odd rows of the unit -- disjoint accumulator columns, so the
//     order in which the tensor pipe takes their instructions does not matter.  Warp 0 also issues the copies (one k-block
//     ahead of its MMAs) and publishes each through an mbarrier (tcgen05.commit) that warp 1 waits on.
// TMEM: 32-pixel rows (N = 32 MMAs: 16 clk with A in TMEM against 40 from shared memory): accumulator columns 0..255, weight
// k-blocks in columns 256..511 (one accumulator buffer: the epilogue's drain is not overlapped).  64- and 128-pixel rows: the copy
// would cost what it saves (768 clk per k-block either way at N = 64), so the weights stay in shared memory, TWO accumulators of
// 256 columns alternate and unit u+1 is issued while the epilogue drains unit u.  L2 -> SM traffic per 256 pixels of a
// 128 -> 128 layer: 288 KB of weights + 110 KB of rows instead of 1152 KB.
//
//   warps 0-7   epilogue  : TMEM (lane = output channel, column = pixel) -> + bias -> ReLU -> bf16 -> global; a warp's store
//                           instruction writes 32 consecutive channels of one pixel (64 contiguous bytes)
//   warps 8-12  producers : padded input rows, 128 (or 64) channels at a time, into the row ring with cp.async (reflect / zero
//                           padding in the source address; per-thread offsets computed once)
//   warp  13    weight streamer: one cp.async.bulk per 16 KB stage out of the packed weight image (mst_pack_conv3x3_weight)
//   warps 14,15 MMA issuers: per unit and channel slice: 9 taps x CS/64 weight k-blocks x R/2 rows each x 4 k-steps; rows are
//                           released to the producers when both have issued their last tap row
#include "../../include/mst_b200.h"
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mst {

constexpr int CM_EPI_WARPS = 8;
constexpr int CM_PROD_WARPS = 5;
constexpr int CM_WSTREAM_WARP = CM_EPI_WARPS + CM_PROD_WARPS;  // 13
constexpr int CM_MMA_WARP = CM_WSTREAM_WARP + 1;               // 14 and 15
constexpr int CM_THREADS = 16 * 32;
constexpr int CM_MAX_RING = 12;
constexpr int CM_WSTAGES = 4;
constexpr int CM_WSTAGE_BYTES = 128 * 128;  // [128 output channels x 64 k] bf16, SWIZZLE_128B image
constexpr int CM_ACC_COLS = 256;            // accumulator columns per unit (R rows x W pixels)
constexpr int CM_TSTAGES = 8;               // weight k-blocks resident in TMEM (32 columns each, columns 256..511)

struct CmGeom {
  int R;              // output rows per unit
  int ring;           // row-ring slots
  int plane_bytes;    // one 64-channel plane of a padded row: ((W + 2) rounded up to 8 pixels) x 128 B
  int slot_bytes;     // planes x plane_bytes
  int nslices;        // Cin / CS
  int yblocks;        // H / R
  int ctiles;         // Cout / 128
  int units;          // B x yblocks x ctiles
  int units_per_cta;
  int bn, nkb;        // tile width and k-blocks of the packed weight image
  int prof;           // MST_CM_PROF=1: in-kernel wait clocks of the MMA warp (experiments)
};

MST_DEVINL void cm_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
MST_DEVINL void cm_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// CS = channels per slice: 128 (two 64-channel planes per staged row; W <= 64) or 64 (one plane; W <= 128)
// RR = rows per unit (256 / W), a template parameter so that the issue loops unroll completely: measured, a lone warp pays ~7 clk
// per uniform-datapath instruction, and an N = 64 MMA with its A operand in TMEM takes 32 clk (tools/micro/umma_rate.cu), so
// the instruction stream between two MMAs has to be a handful of adds.
template <int CS, int RR>
__global__ void __launch_bounds__(CM_THREADS, 1) conv_cm_kernel(const GemmCore p, const CmGeom g) {
  constexpr int PLANES = CS / 64;
  constexpr int CPP = CS / 8;  // 16-byte chunks per pixel of a slice
  // Weight k-block through TMEM only when it is reused often enough: the copy runs at 64 B/clk in the same pipe as the MMAs
  // (256 clk per 16 KB k-block), a shared-memory A operand costs 32 clk per MMA (tools/micro/umma_rate.cu).  Per k-block:
  // RR = 8 (N = 32): 32 x 16 + 256 = 768 clk against 32 x 40 = 1280; RR = 4 (N = 64): 768 against 768; RR = 2 (N = 128): 8 x 64
  // + 256 = 768 against 8 x 64 = 512 (the N = 128 MMA is at its math floor with both operands in shared memory).
  constexpr bool TS = RR >= 8;
  // accumulators: the TMEM k-block slots take columns 256..511, so the TMEM path has ONE accumulator buffer (the epilogue's drain is
  // not overlapped); with the weights read from shared memory both halves of TMEM are accumulators and unit u+1 is issued while
  // the epilogue drains unit u.  At N = 64 the two forms cost the same per k-block (768 clk, see above): the double buffer decides.
  constexpr int NACC = TS ? 1 : 2;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t row_full[CM_MAX_RING], row_free[CM_MAX_RING];
  __shared__ uint64_t w_full[CM_WSTAGES], w_empty[CM_WSTAGES];
  __shared__ uint64_t a_ready[CM_TSTAGES], a_free[CM_TSTAGES];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform for ptxas
  const uint32_t w_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring_base = w_base + CM_WSTAGES * CM_WSTAGE_BYTES;
  const int Wp = p.W + 2;
  const int u_begin = blockIdx.x * g.units_per_cta;
  const int u_end = min(u_begin + g.units_per_cta, g.units);
  const int rows_per_item = g.R + 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.ring; ++s) {
      mbar_init(smem_u32(&row_full[s]), CM_PROD_WARPS * 32);
      mbar_init(smem_u32(&row_free[s]), 2);
    }
    for (int s = 0; s < CM_WSTAGES; ++s) {
      mbar_init(smem_u32(&w_full[s]), 1);
      mbar_init(smem_u32(&w_empty[s]), TS ? 1 : 2);
    }
    for (int s = 0; s < CM_TSTAGES; ++s) {
      mbar_init(smem_u32(&a_ready[s]), 1);
      mbar_init(smem_u32(&a_free[s]), 2);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&acc_full[b]), 2);
      mbar_init(smem_u32(&acc_empty[b]), CM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (p.N < 128) {  // 64 output channels: operand rows 64..127 of every weight stage are zeros
    for (uint32_t i = threadIdx.x; i < (uint32_t)(CM_WSTAGES * CM_WSTAGE_BYTES) / 16u; i += CM_THREADS)
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(w_base + i * 16u), "r"(0u) : "memory");
    fence_proxy_async_smem();
  }
  if (warp == CM_MMA_WARP) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= CM_EPI_WARPS && warp < CM_WSTREAM_WARP) {
    // =========================== row producers ===========================
    const int t = threadIdx.x - CM_EPI_WARPS * 32;
    constexpr int NPROD = CM_PROD_WARPS * 32;
    constexpr int MAXW = CS == 128 ? 64 : 128;
    constexpr int MAXK = ((MAXW + 2) * CPP + NPROD - 1) / NPROD;
    const int row_chunks = Wp * CPP;
    // the same copies for every row: chunk idx = t + k * NPROD -> (padded column, channel chunk)
    int src_off[MAXK];
    uint32_t dst_off[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
      const int idx = t + k * NPROD;
      const int col = idx / CPP;
      const int c = idx - col * CPP;
      int xx = col - 1;
      bool valid = idx < row_chunks;
      if (p.pad_mode == 1) xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
      else valid = valid && (unsigned)xx < (unsigned)p.W;
      if (p.upsample) xx >>= 1;  // nearest x2 (decoder.py:27): the stored input is [B, H/2, W/2, Cin]
      src_off[k] = valid ? xx * p.Cin + c * 8 : -1;
      // plane (64 channels) -> pixel row of 128 B -> chunk XOR the pixel's low three bits (planes are 1024-byte aligned)
      dst_off[k] = (uint32_t)(c >> 3) * (uint32_t)g.plane_bytes + (uint32_t)col * 128u + (uint32_t)(((c & 7) ^ (col & 7)) << 4);
    }
    const int nk = (row_chunks - t + NPROD - 1) / NPROD;
    const bf16* Abase = reinterpret_cast<const bf16*>(p.A);
    const int Hs = p.upsample ? (p.H >> 1) : p.H, Ws = p.upsample ? (p.W >> 1) : p.W;  // stored input size
    int slot = 0;
    uint32_t fphase = 1;  // a fresh row_free passes a wait on parity 1
    for (int u = u_begin; u < u_end; ++u) {
      const int rest = u / g.ctiles;
      const int b = rest / g.yblocks, y0 = (rest - b * g.yblocks) * g.R;
      for (int sl = 0; sl < g.nslices; ++sl) {
        for (int j = 0; j < rows_per_item; ++j) {
          int yy = y0 - 1 + j;
          bool vrow = true;
          if (p.pad_mode == 1) yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
          else vrow = (unsigned)yy < (unsigned)p.H;
          if (p.upsample) yy >>= 1;
          const bf16* rowsrc = Abase + ((long long)b * Hs + (vrow ? yy : 0)) * Ws * p.Cin + sl * CS;
          mbar_wait(smem_u32(&row_free[slot]), fphase);
          const uint32_t rowdst = ring_base + (uint32_t)slot * (uint32_t)g.slot_bytes;
#pragma unroll
          for (int k = 0; k < MAXK; ++k) {
            if (k < nk) {
              const bool valid = vrow && src_off[k] >= 0;
              cp_async16(rowdst + dst_off[k], valid ? rowsrc + src_off[k] : Abase, valid);
            }
          }
          cp_async_mbar_arrive_noinc(smem_u32(&row_full[slot]));
          if (++slot == g.ring) { slot = 0; fphase ^= 1; }
        }
      }
    }
    cp_async_wait_all();
  } else if (warp == CM_WSTREAM_WARP) {
    // =========================== weight streamer ===========================
    if (lane == 0) {
      const bf16* wt = reinterpret_cast<const bf16*>(p.Wt);
      // a 64-channel layer (decoder.py:37) fills the lower half of every [128 x 64] stage; the upper half stays zero (zeroed once)
      const uint32_t w_bytes = p.N >= 128 ? (uint32_t)CM_WSTAGE_BYTES : (uint32_t)(p.N * 128);
      int ws = 0;
      uint32_t wphase = 1;
      for (int u = u_begin; u < u_end; ++u) {
        const int ct = u % g.ctiles;
        const int n_tile = (ct * 128) / g.bn, sub = ((ct * 128) % g.bn) / 128;
        const bf16* wtile = wt + (long long)n_tile * g.nkb * g.bn * 64 + (long long)sub * 128 * 64;
        for (int sl = 0; sl < g.nslices; ++sl) {
          for (int tap = 0; tap < 9; ++tap) {
            for (int pl = 0; pl < PLANES; ++pl) {
              const int kb = (tap * p.Cin + sl * CS + pl * 64) >> 6;
              mbar_wait(smem_u32(&w_empty[ws]), wphase);
              cm_arrive_expect_tx(smem_u32(&w_full[ws]), w_bytes);
              cm_bulk_g2s(w_base + ws * CM_WSTAGE_BYTES, wtile + (long long)kb * g.bn * 64, w_bytes, smem_u32(&w_full[ws]));
              if (++ws == CM_WSTAGES) { ws = 0; wphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp >= CM_MMA_WARP) {
    // =========================== MMA issuers (two warps) ===========================
    // each warp whole, convergent, on warp-uniform values; one lane is elected inside the *_pred helpers
    const int mw = warp - CM_MMA_WARP;  // 0: even rows + the weight copies; 1: odd rows
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.W >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 8 rows x 128 B, version 1, SWIZZLE_128B
    const uint32_t plane16 = (uint32_t)g.plane_bytes >> 4, slot16 = (uint32_t)g.slot_bytes >> 4;
    const uint32_t ring_lo = ((ring_base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w_lo = ((w_base & 0x3FFFFu) >> 4) | (1u << 16);
    const int stages_total = (u_end - u_begin) * g.nslices * 9 * PLANES;
    // the weight k-blocks are ONE linear sequence of stages: stage q sits in shared-memory slot q % CM_WSTAGES and goes to
    // TMEM slot q % CM_TSTAGES
    int q_cp = 0;                        // (warp 0) next stage to copy
    auto copy_stage = [&]() {            // shared memory -> tensor memory, then hand the shared-memory stage back
      const int ws = q_cp % CM_WSTAGES, ts = q_cp % CM_TSTAGES;
      mbar_wait(smem_u32(&w_full[ws]), (uint32_t)((q_cp / CM_WSTAGES) & 1));
      mbar_wait(smem_u32(&a_free[ts]), (uint32_t)(((q_cp / CM_TSTAGES) & 1) ^ 1));  // both warps' MMAs of the slot's previous k-block are done
      tc_fence_after();
      const uint32_t a_lo = w_lo + (uint32_t)((ws * CM_WSTAGE_BYTES) >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) tmem_cp_128x256b_pred(tmem_base + CM_ACC_COLS + ts * 32 + k * 8, ((uint64_t)desc_hi << 32) | (a_lo + k * 2));
      umma_commit_pred(smem_u32(&w_empty[ws]));
      umma_commit_pred(smem_u32(&a_ready[ts]));
      ++q_cp;
    };
    // ring position of row 0 of the current item and the parity of that slot's current use; row j of the item sits j slots
    // further (one wrap at most: an item has fewer rows than the ring) -- no division in the issue loop
    int base_slot = 0;
    uint32_t base_phase = 0;
    auto slot_of = [&](int j, uint32_t& phase) {
      int sidx = base_slot + j;
      const bool wrapped = sidx >= g.ring;
      if (wrapped) sidx -= g.ring;
      phase = base_phase ^ (wrapped ? 1u : 0u);
      return sidx;
    };
    uint32_t d_row[RR / 2];  // accumulator columns of this warp's rows
#pragma unroll
    for (int i = 0; i < RR / 2; ++i) d_row[i] = tmem_base + (uint32_t)((2 * i + mw) * p.W);
    int q = 0;  // current stage
    int ucount = 0;
    long long t_rows = 0, t_w = 0, t_acc = 0, t_all = g.prof ? clock64() : 0, t0 = 0;
    if (TS && mw == 0 && stages_total > 0) copy_stage();
    for (int u = u_begin; u < u_end; ++u, ++ucount) {
      if (g.prof) t0 = clock64();
      const int buf = ucount % NACC, use = ucount / NACC;
      const uint32_t dbuf = (uint32_t)(buf * CM_ACC_COLS);
      mbar_wait(smem_u32(&acc_empty[buf]), (uint32_t)((use & 1) ^ 1));
      if (g.prof) t_acc += clock64() - t0;
      tc_fence_after();
      for (int sl = 0; sl < g.nslices; ++sl) {
        int rows_waited = 0;
        for (int ky = 0; ky < 3; ++ky) {
          if (g.prof) t0 = clock64();
          for (; rows_waited < ky + g.R; ++rows_waited) {  // rows ky .. ky + R - 1 of the item feed this tap row
            uint32_t ph;
            const int sidx = slot_of(rows_waited, ph);
            mbar_wait(smem_u32(&row_full[sidx]), ph);
          }
          if (g.prof) t_rows += clock64() - t0;
          fence_proxy_async_smem();  // the rows were written by cp.async (generic proxy)
          tc_fence_after();
          // descriptor low words of this warp's rows for this tap row (ring slot of item row r + ky), once per ky
          uint32_t rb[RR / 2];
#pragma unroll
          for (int i = 0; i < RR / 2; ++i) {
            uint32_t ph_unused;
            rb[i] = ring_lo + (uint32_t)slot_of(2 * i + mw + ky, ph_unused) * slot16;
          }
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int pl = 0; pl < PLANES; ++pl, ++q) {
              const uint32_t acc = (sl | ky | kx | pl) != 0;  // 0: the unit's first k-block overwrites the accumulator
              if constexpr (TS) {
                const int ts = q % CM_TSTAGES;
                if (mw == 0) {
                  if (q_cp <= q + 1 && q_cp < stages_total) copy_stage();  // one k-block ahead of the MMAs
                } else {
                  if (g.prof) t0 = clock64();
                  mbar_wait(smem_u32(&a_ready[ts]), (uint32_t)((q / CM_TSTAGES) & 1));
                  if (g.prof) t_w += clock64() - t0;
                  tc_fence_after();
                }
                const uint32_t a_tmem = tmem_base + CM_ACC_COLS + ts * 32;
#pragma unroll
                for (int i = 0; i < RR / 2; ++i) {
                  // + kx pixels = kx operand rows of 128 B (8 address units); k-step = 32 B inside the swizzled row
                  const uint32_t b_lo = rb[i] + (uint32_t)pl * plane16 + (uint32_t)(kx * 8);
                  const uint32_t d = d_row[i] + dbuf;
                  umma_ts_pred(d, a_tmem, ((uint64_t)desc_hi << 32) | b_lo, idesc, acc);
                  umma_ts_pred(d, a_tmem + 8, ((uint64_t)desc_hi << 32) | (b_lo + 2), idesc, 1);
                  umma_ts_pred(d, a_tmem + 16, ((uint64_t)desc_hi << 32) | (b_lo + 4), idesc, 1);
                  umma_ts_pred(d, a_tmem + 24, ((uint64_t)desc_hi << 32) | (b_lo + 6), idesc, 1);
                }
                umma_commit_pred(smem_u32(&a_free[ts]));
              } else {
                const int ws = q % CM_WSTAGES;
                if (g.prof) t0 = clock64();
                mbar_wait(smem_u32(&w_full[ws]), (uint32_t)((q / CM_WSTAGES) & 1));
                if (g.prof) t_w += clock64() - t0;
                tc_fence_after();
                const uint32_t a_lo = w_lo + (uint32_t)((ws * CM_WSTAGE_BYTES) >> 4);
#pragma unroll
                for (int i = 0; i < RR / 2; ++i) {
                  const uint32_t b_lo = rb[i] + (uint32_t)pl * plane16 + (uint32_t)(kx * 8);
                  const uint32_t d = d_row[i] + dbuf;
                  umma_bf16_pred(d, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, acc);
                  umma_bf16_pred(d, ((uint64_t)desc_hi << 32) | (a_lo + 2), ((uint64_t)desc_hi << 32) | (b_lo + 2), idesc, 1);
                  umma_bf16_pred(d, ((uint64_t)desc_hi << 32) | (a_lo + 4), ((uint64_t)desc_hi << 32) | (b_lo + 4), idesc, 1);
                  umma_bf16_pred(d, ((uint64_t)desc_hi << 32) | (a_lo + 6), ((uint64_t)desc_hi << 32) | (b_lo + 6), idesc, 1);
                }
                umma_commit_pred(smem_u32(&w_empty[ws]));  // both issuing warps arrive: the stage is free when both are done
              }
            }
          }
          // rows whose last tap row this was go back to the producers: row ky for ky < 2, rows 2 .. R+1 after ky = 2
          uint32_t ph_unused;
          if (ky < 2) {
            umma_commit_pred(smem_u32(&row_free[slot_of(ky, ph_unused)]));
          } else {
            for (int j = 2; j < rows_per_item; ++j) umma_commit_pred(smem_u32(&row_free[slot_of(j, ph_unused)]));
          }
        }
        base_slot += rows_per_item;
        if (base_slot >= g.ring) { base_slot -= g.ring; base_phase ^= 1u; }
      }
      umma_commit_pred(smem_u32(&acc_full[buf]));
    }
    if (g.prof && blockIdx.x == 1 && lane == 0)
      printf("cm prof mma %d: units %d total %lld wait rows %lld a_ready %lld acc_empty %lld\n", mw, ucount, clock64() - t_all, t_rows, t_w, t_acc);
    tc_fence_before();
  } else {
    // =========================== epilogue (warps 0-7) ===========================
    const int quad = warp & 3, half = warp >> 2;
    bf16* out = reinterpret_cast<bf16*>(p.out_bf16);
    int ucount = 0;
    for (int u = u_begin; u < u_end; ++u, ++ucount) {
      const int ct = u % g.ctiles, rest = u / g.ctiles;
      const int b = rest / g.yblocks, y0 = (rest - b * g.yblocks) * g.R;
      const int ch = ct * 128 + quad * 32 + lane;
      const bool ch_ok = ch < p.N;
      const float bias = (p.bias && ch_ok) ? p.bias[ch] : 0.f;
      const int buf = ucount % NACC, use = ucount / NACC;
      if (lane == 0) mbar_wait(smem_u32(&acc_full[buf]), (uint32_t)(use & 1));
      __syncwarp();
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = half * 128 + cc * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + buf * CM_ACC_COLS + col0, v);
        tmem_wait_ld();
        if (cc == 3) {  // accumulator drained by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
        }
        if (!ch_ok) continue;  // (warp-uniform: a quadrant is 32 channels)
        const int r = col0 / p.W, x0 = col0 - r * p.W;  // 32 consecutive pixels of one output row (W >= 32)
        bf16* op = out + ((long long)(b * p.H + y0 + r) * p.W + x0) * p.ld_out16 + ch;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float xv = __uint_as_float(v[e]) + bias;
          if (p.act == MST_ACT_RELU) xv = fmaxf(xv, 0.f);
          op[(long long)e * p.ld_out16] = __float2bfloat16_rn(xv);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == CM_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int cm_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static int cm_slice(int Cin, int W) { return (Cin % 128 == 0 && W <= 64) ? 128 : 64; }

static bool cm_shape_ok(int N, int Cin, int H, int W) {
  if (N <= 0 || (N % 128 != 0 && N != 64) || Cin <= 0 || Cin % 64 != 0) return false;
  if (N == 64 && Cin % 128 != 0) return false;  // 64 -> 64 / 64 -> 32 layers: the row-streaming kernel (conv_band.cu) is faster (49 vs 62 us)
  if (W != 32 && W != 64 && W != 128) return false;
  // 32-pixel rows make N = 32 MMAs, which are bound by the shared-memory read of their 4 KB weight operand (40 clk for 16 clk of
  // math): measured slower than the gathered GEMM's 256-wide tiles except for a single 128-channel tile (decoder.py:25)
  if (W == 32 && N != 128) return false;
  const int R = CM_ACC_COLS / W;
  return H >= 2 && H % R == 0;
}

static bool plan_cm(const MstGemm& g, CmGeom& out) {
  if (!cm_shape_ok(g.N, g.Cin, g.H, g.W)) return false;
  if (g.k_pad != 9 * g.Cin || g.M % (g.H * g.W) != 0) return false;
  const int CS = cm_slice(g.Cin, g.W);
  out.R = CM_ACC_COLS / g.W;
  out.plane_bytes = (g.W + 2 + 7) / 8 * 8 * 128;
  out.slot_bytes = (CS / 64) * out.plane_bytes;
  const long long budget = 220LL * 1024 - 1024 - (long long)CM_WSTAGES * CM_WSTAGE_BYTES;
  long long ring = budget / out.slot_bytes;
  if (ring > CM_MAX_RING) ring = CM_MAX_RING;
  if (ring < out.R + 3) return false;  // an item's R + 2 rows and at least one row of the next
  out.ring = (int)ring;
  out.nslices = g.Cin / CS;
  out.yblocks = g.H / out.R;
  out.ctiles = (g.N + 127) / 128;
  const int B = g.M / (g.H * g.W);
  out.units = B * out.yblocks * out.ctiles;
  const int sms = cm_num_sms();
  out.units_per_cta = (out.units + sms - 1) / sms;
  out.bn = mst_gemm_tile_n(g.N);
  out.nkb = g.k_pad / 64;
  { const char* e = getenv("MST_CM_PROF"); out.prof = e ? atoi(e) : 0; }
  return out.bn >= 128 || (g.N == 64 && out.bn == 64);
}

template <int CS, int RR>
static int launch_cm(const MstGemm& g, const CmGeom& geo, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)CM_WSTAGES * CM_WSTAGE_BYTES + (size_t)geo.ring * geo.slot_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_cm_kernel<CS, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const unsigned grid = (unsigned)((geo.units + geo.units_per_cta - 1) / geo.units_per_cta);
  GemmCore core;
  memcpy(&core, &g, sizeof(GemmCore));
  conv_cm_kernel<CS, RR><<<grid, CM_THREADS, smem, st>>>(core, geo);
  return (int)cudaGetLastError();
}

}  // namespace mst

using namespace mst;

extern "C" int mst_conv3x3_cm_supported(int N, int Cin, int H, int W) { return cm_shape_ok(N, Cin, H, W) ? 1 : 0; }

extern "C" int mst_conv3x3_cm(const MstGemm* g, void* stream) {
  if (!g || !g->A || !g->Wt || !g->out_bf16) return MST_ERR_BAD_ARG;
  if (g->a_mode != MST_A_CONV3X3 || g->M <= 0 || g->H <= 0 || g->W <= 0) return MST_ERR_BAD_ARG;
  // what this kernel does not do: fp32 / NCHW outputs, residuals, the training-step extensions
  if (g->out_f32 || g->res || g->mul || g->out_nchw || g->gate || g->add16 || g->out_pre16 || g->row_scale || g->conv_full)
    return MST_ERR_UNSUPPORTED;
  if (g->act != MST_ACT_NONE && g->act != MST_ACT_RELU) return MST_ERR_UNSUPPORTED;
  if (g->pad_mode != 0 && g->pad_mode != 1) return MST_ERR_BAD_ARG;
  if (g->ld_out16 < g->N) return MST_ERR_BAD_ARG;
  CmGeom geo;
  if (!plan_cm(*g, geo)) return MST_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int cs = cm_slice(g->Cin, g->W);
  if (g->W == 32) return cs == 128 ? launch_cm<128, 8>(*g, geo, st) : launch_cm<64, 8>(*g, geo, st);
  if (g->W == 64) return cs == 128 ? launch_cm<128, 4>(*g, geo, st) : launch_cm<64, 4>(*g, geo, st);
  return launch_cm<64, 2>(*g, geo, st);  // W == 128
}
