// Weight-gradient GEMM for sm_100a (training step):
//     dW[n, k'] += sum_m dY[m, n] * X(m, k')
// for nn.Linear (X(m,k') = X[m*ld_x + k']) and 3x3 Conv2d (X(m,k') = in[b, pad(y+ky-1), pad(x+kx-1), ci] with
// k' = (ky*3+kx)*Cin + ci -- the same gather as the forward implicit GEMM, gemm_tc.cu).
//
// The reduction runs over TOKENS, and both operands are token-major in HBM (a token's channels are contiguous),
// i.e. "MN-major" UMMA operands.  A stage holds 64 tokens: dY as two [64 tokens x 64 channels] panels and X as
// BN/64 such panels, each panel = 64 rows of 128 bytes with the 16-byte chunk index XOR-ed by (token & 7) --
// the canonical SWIZZLE_128B MN-major layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units with
// LBO = one panel (8 KB) and SBO = 1024 B, so a token row lands with plain 16-byte cp.async copies and no transpose.
// tcgen05.mma (M=128 dY channels, N=BN X channels, K=16 tokens) accumulates fp32 in TMEM.
//
// Split-K: grid = n_tiles x k_tiles x splits; each CTA reduces its token range and adds its [128 x BN] tile to dW
// with coalesced fp32 RED (the tile is transposed through shared memory first), so shared weights used several
// times per step (the style transformer's shared MHA) simply accumulate.
//
//   warps 0-3  : epilogue (TMEM lane quadrant each)      warps 4-11 : producers (cp.async gathers)
//   warp 12    : TMEM allocation + MMA issue
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

constexpr int WG_BM = 128;
constexpr int WG_TOK = 64;
constexpr int WG_PANEL = WG_TOK * 128;  // bytes of one [64 tokens x 64 channels] bf16 panel
constexpr int WG_EPI_WARPS = 4;
constexpr int WG_PROD_WARPS = 8;
constexpr int WG_THREADS = (WG_EPI_WARPS + WG_PROD_WARPS + 1) * 32;
constexpr int WG_A_BYTES = 2 * WG_PANEL;

template <int BN>
struct WgCfg {
  static constexpr int B_PANELS = BN / 64;
  static constexpr int B_BYTES = B_PANELS * WG_PANEL;
  static constexpr int STAGE_BYTES = WG_A_BYTES + B_BYTES;
  static constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

// MN-major SWIZZLE_128B shared-memory descriptor: LBO = byte distance between 64-element groups along M/N,
// SBO = byte distance between 8-row groups along K (1024: eight 128-byte token rows).
MST_DEVINL uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const MstWgrad p, const int n_tiles, const int k_tiles,
                                                                 const int splits, const int blocks_per_split) {
  using Cfg = WgCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform for ptxas
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  int bid = blockIdx.x;
  const int split = bid % splits; bid /= splits;
  const int k_tile = bid % k_tiles;
  const int n_tile = bid / k_tiles;
  const int n0 = n_tile * WG_BM, k0 = k_tile * BN;
  const int total_blocks = (p.M + WG_TOK - 1) / WG_TOK;
  const int blk0 = split * blocks_per_split;
  int nblk = total_blocks - blk0;
  if (nblk > blocks_per_split) nblk = blocks_per_split;
  if (nblk <= 0) return;  // uniform per CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), WG_PROD_WARPS * 32);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    mbar_fence_init();
  }
  if (warp == WG_EPI_WARPS + WG_PROD_WARPS) {
    tmem_alloc(smem_u32(&tmem_base_slot), BN < 32 ? 32 : BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= WG_EPI_WARPS && warp < WG_EPI_WARPS + WG_PROD_WARPS) {
    // =========================== producers ===========================
    const int t = threadIdx.x - WG_EPI_WARPS * 32;
    const int r = t >> 2;    // token row inside the stage (0..63)
    const int sub = t & 3;   // this thread's chunks: sub and sub + 4 of every panel
    const bf16* dYb = reinterpret_cast<const bf16*>(p.dY);
    const bf16* Xb = reinterpret_cast<const bf16*>(p.X);
    const bool conv = p.x_mode != MST_A_PLAIN;
    const int Hs = p.upsample ? (p.H >> 1) : p.H;
    const int Ws = p.upsample ? (p.W >> 1) : p.W;
    const int hw = conv ? p.H * p.W : 1;
    // per-chunk constants (do not depend on the token): channel offsets and conv taps
    int a_col[4];
    bool a_ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + (j >> 1) * 64 + (sub + 4 * (j & 1)) * 8;
      a_col[j] = col;
      a_ok[j] = col < p.N;
    }
    int b_col[2 * Cfg::B_PANELS], b_tap[2 * Cfg::B_PANELS];
    bool b_ok[2 * Cfg::B_PANELS];
#pragma unroll
    for (int j = 0; j < 2 * Cfg::B_PANELS; ++j) {
      const int kk = k0 + (j >> 1) * 64 + (sub + 4 * (j & 1)) * 8;
      b_ok[j] = kk < p.K;
      if (conv) {
        const int tap = kk / p.Cin;
        b_tap[j] = tap;
        b_col[j] = kk - tap * p.Cin;
      } else {
        b_tap[j] = 0;
        b_col[j] = kk;
      }
    }
    const uint32_t dst_lo = sw128_offset(r, sub), dst_hi = sw128_offset(r, sub + 4);
    int stage = 0;
    uint32_t pphase = 1;
    for (int blk = 0; blk < nblk; ++blk) {
      const int m = (blk0 + blk) * WG_TOK + r;
      const bool m_ok = m < p.M;
      const int s = stage;
      mbar_wait(smem_u32(&empty_bar[s]), pphase);
      if (++stage == STAGES) { stage = 0; pphase ^= 1; }
      const uint32_t a_stage = smem_base + s * Cfg::STAGE_BYTES;
      const uint32_t b_stage = a_stage + WG_A_BYTES;
      const bf16* dyrow = dYb + (long long)(m_ok ? m : 0) * p.ld_dy;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool v = m_ok && a_ok[j];
        cp_async16(a_stage + (j >> 1) * WG_PANEL + ((j & 1) ? dst_hi : dst_lo), v ? dyrow + a_col[j] : dYb, v);
      }
      if (!conv) {
        const bf16* xrow = Xb + (long long)(m_ok ? m : 0) * p.ld_x;
#pragma unroll
        for (int j = 0; j < 2 * Cfg::B_PANELS; ++j) {
          const bool v = m_ok && b_ok[j];
          cp_async16(b_stage + (j >> 1) * WG_PANEL + ((j & 1) ? dst_hi : dst_lo), v ? xrow + b_col[j] : Xb, v);
        }
      } else {
        int yo[3] = {0, 0, 0}, xo[3] = {0, 0, 0};
        uint32_t vmask = 0;
        const bf16* img = Xb;
        if (m_ok) {
          const int b = m / hw;
          const int rem = m - b * hw;
          const int y = rem / p.W;
          const int x = rem - y * p.W;
          img = Xb + (long long)b * Hs * Ws * p.Cin;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            int yy = y + j - 1, xx = x + j - 1;
            bool vy = true, vx = true;
            if (p.pad_mode == 1) {
              yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
              xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
            } else {
              vy = (unsigned)yy < (unsigned)p.H;
              vx = (unsigned)xx < (unsigned)p.W;
            }
            if (p.upsample) { yy >>= 1; xx >>= 1; }
            yo[j] = vy ? yy * Ws * p.Cin : 0;
            xo[j] = vx ? xx * p.Cin : 0;
            vmask |= (vy ? 1u : 0u) << j | (vx ? 1u : 0u) << (3 + j);
          }
        }
#pragma unroll
        for (int j = 0; j < 2 * Cfg::B_PANELS; ++j) {
          const int tap = b_tap[j];
          const int ky = tap / 3, kx = tap - ky * 3;
          const int yoff = ky == 0 ? yo[0] : (ky == 1 ? yo[1] : yo[2]);
          const int xoff = kx == 0 ? xo[0] : (kx == 1 ? xo[1] : xo[2]);
          const bool v = b_ok[j] && ((vmask >> ky) & (vmask >> (3 + kx)) & 1u);
          cp_async16_ca(b_stage + (j >> 1) * WG_PANEL + ((j & 1) ? dst_hi : dst_lo), v ? img + (yoff + xoff + b_col[j]) : Xb, v);
        }
      }
      cp_async_mbar_arrive_noinc(smem_u32(&full_bar[s]));
    }
    cp_async_wait_all();
  } else if (warp == WG_EPI_WARPS + WG_PROD_WARPS) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = umma_idesc_bf16(WG_BM, BN) | (1u << 15) | (1u << 16);  // A and B MN-major
    int stage = 0;
    uint32_t cphase = 0;
    for (int blk = 0; blk < nblk; ++blk) {
      const int s = stage;
      mbar_wait(smem_u32(&full_bar[s]), cphase);
      if (++stage == STAGES) { stage = 0; cphase ^= 1; }
      tc_fence_after();
      {  // whole warp, warp-uniform values, one lane elected inside the asm (uniform-register issue loop, see gemm_tc.cu)
        const uint32_t a_stage = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t b_stage = a_stage + WG_A_BYTES;
#pragma unroll
        for (int k = 0; k < WG_TOK / 16; ++k) {  // 16 tokens = two 8-row groups = 2048 bytes
          umma_bf16_pred(tmem_base, umma_desc_mn_sw128(a_stage + k * 2048, WG_PANEL), umma_desc_mn_sw128(b_stage + k * 2048, WG_PANEL),
                         idesc, (blk | k) != 0);
        }
        umma_commit_pred(smem_u32(&empty_bar[s]));
        if (blk == nblk - 1) umma_commit_pred(smem_u32(&tmem_full_bar));
      }
    }
    tc_fence_before();
  } else {
    // =========================== epilogue (warps 0-3) ===========================
    // All MMAs (and therefore all shared-memory operand reads) are complete once tmem_full fires, so the pipeline
    // stages are free: each warp transposes 32x32 fp32 blocks through its own 4.1 KB slice and issues coalesced REDs.
    if (lane == 0) mbar_wait(smem_u32(&tmem_full_bar), 0);
    __syncwarp();
    tc_fence_after();
    float* tr = reinterpret_cast<float*>(smem_raw + ((smem_base - smem_u32(smem_raw)))) + warp * (32 * 33);
    const int quad = warp;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int nrow0 = n0 + quad * 32;
    const int n_rows = p.n_real > 0 ? p.n_real : p.N;
#pragma unroll 1
    for (int col0 = 0; col0 < BN; col0 += 32) {
      if (k0 + col0 >= p.K) break;
      uint32_t v[32];
      tmem_ld32(t_row + col0, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int kk = k0 + col0 + lane;
      if (kk < p.K) {
        long long col_off;
        int row_stride;
        if (p.x_mode == MST_A_PLAIN) {
          col_off = kk;
          row_stride = p.K;
        } else {  // conv weight [N][Cin][3][3]
          const int tap = kk / p.Cin, ci = kk - tap * p.Cin;
          col_off = (long long)ci * 9 + tap;
          row_stride = p.Cin * 9;
        }
        for (int rr = 0; rr < 32; ++rr) {
          const int n = nrow0 + rr;
          if (n >= n_rows) break;
          atomicAdd(p.dW + (long long)n * row_stride + col_off, tr[rr * 33 + lane]);
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == WG_EPI_WARPS + WG_PROD_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

static int wg_num_sms = 0;

template <int BN>
static int launch_wgrad(const MstWgrad& g, cudaStream_t st) {
  using Cfg = WgCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (wg_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&wg_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (wg_num_sms <= 0) wg_num_sms = 148;
  }
  const int n_tiles = (g.N + WG_BM - 1) / WG_BM;
  const int k_tiles = (g.K + BN - 1) / BN;
  const int tiles = n_tiles * k_tiles;
  const int blocks = (g.M + WG_TOK - 1) / WG_TOK;
  // split the token range over the SMs, but keep at least 8 stages (512 tokens) per CTA so the RED epilogue amortises
  int splits = wg_num_sms / tiles;
  if (splits < 1) splits = 1;
  const int max_splits = (blocks + 7) / 8;
  if (splits > max_splits) splits = max_splits;
  const int bps = (blocks + splits - 1) / splits;
  splits = (blocks + bps - 1) / bps;
  wgrad_tc_kernel<BN><<<(unsigned)(tiles * splits), WG_THREADS, Cfg::SMEM_BYTES, st>>>(g, n_tiles, k_tiles, splits, bps);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- column sums (bias gradients)
// out[n] += sum_m dY[m, n].  CTA = 256 rows x 64 columns slab; thread = (8-channel chunk, row lane); smem reduce, then RED.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dY, int M, int N, int ld, float* __restrict__ out,
                                                     int rows_per_cta) {
  __shared__ float red[32][65];
  const int c8 = threadIdx.x & 7, rl = threadIdx.x >> 3;  // 8 chunks x 32 row lanes
  const int col = blockIdx.y * 64 + c8 * 8;
  const int m0 = blockIdx.x * rows_per_cta;
  const int m1 = min(M, m0 + rows_per_cta);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    for (int m = m0 + rl; m < m1; m += 32) {
      const uint4 v = *reinterpret_cast<const uint4*>(dY + (long long)m * ld + col);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] += __uint_as_float(w[e] << 16);
        acc[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][c8 * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    const int n = blockIdx.y * 64 + threadIdx.x;
    if (n < N) atomicAdd(out + n, s);
  }
}

}  // namespace mst

using namespace mst;

extern "C" int mst_wgrad(const MstWgrad* g, void* stream) {
  if (!g || !g->dY || !g->X || !g->dW) return MST_ERR_BAD_ARG;
  if (g->M <= 0 || g->N <= 0 || g->K <= 0 || g->n_real < 0 || g->n_real > g->N) return MST_ERR_BAD_ARG;
  if (g->N % 8 != 0 || g->ld_dy % 8 != 0 || g->ld_dy < g->N) return MST_ERR_BAD_ARG;
  if (g->x_mode == MST_A_PLAIN) {
    if (g->K % 8 != 0 || g->ld_x % 8 != 0 || g->ld_x < g->K) return MST_ERR_BAD_ARG;
  } else if (g->x_mode == MST_A_CONV3X3) {
    if (g->Cin % 8 != 0 || g->K != 9 * g->Cin || g->H < 2 || g->W < 2) return MST_ERR_BAD_ARG;
    if (g->M % (g->H * g->W) != 0) return MST_ERR_BAD_ARG;
    if (g->upsample && ((g->H | g->W) & 1)) return MST_ERR_BAD_ARG;
  } else {
    return MST_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (g->K > 128) return launch_wgrad<256>(*g, st);
  if (g->K > 64) return launch_wgrad<128>(*g, st);
  return launch_wgrad<64>(*g, st);
}

extern "C" int mst_colsum(const mst_bf16* dY, int M, int N, int ld, float* out, void* stream) {
  if (!dY || !out || M <= 0 || N <= 0 || N % 8 != 0 || ld % 8 != 0 || ld < N) return MST_ERR_BAD_ARG;
  int rows_per_cta = 1024;
  dim3 grid((unsigned)((M + rows_per_cta - 1) / rows_per_cta), (unsigned)((N + 63) / 64));
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const bf16*>(dY), M, N, ld, out, rows_per_cta);
  return (int)cudaGetLastError();
}
