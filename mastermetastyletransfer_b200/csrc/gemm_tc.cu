// Tensor-core GEMM for sm_100a: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) with a gathering A-operand
// producer and a fused epilogue.  One CTA computes a 128 x BN output tile.
//
//   warps 0-3 : producer (cp.async 16-byte gathers into 128B-swizzled K-major smem stages), then epilogue
//               (tcgen05.ld of their 32-lane TMEM quadrant -> bias/activation/residual/blend -> global)
//   warp  4   : TMEM allocation, then one elected lane issues tcgen05.mma and commits to mbarriers
//
// The gather makes the same kernel serve F.linear on token-major activations and 3x3 convolutions as
// implicit GEMM (zero or reflect padding, optional nearest x2 upsample folded into the read).
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * 128;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 3 : 4);
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + slack for manual 1024 B alignment
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN>
__global__ void __launch_bounds__(160) gemm_tc_kernel(const MstGemm p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int LAG = STAGES - 1;  // cp.async groups kept in flight per producer thread

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  const int n_tiles = p.N / BN;
  const int n_tile = blockIdx.x % n_tiles;
  const int m_tile = blockIdx.x / n_tiles;
  const int m0 = m_tile * BM;
  const int n0 = n_tile * BN;
  const int nkb = p.k_pad / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 128);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&tmem_base_slot), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 4) {
    // =========================== producer ===========================
    const int t = threadIdx.x;
    const int c = t & 7;    // 16-byte chunk column inside the 128-byte k-slab
    const int r0 = t >> 3;  // first of this thread's rows; rows r0 + 16*i
    // per-row source bookkeeping
    long long row_base[8];  // PLAIN: element offset of the row; CONV: image base offset (b*Hs*Ws*Cin)
    int row_yx[8];          // CONV: y | x << 16 ; -1 if row >= M
    const int Hs = p.upsample ? (p.H >> 1) : p.H;
    const int Ws = p.upsample ? (p.W >> 1) : p.W;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + r0 + 16 * i;
      if (m >= p.M) {
        row_base[i] = 0;
        row_yx[i] = -1;
      } else if (p.a_mode == MST_A_PLAIN) {
        row_base[i] = (long long)m * p.lda;
        row_yx[i] = 0;
      } else {
        const int hw = p.H * p.W;
        const int b = m / hw;
        const int rem = m - b * hw;
        const int y = rem / p.W;
        const int x = rem - y * p.W;
        row_base[i] = (long long)b * Hs * Ws * p.Cin;
        row_yx[i] = y | (x << 16);
      }
    }
    const uint32_t a_dst0 = sw128_offset(r0, c);  // + i*2048 for row r0+16i (two 8-row groups further)

    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      if (kb >= STAGES) mbar_wait(smem_u32(&empty_bar[s]), ((kb / STAGES) - 1) & 1);
      const uint32_t a_stage = smem_base + s * Cfg::STAGE_BYTES;
      const uint32_t b_stage = a_stage + A_STAGE_BYTES;
      const int k0 = kb * BK + c * 8;
      // ---- A tile: 128 rows x 64 k ----
      if (p.a_mode == MST_A_PLAIN) {
        const bool kvalid = k0 < p.K;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool valid = kvalid && row_yx[i] >= 0;
          const bf16* src = reinterpret_cast<const bf16*>(p.A) + (valid ? row_base[i] + k0 : 0);
          cp_async16(a_stage + a_dst0 + i * 2048, src, valid);
        }
      } else {
        const int tap = k0 / p.Cin;
        const int ch = k0 - tap * p.Cin;
        const int ky = tap / 3, kx = tap - ky * 3;
        const bool kvalid = tap < 9;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          bool valid = kvalid && row_yx[i] >= 0;
          int yy = (row_yx[i] & 0xFFFF) + ky - 1;
          int xx = (row_yx[i] >> 16) + kx - 1;
          if (p.pad_mode == 1) {  // reflect (no edge repeat): -1 -> 1, H -> H-2
            yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
            xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
          } else {
            valid = valid && (unsigned)yy < (unsigned)p.H && (unsigned)xx < (unsigned)p.W;
          }
          if (p.upsample) { yy >>= 1; xx >>= 1; }
          const long long off = valid ? row_base[i] + ((long long)(yy * Ws + xx)) * p.Cin + ch : 0;
          cp_async16(a_stage + a_dst0 + i * 2048, reinterpret_cast<const bf16*>(p.A) + off, valid);
        }
      }
      // ---- B tile: BN rows (output channels) x 64 k, always in range (weights are padded) ----
      {
        const bf16* wsrc = reinterpret_cast<const bf16*>(p.Wt) + (long long)(n0 + r0) * p.k_pad + kb * BK + c * 8;
#pragma unroll
        for (int i = 0; i < (BN + 15) / 16; ++i) {
          if (BN >= 16 || r0 + 16 * i < BN)
            cp_async16(b_stage + a_dst0 + i * 2048, wsrc + (long long)i * 16 * p.k_pad, true);
        }
      }
      cp_async_commit();
      if (kb >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&full_bar[(kb - LAG) % STAGES]));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    for (int kb = (nkb > LAG ? nkb - LAG : 0); kb < nkb; ++kb) mbar_arrive(smem_u32(&full_bar[kb % STAGES]));

    // =========================== epilogue ===========================
    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + lane;
    const bool row_ok = row < p.M;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    long long nchw_base = 0;
    int hw = 0;
    if (p.out_nchw && row_ok) {
      hw = p.H * p.W;
      const int b = row / hw;
      nchw_base = (long long)b * p.n_real * hw + (row - b * hw);
    }
#pragma unroll 1
    for (int col0 = 0; col0 < BN; col0 += 16) {
      uint32_t v[16];
      tmem_ld16(t_row + col0, v);
      tmem_wait_ld();
      if (!row_ok) continue;
      const int n = n0 + col0;
      float x[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        x[j] = __uint_as_float(v[j]);
        if (p.bias) x[j] += __ldg(p.bias + n + j);
        if (p.act == MST_ACT_RELU) x[j] = fmaxf(x[j], 0.0f);
        else if (p.act == MST_ACT_GELU) x[j] = gelu_erf(x[j]);
      }
      if (p.res) {
        const float4* r4 = reinterpret_cast<const float4*>(p.res + (long long)row * p.ld_res + n);
        if (p.mul) {
          const float4* m4 = reinterpret_cast<const float4*>(p.mul + (long long)row * p.ld_res + n);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 r = r4[j], m = m4[j];
            x[4 * j + 0] += r.x * m.x; x[4 * j + 1] += r.y * m.y; x[4 * j + 2] += r.z * m.z; x[4 * j + 3] += r.w * m.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 r = r4[j];
            x[4 * j + 0] += r.x; x[4 * j + 1] += r.y; x[4 * j + 2] += r.z; x[4 * j + 3] += r.w;
          }
        }
      }
      if (p.out_nchw) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n + j < p.n_real) p.out_f32[nchw_base + (long long)(n + j) * hw] = x[j];
      } else {
        if (p.out_f32) {
          float4* o4 = reinterpret_cast<float4*>(p.out_f32 + (long long)row * p.ld_out32 + n);
#pragma unroll
          for (int j = 0; j < 4; ++j) o4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        }
        if (p.out_bf16) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out_bf16) + (long long)row * p.ld_out16 + n);
          o4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          o4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
    tc_fence_before();
  } else {
    // =========================== MMA issuer (warp 4) ===========================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(smem_u32(&full_bar[s]), (kb / STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_stage = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t b_stage = a_stage + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr>>4) field
          const uint64_t ad = umma_desc_sw128(a_stage + k * 32);
          const uint64_t bd = umma_desc_sw128(b_stage + k * 32);
          umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0);
        }
        umma_commit(smem_u32(&empty_bar[s]));                  // smem stage reusable once these MMAs retire
        if (kb == nkb - 1) umma_commit(smem_u32(&accum_bar));  // accumulator complete
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm(const MstGemm& g, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long tiles = (long long)((g.M + BM - 1) / BM) * (g.N / BN);
  if (tiles <= 0 || tiles > 0x7fffffffLL) return MST_ERR_BAD_ARG;
  gemm_tc_kernel<BN><<<(unsigned)tiles, 160, Cfg::SMEM_BYTES, st>>>(g);
  return (int)cudaGetLastError();
}

}  // namespace mst

extern "C" int mst_gemm(const MstGemm* g, void* stream) {
  using namespace mst;
  if (!g || !g->A || !g->Wt) return MST_ERR_BAD_ARG;
  if (g->M <= 0 || g->N <= 0 || g->K <= 0 || g->k_pad % BK != 0 || g->k_pad < g->K) return MST_ERR_BAD_ARG;
  if (g->N % 16 != 0 || g->K % 8 != 0) return MST_ERR_UNSUPPORTED;
  if (!g->out_f32 && !g->out_bf16) return MST_ERR_BAD_ARG;
  if (g->mul && !g->res) return MST_ERR_BAD_ARG;
  if (g->a_mode == MST_A_PLAIN) {
    if (g->lda % 8 != 0 || g->lda < g->K) return MST_ERR_BAD_ARG;
  } else if (g->a_mode == MST_A_CONV3X3) {
    if (g->Cin % 8 != 0 || g->K != 9 * g->Cin || g->H < 2 || g->W < 2 || g->H > 32767 || g->W > 32767) return MST_ERR_BAD_ARG;
    if (g->M % (g->H * g->W) != 0) return MST_ERR_BAD_ARG;
    if (g->upsample && ((g->H | g->W) & 1)) return MST_ERR_BAD_ARG;
  } else {
    return MST_ERR_BAD_ARG;
  }
  if (g->out_nchw) {
    if (!g->out_f32 || g->a_mode != MST_A_CONV3X3 || g->n_real <= 0 || g->n_real > g->N || g->res) return MST_ERR_BAD_ARG;
  } else {
    if (g->out_f32 && g->ld_out32 % 4 != 0) return MST_ERR_BAD_ARG;
    if (g->out_bf16 && g->ld_out16 % 8 != 0) return MST_ERR_BAD_ARG;
    if (g->res && g->ld_res % 4 != 0) return MST_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int N = g->N;
  if (N % 128 == 0) return launch_gemm<128>(*g, st);
  if (N % 64 == 0) return launch_gemm<64>(*g, st);
  if (N % 32 == 0) return launch_gemm<32>(*g, st);
  return launch_gemm<16>(*g, st);
}
