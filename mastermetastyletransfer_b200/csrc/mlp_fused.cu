// Fused transformer MLP for sm_100a:  out = res + fc2(GELU(fc1(A) + b1)) + b2   (torchvision MLP, the five
// 256->1024->256 MLPs of a style-transformer layer and the four Swin-encoder MLPs).  The [tokens x 4C] hidden
// activation never touches HBM: per 128-token tile it is produced 128 hidden units at a time into TMEM (fc1),
// pulled through bias + GELU by the epilogue warps into a 128B-swizzled bf16 shared-memory tile, and consumed
// from there as the A operand of the second tcgen05 GEMM, which accumulates the [128 x C] output in TMEM
// across all hidden chunks.
//
//   warps 0-7  : epilogue.  fc1 chunk: TMEM -> +b1 -> GELU -> bf16 -> swizzled smem (Hs, double buffered);
//                tile end: TMEM -> +b2 -> +residual -> fp32 / bf16 global stores.
//   warps 8-11 : A-tile producers (cp.async, [128 x C] bf16 stays resident for the whole tile).
//   warp 12    : MMA issuer (one elected lane): MMA1(t) then MMA2(t-1), so the tensor core always has the next
//                fc1 chunk to chew on while the epilogue warps run GELU on the previous one.
//   warp 13    : weight streamer: both weight matrices are pre-packed as ONE linear stream of 32 KB stages in
//                exactly the order the MMA warp consumes them (each stage is the swizzled smem image), so the
//                producer is a loop of cp.async.bulk copies.
// TMEM: fc1 accumulator double buffered (2 x 128 columns) + fc2 accumulator (C columns) <= 512 columns.
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

constexpr int ML_EPI_WARPS = 8;
constexpr int ML_PROD_WARPS = 4;
constexpr int ML_THREADS = (ML_EPI_WARPS + ML_PROD_WARPS + 2) * 32;
constexpr int ML_HC = 128;               // hidden units per chunk
constexpr int ML_STAGE_BYTES = 32 * 1024;

MST_DEVINL void ml_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
MST_DEVINL void ml_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

template <int C>
struct MlpCfg {
  static constexpr int KB1 = C / 64;                 // k-blocks of fc1 (K = C)
  static constexpr int S1 = C / 128;                 // weight stages per fc1 chunk (two [128 x 64] k-blocks per stage)
  static constexpr int S2 = C == 256 ? 2 : 1;        // weight stages per fc2 chunk
  static constexpr int NSTG = C == 256 ? 2 : 3;      // ring depth
  static constexpr int A_BYTES = KB1 * 128 * 128;    // resident A tile
  static constexpr int HS_BYTES = 2 * 128 * 128;     // one Hs buffer: [128 x 128] bf16 as two k-blocks
  static constexpr int SMEM_BYTES = 1024 + A_BYTES + 2 * HS_BYTES + NSTG * ML_STAGE_BYTES;
  static constexpr int ACC2_COL = 256;
};

template <int C>
__global__ void __launch_bounds__(ML_THREADS, 1) mlp_fused_kernel(const MstMlp p, const int num_tiles) {
  using Cfg = MlpCfg<C>;
  constexpr int NSTG = Cfg::NSTG;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full, a_empty, acc2_full, acc2_empty;
  __shared__ uint64_t w_full[NSTG], w_empty[NSTG];
  __shared__ uint64_t acc1_full[2], acc1_empty[2], hs_full[2], hs_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float b1_s[1024];
  __shared__ float b2_s[256];

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform for ptxas
  const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t hs_base = a_base + Cfg::A_BYTES;
  const uint32_t ring_base = hs_base + 2 * Cfg::HS_BYTES;
  const int hidden = 4 * C;
  const int NCH = hidden / ML_HC;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&a_full), ML_PROD_WARPS * 32);
    mbar_init(smem_u32(&a_empty), 1);
    mbar_init(smem_u32(&acc2_full), 1);
    mbar_init(smem_u32(&acc2_empty), ML_EPI_WARPS);
    for (int s = 0; s < NSTG; ++s) { mbar_init(smem_u32(&w_full[s]), 1); mbar_init(smem_u32(&w_empty[s]), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&acc1_full[b]), 1);
      mbar_init(smem_u32(&acc1_empty[b]), ML_EPI_WARPS);
      mbar_init(smem_u32(&hs_full[b]), ML_EPI_WARPS);
      mbar_init(smem_u32(&hs_empty[b]), 1);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < hidden; i += ML_THREADS) b1_s[i] = p.b1[i];
  for (int i = threadIdx.x; i < C; i += ML_THREADS) b2_s[i] = p.b2[i];
  if (warp == 12) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp >= ML_EPI_WARPS && warp < ML_EPI_WARPS + ML_PROD_WARPS) {
    // =========================== A-tile producers ===========================
    const int t = threadIdx.x - ML_EPI_WARPS * 32;
    const int c = t & 7, r0 = t >> 3;  // rows r0 + 16*i
    const bf16* Abase = reinterpret_cast<const bf16*>(p.A);
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      if (lt >= 1) mbar_wait(smem_u32(&a_empty), (lt - 1) & 1);
      const int m0 = tile * 128;
#pragma unroll
      for (int kb = 0; kb < Cfg::KB1; ++kb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const int m = m0 + r;
          const bool valid = m < p.M;
          cp_async16(a_base + kb * 16384 + sw128_offset(r, c), valid ? Abase + (long long)m * p.lda + kb * 64 + c * 8 : Abase, valid);
        }
      }
      cp_async_mbar_arrive_noinc(smem_u32(&a_full));
    }
    cp_async_wait_all();
  } else if (warp == 13) {
    // =========================== weight streamer ===========================
    if (lane == 0) {
      const int stages_per_tile = NCH * (Cfg::S1 + Cfg::S2);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.Wstream);
      int ws = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int s = 0; s < stages_per_tile; ++s, ++ws) {
          const int slot = ws % NSTG;
          if (ws >= NSTG) mbar_wait(smem_u32(&w_empty[slot]), ((ws / NSTG) - 1) & 1);
          ml_arrive_expect_tx(smem_u32(&w_full[slot]), ML_STAGE_BYTES);
          ml_bulk_g2s(ring_base + slot * ML_STAGE_BYTES, wsrc + (size_t)s * ML_STAGE_BYTES, ML_STAGE_BYTES, smem_u32(&w_full[slot]));
        }
      }
    }
  } else if (warp == 12) {
    // =========================== MMA issuer ===========================
    // whole warp, convergent, warp-uniform values; one lane elected inside umma_bf16_pred / umma_commit_pred so that
    // ptxas keeps the descriptors in uniform registers (see gemm_tc.cu)
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, ML_HC);
    constexpr uint32_t idesc2 = umma_idesc_bf16(128, C);
    int ws = 0, lt = 0;
    auto wait_stage = [&](int& slot) {
      slot = ws % NSTG;
      mbar_wait(smem_u32(&w_full[slot]), (ws / NSTG) & 1);
      tc_fence_after();
    };
    auto mma2 = [&](int j) {  // acc2 += Hs(j) . W2_j^T
      const int gc = lt * NCH + j, buf = j & 1, u = gc >> 1;
      mbar_wait(smem_u32(&hs_full[buf]), u & 1);
      if (j == 0) mbar_wait(smem_u32(&acc2_empty), (lt & 1) ^ 1);
      tc_fence_after();
      const uint32_t hs = hs_base + buf * Cfg::HS_BYTES;
      for (int s = 0; s < Cfg::S2; ++s, ++ws) {
        int slot;
        wait_stage(slot);
        {
          const uint32_t wb = ring_base + slot * ML_STAGE_BYTES;
          if constexpr (C == 256) {  // stage = k-block s of the chunk: [256 x 64]
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_pred(tmem_base + Cfg::ACC2_COL, umma_desc_sw128(hs + s * 16384 + k * 32), umma_desc_sw128(wb + k * 32), idesc2,
                        (j | s | k) != 0);
          } else {  // stage = both k-blocks: 2 x [128 x 64]
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_pred(tmem_base + Cfg::ACC2_COL, umma_desc_sw128(hs + kb * 16384 + k * 32), umma_desc_sw128(wb + kb * 16384 + k * 32),
                          idesc2, (j | kb | k) != 0);
          }
          umma_commit_pred(smem_u32(&w_empty[slot]));
        }
      }
      umma_commit_pred(smem_u32(&hs_empty[buf]));
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      mbar_wait(smem_u32(&a_full), lt & 1);
      tc_fence_after();
      for (int t = 0; t < NCH; ++t) {
        const int gc = lt * NCH + t, buf = t & 1, u = gc >> 1;
        mbar_wait(smem_u32(&acc1_empty[buf]), (u & 1) ^ 1);
        tc_fence_after();
        // ---- MMA1(t): acc1[buf] = A . W1_t^T ----
        for (int s = 0; s < Cfg::S1; ++s, ++ws) {
          int slot;
          wait_stage(slot);
          {
            const uint32_t wb = ring_base + slot * ML_STAGE_BYTES;
#pragma unroll
            for (int kb2 = 0; kb2 < 2; ++kb2) {
              const int kb = s * 2 + kb2;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_pred(tmem_base + buf * ML_HC, umma_desc_sw128(a_base + kb * 16384 + k * 32),
                          umma_desc_sw128(wb + kb2 * 16384 + k * 32), idesc1, (kb | k) != 0);
            }
            umma_commit_pred(smem_u32(&w_empty[slot]));
          }
        }
        umma_commit_pred(smem_u32(&acc1_full[buf]));
        if (t >= 1) mma2(t - 1);
      }
      mma2(NCH - 1);
      umma_commit_pred(smem_u32(&acc2_full));
      umma_commit_pred(smem_u32(&a_empty));
    }
    tc_fence_before();
  } else {
    // =========================== epilogue warps 0-7 ===========================
    const int quad = warp & 3, half = warp >> 2;
    const int row_in_tile = quad * 32 + lane;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int row = tile * 128 + row_in_tile;
      const bool row_ok = row < p.M;
      for (int j = 0; j < NCH; ++j) {
        const int gc = lt * NCH + j, buf = j & 1, u = gc >> 1;
        if (lane == 0) mbar_wait(smem_u32(&acc1_full[buf]), u & 1);
        __syncwarp();
        tc_fence_after();
        // this warp's 64 hidden units = k-block `half` of the Hs tile
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * ML_HC + half * 64;
        uint32_t packed[32];
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          uint32_t v[32];
          tmem_ld32(taddr + part * 32, v);
          tmem_wait_ld();
          if (part == 1) {  // both loads done: the fc1 accumulator can be overwritten by MMA1(j+2)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc1_empty[buf]));
          }
          const float* bb = b1_s + j * ML_HC + half * 64 + part * 32;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float x0 = gelu_erf(__uint_as_float(v[2 * e]) + bb[2 * e]);
            const float x1 = gelu_erf(__uint_as_float(v[2 * e + 1]) + bb[2 * e + 1]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
            packed[part * 16 + e] = *reinterpret_cast<uint32_t*>(&h2);
          }
        }
        if (lane == 0) mbar_wait(smem_u32(&hs_empty[buf]), (u & 1) ^ 1);  // MMA2(j-2) has finished reading this buffer
        __syncwarp();
        const uint32_t hrow = hs_base + buf * Cfg::HS_BYTES + half * 16384;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t addr = hrow + sw128_offset(row_in_tile, c);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(packed[4 * c]), "r"(packed[4 * c + 1]),
                       "r"(packed[4 * c + 2]), "r"(packed[4 * c + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&hs_full[buf]));
      }
      // ---- tile end: fc2 accumulator -> + b2 + residual -> global ----
      if (lane == 0) mbar_wait(smem_u32(&acc2_full), lt & 1);
      __syncwarp();
      tc_fence_after();
      constexpr int CPW = C / 2;
      const bool wide_res = p.res && ((reinterpret_cast<uintptr_t>(p.res) | (uintptr_t)(p.ld_res * 4)) & 31) == 0;
      const bool wide_o32 = p.out_f32 && ((reinterpret_cast<uintptr_t>(p.out_f32) | (uintptr_t)(p.ld_out32 * 4)) & 31) == 0;
      const bool wide_o16 = p.out_bf16 && ((reinterpret_cast<uintptr_t>(p.out_bf16) | (uintptr_t)(p.ld_out16 * 2)) & 31) == 0;
      const uint32_t t2 = tmem_base + ((uint32_t)(quad * 32) << 16) + Cfg::ACC2_COL + half * CPW;
#pragma unroll 1
      for (int col0 = 0; col0 < CPW; col0 += 32) {
        uint32_t v[32];
        tmem_ld32(t2 + col0, v);
        tmem_wait_ld();
        if (col0 + 32 >= CPW) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&acc2_empty));
        }
        if (!row_ok) continue;
        const int n = half * CPW + col0;
        float x[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = __uint_as_float(v[e]) + b2_s[n + e];
        // 256-bit accesses (one 32-byte sector per lane and instruction) when the row segments are 32-byte aligned
        if (p.res) {
          const float* rp = p.res + (long long)row * p.ld_res + n;
          if (wide_res) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float r[8];
              ld_global_256f(rp + 8 * e, r);
#pragma unroll
              for (int q = 0; q < 8; ++q) x[8 * e + q] += r[q];
            }
          } else {
            const float4* r4 = reinterpret_cast<const float4*>(rp);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 r = r4[e];
              x[4 * e] += r.x; x[4 * e + 1] += r.y; x[4 * e + 2] += r.z; x[4 * e + 3] += r.w;
            }
          }
        }
        if (p.out_f32) {
          float* op = p.out_f32 + (long long)row * p.ld_out32 + n;
          if (wide_o32) {
#pragma unroll
            for (int e = 0; e < 4; ++e) st_global_256f(op + 8 * e, x + 8 * e);
          } else {
            float4* o4 = reinterpret_cast<float4*>(op);
#pragma unroll
            for (int e = 0; e < 8; ++e) o4[e] = make_float4(x[4 * e], x[4 * e + 1], x[4 * e + 2], x[4 * e + 3]);
          }
        }
        if (p.out_bf16) {
          bf16* op = reinterpret_cast<bf16*>(p.out_bf16) + (long long)row * p.ld_out16 + n;
          if (wide_o16) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              uint32_t pk[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(x[16 * e + 2 * q], x[16 * e + 2 * q + 1]);
                pk[q] = *reinterpret_cast<uint32_t*>(&h2);
              }
              st_global_256(op + 16 * e, pk);
            }
          } else {
            uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uint32_t pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(x[8 * e + 2 * q], x[8 * e + 2 * q + 1]);
                pk[q] = *reinterpret_cast<uint32_t*>(&h2);
              }
              o4[e] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- weight stream packing
// Stream order (one entry = one 32 KB stage): W1(0); then for t = 1..NCH-1: W1(t), W2(t-1); then W2(NCH-1).
// W1(t) stages: s in [0,S1): k-blocks 2s, 2s+1 of the [128 hidden x C] tile, each a [128 x 64] SW128 image.
// W2(j) stages: C == 256: s in {0,1}: [256 out x 64 hidden] image of hidden units 128j + 64s ..;
//               C == 128: one stage with two [128 out x 64 hidden] images.
__device__ __forceinline__ int mlp_stage_of_w1(int t, int S1, int S2) { return t == 0 ? 0 : t * S1 + (t - 1) * S2; }
__device__ __forceinline__ int mlp_stage_of_w2(int j, int NCH, int S1, int S2) {
  return j == NCH - 1 ? NCH * S1 + (NCH - 1) * S2 : (j + 2) * S1 + j * S2;
}
__global__ void pack_mlp_kernel(const float* __restrict__ w1, const float* __restrict__ w2, bf16* __restrict__ dst, int C) {
  const int hidden = 4 * C, NCH = hidden / ML_HC;
  const int S1 = C / 128, S2 = C == 256 ? 2 : 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n1 = (long long)hidden * C;
  if (i < n1) {  // W1[hid][k]
    const int hid = (int)(i / C), k = (int)(i % C);
    const int t = hid / ML_HC, r = hid % ML_HC;
    const int kb = k / 64, kk = k % 64, c = kk >> 3, e = kk & 7;
    const long long stage = mlp_stage_of_w1(t, S1, S2) + kb / 2;
    const long long off = stage * (ML_STAGE_BYTES / 2) + (kb & 1) * 8192 + ((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) / 2 + e;
    dst[off] = __float2bfloat16(w1[i]);
  } else if (i < 2 * n1) {  // W2[n][hid]
    const long long q = i - n1;
    const int n = (int)(q / hidden), hid = (int)(q % hidden);
    const int j = hid / ML_HC, hh = hid % ML_HC;
    const int kb = hh / 64, kk = hh % 64, c = kk >> 3, e = kk & 7;
    long long stage = mlp_stage_of_w2(j, NCH, S1, S2);
    long long inner = ((n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4)) / 2 + e;
    if (C == 256) stage += kb; else inner += kb * 8192;
    dst[stage * (ML_STAGE_BYTES / 2) + inner] = __float2bfloat16(w2[q]);
  }
}

static int ml_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

template <int C>
static int launch_mlp(const MstMlp& p, cudaStream_t st) {
  using Cfg = MlpCfg<C>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int tiles = (p.M + 127) / 128;
  const unsigned grid = (unsigned)(tiles < ml_num_sms() ? tiles : ml_num_sms());
  mlp_fused_kernel<C><<<grid, ML_THREADS, Cfg::SMEM_BYTES, st>>>(p, tiles);
  return (int)cudaGetLastError();
}

}  // namespace mst

using namespace mst;

extern "C" size_t mst_mlp_stream_bytes(int C) { return (C == 128 || C == 256) ? (size_t)2 * 4 * C * C * 2 : 0; }

extern "C" int mst_pack_mlp_weights(const float* w1, const float* w2, mst_bf16* dst, int C, void* stream) {
  if (!w1 || !w2 || !dst || (C != 128 && C != 256)) return MST_ERR_BAD_ARG;
  const long long n = 2LL * 4 * C * C;
  pack_mlp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w1, w2, reinterpret_cast<bf16*>(dst), C);
  return (int)cudaGetLastError();
}

extern "C" int mst_mlp_fused(const MstMlp* p, void* stream) {
  if (!p || !p->A || !p->Wstream || !p->b1 || !p->b2) return MST_ERR_BAD_ARG;
  if (p->M <= 0 || p->lda % 8 || p->lda < p->C) return MST_ERR_BAD_ARG;
  if (!p->out_f32 && !p->out_bf16) return MST_ERR_BAD_ARG;
  if ((p->out_f32 && p->ld_out32 % 4) || (p->out_bf16 && p->ld_out16 % 8) || (p->res && p->ld_res % 4)) return MST_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->C == 256) return launch_mlp<256>(*p, st);
  if (p->C == 128) return launch_mlp<128>(*p, st);
  return MST_ERR_UNSUPPORTED;
}
