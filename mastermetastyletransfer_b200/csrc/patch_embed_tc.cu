// Patch embedding of the torchvision Swin encoder (features[0]: Conv2d(3, 128, 4, 4) + LayerNorm(128)) with the first block's
// norm1 fused behind it, on tcgen05 (round 2).  The kernel writes 100 MB (fp32 x + bf16 LN1(x)) for 25 MB read.  With each lane
// storing its own token row (32-byte sectors of 32 different rows per instruction) it took 44-49 us per launch at batch 32 / 256^2,
// the SAME as the mma.sync kernel of norm_misc.cu it replaces, with 8 or with 16 epilogue warps -- which read as "write-bound at
// 2.3 TB/s"; it was the store PATTERN: through TMA tensor stores (below) the same kernel takes 37 us (2.7 TB/s of stores).
//
// The 4x4 / stride-4 conv is a [tokens x 48] . [48 x 128] GEMM, k = ci*16 + ky*4 + kx.  Per 128-token tile:
//   warps 16-19 producers: thread = token; twelve 16-byte loads straight from the NCHW image (one per input channel and patch
//               row: consecutive threads read consecutive 16 bytes), each value split hi + lo into two bf16 (16 mantissa bits of
//               the fp32 image survive) and written into the swizzled K-major A tile: k-block 0 = the 48 hi parts (+16 zeros),
//               k-block 1 = the 48 lo parts; the B tile holds the bf16 weights twice, so  A.B^T = (hi + lo).W
//   warp 20     MMA issuer: 8 x tcgen05.mma (M = 128, N = 128, K = 16) into one of four TMEM accumulators
//   warps 0-15  epilogue, four groups of four (one tile each, round robin: the LayerNorm arithmetic, ~20 operations per element
//               and pass, is what the kernel costs, so it gets the warps): thread = token, the whole 128-channel row is in the
//               thread's TMEM lane, so both LayerNorms are thread-local (no shuffles, no exchange): shifted one-pass statistics,
//               then  x = LN0(acc + bias) -> fp32 [T,128]  and  y = LN1(x) -> bf16 [T,128]  with the parameters read as
//               warp-uniform shared-memory broadcasts; rows leave through TMA tensor stores (32-byte sectors per lane without them)
//
// Output through TMA tensor stores (end of round 2): a lane owns a token row, so direct stores touch 32 different rows per
// instruction.  Each epilogue warp stages its [32 tokens x 32 channels] block in a private 4 KB slab (fp32: 128-byte rows, 128B
// swizzle; bf16: 64-byte rows, 64B swizzle -- both conflict-free for row-per-lane 16-byte writes) and one lane hands it to the TMA
// engine; rows past the last token are clipped by the tensor map.  The weight tile is stored once (both k-blocks read the same 16 KB),
// which pays for the slabs.
#include "../../include/mst_b200.h"
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace mst {

constexpr int PT_EPI_WARPS = 16;                            // four groups of four: four tiles in their epilogue at a time
constexpr int PT_NBUF = PT_EPI_WARPS / 4;                   // A-tile buffers = TMEM accumulators = epilogue groups
constexpr int PT_PROD_WARP0 = PT_EPI_WARPS;
constexpr int PT_PROD_WARPS = 4;
constexpr int PT_MMA_WARP = PT_PROD_WARP0 + PT_PROD_WARPS;  // 12
constexpr int PT_THREADS = (PT_MMA_WARP + 1) * 32;          // 672
constexpr int PT_KB_BYTES = 128 * 128;                      // one [128 rows x 64 k] bf16 k-block
constexpr int PT_A_BYTES = 2 * PT_KB_BYTES;                 // hi | lo
constexpr int PT_SLAB_BYTES = 4096;                         // one staging slab per epilogue warp
constexpr int PT_SMEM_BYTES = 1024 + PT_NBUF * PT_A_BYTES + PT_KB_BYTES + PT_EPI_WARPS * PT_SLAB_BYTES;  // A buffers + weight tile + slabs

MST_DEVINL void pt_tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
MST_DEVINL void pt_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
MST_DEVINL void pt_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// uint8 [B,S,S,3] image input (img8 != nullptr): the producers read the interleaved bytes and apply ToTensor + Normalize themselves
// (the values images_u8_to_nchw_kernel would have written, norm_misc.cu), so the fp32 NCHW image never exists in HBM.
struct PeU8 {
  const uint32_t* img8;
  float mean[3], sd[3];
  int normalize;
};

__global__ void __launch_bounds__(PT_THREADS, 1) patch_embed_tc_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                                      const float* __restrict__ bias, const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta, float* __restrict__ out,
                                                                      const float* __restrict__ gamma1, const float* __restrict__ beta1,
                                                                      bf16* __restrict__ y16, int S, int n_tiles, long long total,
                                                                      PeU8 u8, const __grid_constant__ CUtensorMap tm_x,
                                                                      const __grid_constant__ CUtensorMap tm_y, int tma) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float lut[3][256];  // u8 images: ((v / 255) - mean_c) / std_c for every byte value, in images_u8_to_nchw's operation order
  __shared__ uint64_t a_full[PT_NBUF], a_empty[PT_NBUF], acc_full[PT_NBUF], acc_empty[PT_NBUF];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float prm[5][128];  // bias, gamma, beta, gamma1, beta1

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t a_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = a_base + PT_NBUF * PT_A_BYTES;
  uint8_t* gen = smem_raw + (a_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    for (int b = 0; b < PT_NBUF; ++b) {
      mbar_init(smem_u32(&a_full[b]), PT_PROD_WARPS * 32);
      mbar_init(smem_u32(&a_empty[b]), 1);
      mbar_init(smem_u32(&acc_full[b]), 1);
      mbar_init(smem_u32(&acc_empty[b]), 4);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 128; i += PT_THREADS) {
    prm[0][i] = bias[i]; prm[1][i] = gamma[i]; prm[2][i] = beta[i];
    prm[3][i] = y16 ? gamma1[i] : 0.f; prm[4][i] = y16 ? beta1[i] : 0.f;
  }
  if (u8.img8)
    for (int i = threadIdx.x; i < 3 * 256; i += PT_THREADS) {
      const int c = i >> 8;
      float v = __fdiv_rn((float)(i & 255), 255.0f);
      const float mc = c == 0 ? u8.mean[0] : c == 1 ? u8.mean[1] : u8.mean[2], sc = c == 0 ? u8.sd[0] : c == 1 ? u8.sd[1] : u8.sd[2];
      if (u8.normalize) v = __fdiv_rn(__fsub_rn(v, mc), sc);
      lut[c][i & 255] = v;
    }
  // weight tile: row n (output channel) = W[n][0..47] then zeros; both k-blocks of A (hi parts, lo parts) multiply this one tile
  for (int i = threadIdx.x; i < 128 * 8; i += PT_THREADS) {
    const int n = i >> 3, ch = i & 7;  // 16-byte chunk ch of the row's 64 k
    const int kb = 0;
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    if (ch < 6) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float* wp = w + n * 48 + ch * 8 + 2 * e;
        __nv_bfloat162 h = __floats2bfloat162_rn(wp[0], wp[1]);
        pk[e] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    *reinterpret_cast<uint4*>(gen + PT_NBUF * PT_A_BYTES + kb * PT_KB_BYTES + sw128_offset(n, ch)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  // the A tiles' chunks 6, 7 of every row (k = 48..63) are never written by the producers: zero them once
  for (int i = threadIdx.x; i < PT_NBUF * 2 * 128 * 2; i += PT_THREADS) {
    const int buf = i >> 9, rem = i & 511, kb = rem >> 8, r = (rem >> 1) & 127, ch = 6 + (rem & 1);
    *reinterpret_cast<uint4*>(gen + buf * PT_A_BYTES + kb * PT_KB_BYTES + sw128_offset(r, ch)) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  if (warp == PT_MMA_WARP) {
    tmem_alloc(smem_u32(&tmem_base_slot), 128 * PT_NBUF);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int n_my = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int P = S >> 2;
  const long long plane = (long long)S * S;

  if (warp >= PT_PROD_WARP0 && warp < PT_MMA_WARP) {
    // =========================== producers: thread = token ===========================
    const int t = threadIdx.x - PT_PROD_WARP0 * 32;
    const uint32_t rowoff = (uint32_t)((t >> 3) * 1024 + (t & 7) * 128);
    for (int lt = 0; lt < n_my; ++lt) {
      const long long tile = blockIdx.x + (long long)lt * gridDim.x;
      const int buf = lt % PT_NBUF, u = lt / PT_NBUF;
      if (u >= 1) mbar_wait(smem_u32(&a_empty[buf]), (u - 1) & 1);
      const long long tok = tile * 128 + t;
      uint8_t* ab = gen + buf * PT_A_BYTES;
      if (tok < total) {
        const long long rowi = tok / P;  // b * P + py
        const int px = (int)(tok - rowi * P);
        const long long b = rowi / P;
        const int py = (int)(rowi - b * P);
        float4 v[12];
        if (u8.img8) {  // 4 rows x 12 bytes (4 pixels x RGB) -> the same 12 float4 (channel, row)
          const uint32_t* q0 = u8.img8 + ((b * S + (long long)py * 4) * S * 3) / 4 + px * 3;
          uint32_t wv[12];
#pragma unroll
          for (int ky = 0; ky < 4; ++ky)
#pragma unroll
            for (int e = 0; e < 3; ++e) wv[ky * 3 + e] = __ldg(q0 + (long long)ky * (S * 3 / 4) + e);
#pragma unroll
          for (int ky = 0; ky < 4; ++ky)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
              float f[4];
#pragma unroll
              for (int kx = 0; kx < 4; ++kx) {
                const int byte = 3 * kx + ci;
                f[kx] = lut[ci][(wv[ky * 3 + (byte >> 2)] >> ((byte & 3) * 8)) & 255u];
              }
              v[ci * 4 + ky] = make_float4(f[0], f[1], f[2], f[3]);
            }
        } else {
          const float* p0 = img + (b * 3 * S + (long long)py * 4) * S + px * 4;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) v[ci * 4 + ky] = *reinterpret_cast<const float4*>(p0 + ci * plane + (long long)ky * S);
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) {  // k = 4 j .. 4 j + 3: half of 16-byte chunk j >> 1
          const float f[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          uint32_t hi[2], lo[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * e]), h1 = __float2bfloat16_rn(f[2 * e + 1]);
            __nv_bfloat162 hh; hh.x = h0; hh.y = h1;
            __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * e] - __bfloat162float(h0), f[2 * e + 1] - __bfloat162float(h1));
            hi[e] = *reinterpret_cast<uint32_t*>(&hh);
            lo[e] = *reinterpret_cast<uint32_t*>(&ll);
          }
          const uint32_t off = rowoff + ((uint32_t)((j >> 1) ^ (t & 7)) << 4) + (uint32_t)(j & 1) * 8u;
          *reinterpret_cast<uint2*>(ab + off) = make_uint2(hi[0], hi[1]);
          *reinterpret_cast<uint2*>(ab + PT_KB_BYTES + off) = make_uint2(lo[0], lo[1]);
        }
      } else {  // rows past the last token: zeros (their results are not stored)
#pragma unroll
        for (int ch = 0; ch < 6; ++ch) {
          const uint32_t off = rowoff + ((uint32_t)(ch ^ (t & 7)) << 4);
          *reinterpret_cast<uint4*>(ab + off) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(ab + PT_KB_BYTES + off) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      fence_proxy_async_smem();  // generic-proxy stores, read by the tensor core
      mbar_arrive(smem_u32(&a_full[buf]));
    }
  } else if (warp == PT_MMA_WARP) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    for (int lt = 0; lt < n_my; ++lt) {
      const int buf = lt % PT_NBUF, u = lt / PT_NBUF;
      mbar_wait(smem_u32(&a_full[buf]), u & 1);
      mbar_wait(smem_u32(&acc_empty[buf]), (u & 1) ^ 1);
      tc_fence_after();
      const uint32_t ab = a_base + buf * PT_A_BYTES;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_pred(tmem_base + buf * 128, umma_desc_sw128(ab + kb * PT_KB_BYTES + k * 32), umma_desc_sw128(b_base + k * 32),
                         idesc, (kb | k) != 0);
      umma_commit_pred(smem_u32(&a_empty[buf]));
      umma_commit_pred(smem_u32(&acc_full[buf]));
    }
    tc_fence_before();
  } else {
    // =========================== epilogue: group = warp >> 2 takes tiles lt = group, group + 2, ... ===========================
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t slab = b_base + PT_KB_BYTES + (uint32_t)warp * PT_SLAB_BYTES;
    for (int lt = grp; lt < n_my; lt += PT_NBUF) {
      const long long tile = blockIdx.x + (long long)lt * gridDim.x;
      const int buf = lt % PT_NBUF, u = lt / PT_NBUF;  // buf == grp
      const long long tok = tile * 128 + row;
      if (lane == 0) mbar_wait(smem_u32(&acc_full[buf]), u & 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t acc = lane_addr + buf * 128;
      // ---- pass 1: statistics of x0 = acc + bias (shifted by the row's first value: one pass, well conditioned) ----
      float k0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(acc + c0, v);
        tmem_wait_ld();
        if (c0 == 0) k0 = __uint_as_float(v[0]) + prm[0][0];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float d = __uint_as_float(v[e]) + prm[0][c0 + e] - k0;
          s1 += d;
          s2 = fmaf(d, d, s2);
        }
      }
      const float ms = s1 * (1.f / 128.f);
      const float mean0 = k0 + ms;
      const float rstd0 = rsqrtf(fmaxf(s2 * (1.f / 128.f) - ms * ms, 0.f) + 1e-5f);
      // ---- pass 2: x = LN0(x0) -> fp32 out; statistics of x for the fused norm1 ----
      float k1 = 0.f, t1 = 0.f, t2 = 0.f;
      float* op = out + tok * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(acc + c0, v);
        tmem_wait_ld();
        float x[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = (__uint_as_float(v[e]) + prm[0][c0 + e] - mean0) * rstd0 * prm[1][c0 + e] + prm[2][c0 + e];
        if (c0 == 0) k1 = x[0];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float d = x[e] - k1;
          t1 += d;
          t2 = fmaf(d, d, t2);
        }
        if (tma) {  // [32 tokens x 32 channels] fp32 through the slab: 128-byte rows, 128B swizzle
          if (lane == 0) pt_bulk_wait_read();  // the slab's previous store has been read
          __syncwarp();
          const uint32_t srow = slab + (uint32_t)lane * 128u;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(srow + ((uint32_t)(q ^ (lane & 7)) << 4)), "f"(x[4 * q]),
                         "f"(x[4 * q + 1]), "f"(x[4 * q + 2]), "f"(x[4 * q + 3]) : "memory");
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            pt_tma_store_2d(&tm_x, slab, c0, (int)(tile * 128 + quad * 32));
            pt_bulk_commit();
          }
        } else if (tok < total) {
#pragma unroll
          for (int e = 0; e < 4; ++e) st_global_256f(op + c0 + 8 * e, x + 8 * e);
        }
      }
      if (y16) {
        const float ms1 = t1 * (1.f / 128.f);
        const float mean1 = k1 + ms1;
        const float rstd1 = rsqrtf(fmaxf(t2 * (1.f / 128.f) - ms1 * ms1, 0.f) + 1e-5f);
        bf16* yp = y16 + tok * 128;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(acc + c0, v);
          tmem_wait_ld();
          if (c0 == 96) {  // accumulator drained by this warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
          }
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float z[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = c0 + 2 * e + h;
              const float x = (__uint_as_float(v[2 * e + h]) + prm[0][c] - mean0) * rstd0 * prm[1][c] + prm[2][c];
              z[h] = (x - mean1) * rstd1 * prm[3][c] + prm[4][c];
            }
            __nv_bfloat162 hh = __floats2bfloat162_rn(z[0], z[1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&hh);
          }
          if (tma) {  // [32 tokens x 32 channels] bf16: 64-byte rows, 64B swizzle (16-byte chunk ^= (row >> 1) & 3)
            if (lane == 0) pt_bulk_wait_read();
            __syncwarp();
            const uint32_t srow = slab + (uint32_t)lane * 64u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(srow + ((uint32_t)(q ^ ((lane >> 1) & 3)) << 4)), "r"(pk[4 * q]),
                           "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              pt_tma_store_2d(&tm_y, slab, c0, (int)(tile * 128 + quad * 32));
              pt_bulk_commit();
            }
          } else if (tok < total) {
            uint32_t a8[8], b8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { a8[e] = pk[e]; b8[e] = pk[8 + e]; }
            st_global_256(yp + c0, a8);
            st_global_256(yp + c0 + 16, b8);
          }
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
      }
    }
    if (tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA exits
    tc_fence_before();
  }
  __syncthreads();
  if (warp == PT_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128 * PT_NBUF);
  }
}

typedef CUresult (*PtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PtEncodeTiledFn pt_tma_encoder() {
  static PtEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PtEncodeTiledFn>(q);
  }
  return fn;
}

// Called first by mst_patch_embed_ln (norm_misc.cu); handled = false leaves the call to the other kernels.
// img8 != nullptr: uint8 [B,S,S,3] images with mean3 / std3 (host floats; mean3 == nullptr stops after / 255) instead of img.
int patch_embed_tc_try(const float* img, const float* w, const float* b, const float* gamma, const float* beta, float* x,
                       const float* gamma1, const float* beta1, bf16* y16, int B, int S, cudaStream_t st, bool& handled,
                       const uint8_t* img8, const float* mean3, const float* std3) {
  handled = false;
  static int allow = -1;
  if (allow < 0) { const char* e = getenv("MST_PATCH_EMBED_TC"); allow = e ? atoi(e) : 1; }  // 0: the mma.sync kernel (experiments)
  if ((!allow && !img8) || S % 16 != 0) return 0;  // 16-byte image loads: px * 4 floats at 16-byte alignment needs S % 4; rows of S floats: S % 4
  if ((img8 ? (reinterpret_cast<uintptr_t>(img8) & 3) : (reinterpret_cast<uintptr_t>(img) & 15)) || (reinterpret_cast<uintptr_t>(x) & 31) ||
      (reinterpret_cast<uintptr_t>(y16) & 31))
    return 0;
  handled = true;
  PeU8 u8{};
  u8.img8 = reinterpret_cast<const uint32_t*>(img8);
  if (img8 && mean3 && std3) {
    u8.normalize = 1;
    for (int c = 0; c < 3; ++c) { u8.mean[c] = mean3[c]; u8.sd[c] = std3[c]; }
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long total = (long long)B * (S / 4) * (S / 4);
  const long long tiles = (total + 127) / 128;
  if (tiles > 0x7fffffffLL) return MST_ERR_BAD_ARG;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  // tensor maps of the two outputs ([128 channels, tokens], box 32 x 32); without the driver entry point the lanes store directly
  alignas(64) CUtensorMap tm_x, tm_y;
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_y, 0, sizeof(tm_y));
  int tma = 0;
  static int tma_allow = -1;
  if (tma_allow < 0) { const char* e = getenv("MST_PATCH_EMBED_TMA"); tma_allow = e ? atoi(e) : 1; }
  if (tma_allow && total * 128 <= 0x7fffffffLL * 64) {
    if (PtEncodeTiledFn enc = pt_tma_encoder()) {
      const cuuint64_t gdim[2] = {128, (cuuint64_t)total};
      const cuuint32_t box[2] = {32, 32};
      const cuuint32_t estr[2] = {1, 1};
      const cuuint64_t gs32[1] = {128 * 4}, gs16[1] = {128 * 2};
      tma = enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gs32, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      if (tma && y16)
        tma = enc(&tm_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y16, gdim, gs16, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
  }
  patch_embed_tc_kernel<<<grid, PT_THREADS, PT_SMEM_BYTES, st>>>(img, w, b, gamma, beta, x, gamma1, beta1, y16, S, (int)tiles, total, u8, tm_x, tm_y,
                                                                 tma);
  return (int)cudaGetLastError();
}

}  // namespace mst
