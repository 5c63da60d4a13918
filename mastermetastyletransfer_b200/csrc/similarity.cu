// Similarity loss of the paper (codes/utils.py:105-133 get_scaled_self_cosine_distance_map_lower_triangle, codes/loss.py:137-146,
// 321-336) as a tensor-core contraction that never materialises the B x N x N maps:
//
//     D[i][j]  = cos(a_i, a_j) = ahat_i . ahat_j,      ahat_i = a_i / max(|a_i|, 1e-8)          (torch.cosine_similarity)
//     S[i][j]  = D[i][j] / (sum_k D[k][j] + 1e-6)                                                (column-normalised)
//     term     = mean over ALL B*N*N entries of |tril(S_content, -1) - tril(S_output, -1)|       (or the square)
//
// for the relu3_1 (N = 4096 tokens at 256x256) and relu4_1 taps.  D is symmetric, so its column sums need no N x N pass:
// sum_k D[k][j] = ahat_j . (sum_k ahat_k) -- one vector per image.  Three kernels:
//   sim_normalize_kernel : ahat (bf16, the tensor-core operand) and the per-image sum vector s = sum_k bf16(ahat_k) (fp32 atomics)
//   sim_colscale_kernel  : inv_cs[j] = 1 / (bf16(ahat_j) . s + 1e-6)   (the sums of exactly the values the tensor core multiplies)
//   sim_tile_kernel      : one CTA per (image, 128x128 tile of the lower triangle): D_content and D_output tiles by tcgen05.mma
//                          into two TMEM accumulators (operands by TMA, 3-stage ring), epilogue = |D_c*inv_c[j] - D_o*inv_o[j]|
//                          over i > j, reduced to one partial per CTA; mst_sim_finalize adds the partials in fp64.
// (The reference broadcasts a [B, N, N, C] tensor inside cosine_similarity -- 17 TB at N = 4096 -- and, as written, compares the
// content map with itself, loss.py:333-334, i.e. returns 0: SURVEY 8f-3.)
#include "../../include/mst_b200.h"
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace mst {

constexpr int SM_THREADS = 192;          // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA
constexpr int SM_STAGES = 3;
constexpr int SM_OP_BYTES = 128 * 128;   // one [128 rows x 64 channels] bf16 k-block
constexpr int SM_STAGE_BYTES = 4 * SM_OP_BYTES;
constexpr int SM_SMEM_BYTES = SM_STAGES * SM_STAGE_BYTES + 1024;

MST_DEVINL void sm_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
MST_DEVINL void sm_tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

// one warp per token row: norm, ahat (bf16), and the block's contribution to the per-image sum vector.  A lane always owns the
// same channels (8 consecutive ones per 256-channel group), so the sums stay in registers until the block is done.
constexpr int SM_MAX_GROUPS = 8;  // C <= 2048
__global__ void __launch_bounds__(256) sim_normalize_kernel(const bf16* __restrict__ feat, bf16* __restrict__ ahat, float* __restrict__ svec, int N,
                                                            int C, int rows_per_block) {
  extern __shared__ float s_acc[];  // [C]
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_acc[c] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_end = min(N, (int)(blockIdx.x + 1) * rows_per_block);
  float acc[SM_MAX_GROUPS][8];
#pragma unroll
  for (int m = 0; m < SM_MAX_GROUPS; ++m)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[m][e] = 0.f;
  for (int r = blockIdx.x * rows_per_block + warp; r < r_end; r += 8) {
    const bf16* row = feat + ((long long)b * N + r) * C;
    bf16* orow = ahat + ((long long)b * N + r) * C;
    uint4 v[SM_MAX_GROUPS];
    float ss = 0.f;
#pragma unroll
    for (int m = 0; m < SM_MAX_GROUPS; ++m) {
      if (m * 256 < C) {
        v[m] = *reinterpret_cast<const uint4*>(row + m * 256 + lane * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[m]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
        }
      }
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
#pragma unroll
    for (int m = 0; m < SM_MAX_GROUPS; ++m) {
      if (m * 256 < C) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[m]);
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          oh[e] = __floats2bfloat162_rn(f.x * inv, f.y * inv);
          const float2 g = __bfloat1622float2(oh[e]);  // the rounded values are what the tensor core sums
          acc[m][2 * e] += g.x;
          acc[m][2 * e + 1] += g.y;
        }
        *reinterpret_cast<uint4*>(orow + m * 256 + lane * 8) = o;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < SM_MAX_GROUPS; ++m)
    if (m * 256 < C) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&s_acc[m * 256 + lane * 8 + e], acc[m][e]);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&svec[(long long)b * C + c], s_acc[c]);
}

__global__ void __launch_bounds__(256) sim_colscale_kernel(const bf16* __restrict__ ahat, const float* __restrict__ svec, float* __restrict__ inv_cs,
                                                           int N, int C) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= N) return;
  const bf16* row = ahat + ((long long)b * N + r) * C;
  const float* s = svec + (long long)b * C;
  float d = 0.f;
  for (int c = lane * 8; c < C; c += 256) {
    const uint4 v = *reinterpret_cast<const uint4*>(row + c);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h[e]);
      d = fmaf(f.x, s[c + 2 * e], fmaf(f.y, s[c + 2 * e + 1], d));
    }
  }
  d = warp_sum(d);
  if (lane == 0) inv_cs[(long long)b * N + r] = 1.0f / (d + 1e-6f);
}

// grid = B * nb*(nb+1)/2 CTAs (nb = N/128): tile (ti, tj) with ti >= tj of image b
__global__ void __launch_bounds__(SM_THREADS, 1) sim_tile_kernel(const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_o,
                                                                 const float* __restrict__ inv_c, const float* __restrict__ inv_o,
                                                                 float* __restrict__ partials, const int N, const int C, const int squared,
                                                                 const int tiles_per_image) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[SM_STAGES], empty_bar[SM_STAGES], acc_full;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float red[4];
  __shared__ float invs[2][128];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b = blockIdx.x / tiles_per_image;
  int t = blockIdx.x - b * tiles_per_image;
  // (ti, tj), ti >= tj, from the linear lower-triangle index t = ti*(ti+1)/2 + tj
  int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int nkb = C / 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SM_STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&acc_full), 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 128) {
    invs[0][threadIdx.x] = inv_c[(long long)b * N + tj * 128 + threadIdx.x];
    invs[1][threadIdx.x] = inv_o[(long long)b * N + tj * 128 + threadIdx.x];
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(&tmem_base_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 4) {
    if (lane == 0) {
      const int row_i = b * N + ti * 128, row_j = b * N + tj * 128;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % SM_STAGES, u = kb / SM_STAGES;
        if (u >= 1) mbar_wait(smem_u32(&empty_bar[s]), (u - 1) & 1);
        const uint32_t st = base + s * SM_STAGE_BYTES, bar = smem_u32(&full_bar[s]);
        sm_arrive_expect_tx(bar, SM_STAGE_BYTES);
        sm_tma_load_2d(st, &tm_c, kb * 64, row_i, bar);
        sm_tma_load_2d(st + SM_OP_BYTES, &tm_c, kb * 64, row_j, bar);
        sm_tma_load_2d(st + 2 * SM_OP_BYTES, &tm_o, kb * 64, row_i, bar);
        sm_tma_load_2d(st + 3 * SM_OP_BYTES, &tm_o, kb * 64, row_j, bar);
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % SM_STAGES, u = kb / SM_STAGES;
      mbar_wait(smem_u32(&full_bar[s]), u & 1);
      tc_fence_after();
      const uint32_t st = base + s * SM_STAGE_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_bf16_pred(tmem_base, umma_desc_sw128(st + k * 32), umma_desc_sw128(st + SM_OP_BYTES + k * 32), idesc, (kb | k) != 0);
        umma_bf16_pred(tmem_base + 128, umma_desc_sw128(st + 2 * SM_OP_BYTES + k * 32), umma_desc_sw128(st + 3 * SM_OP_BYTES + k * 32), idesc,
                       (kb | k) != 0);
      }
      umma_commit_pred(smem_u32(&empty_bar[s]));
    }
    umma_commit_pred(smem_u32(&acc_full));
    tc_fence_before();
  } else {
    // epilogue: thread = row i of the tile (TMEM lane), columns j in chunks of 32
    const int i_loc = warp * 32 + lane;
    if (lane == 0) mbar_wait(smem_u32(&acc_full), 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool diag = ti == tj;
    float acc = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      if (diag && c0 > warp * 32) break;  // columns j > every row of this warp
      uint32_t vc[32], vo[32];
      tmem_ld32(lane_addr + c0, vc);
      tmem_ld32(lane_addr + 128 + c0, vo);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int j = c0 + e;
        float d = __uint_as_float(vc[e]) * invs[0][j] - __uint_as_float(vo[e]) * invs[1][j];
        d = squared ? d * d : fabsf(d);
        if (!diag || j < i_loc) acc += d;
      }
    }
    tc_fence_before();
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) partials[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// out[0] = sum over taps of  (sum of the tap's partials) / count   (fp64 accumulation, one block)
__global__ void __launch_bounds__(256) sim_finalize_kernel(const float* __restrict__ p0, int n0, double inv_count0, const float* __restrict__ p1, int n1,
                                                           double inv_count1, float* __restrict__ out) {
  __shared__ double red[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n0; i += 256) a += (double)p0[i];
  for (int i = threadIdx.x; i < n1; i += 256) b += (double)p1[i];
  red[threadIdx.x] = a * inv_count0 + b * inv_count1;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)red[0];
}

typedef CUresult (*SmEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static SmEncodeTiledFn sm_tma_encoder() {
  static SmEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<SmEncodeTiledFn>(q);
  }
  return fn;
}

static bool sm_make_map(const void* ptr, long long rows, int C, CUtensorMap* tm) {
  SmEncodeTiledFn enc = sm_tma_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)C * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mst

using namespace mst;

extern "C" int mst_sim_num_tiles(int B, int N) {
  if (B <= 0 || N <= 0 || N % 128 != 0) return MST_ERR_BAD_ARG;
  const long long nb = N / 128, t = (long long)B * nb * (nb + 1) / 2;
  return t > 0x7fffffffLL ? MST_ERR_BAD_ARG : (int)t;
}

extern "C" int mst_sim_prepare(const mst_bf16* feat, int B, int N, int C, mst_bf16* ahat, float* svec, float* inv_cs, void* stream) {
  if (!feat || !ahat || !svec || !inv_cs || B <= 0 || N <= 0 || C <= 0 || C % 256 != 0 || C > 2048) return MST_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(ahat)) & 15) return MST_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(svec, 0, (size_t)B * C * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const int rows_per_block = 64;
  dim3 g1((N + rows_per_block - 1) / rows_per_block, B);
  sim_normalize_kernel<<<g1, 256, C * sizeof(float), st>>>(reinterpret_cast<const bf16*>(feat), reinterpret_cast<bf16*>(ahat), svec, N, C, rows_per_block);
  dim3 g2((N + 7) / 8, B);
  sim_colscale_kernel<<<g2, 256, 0, st>>>(reinterpret_cast<const bf16*>(ahat), svec, inv_cs, N, C);
  return (int)cudaGetLastError();
}

extern "C" int mst_sim_tiles(const mst_bf16* ahat_c, const float* inv_cs_c, const mst_bf16* ahat_o, const float* inv_cs_o, int B, int N, int C,
                             int squared, float* partials, int n_partials, void* stream) {
  if (!ahat_c || !ahat_o || !inv_cs_c || !inv_cs_o || !partials) return MST_ERR_BAD_ARG;
  const int tiles = mst_sim_num_tiles(B, N);
  if (tiles <= 0 || n_partials < tiles || C % 64 != 0 || C <= 0) return MST_ERR_BAD_ARG;
  alignas(64) CUtensorMap tm_c, tm_o;
  if (!sm_make_map(ahat_c, (long long)B * N, C, &tm_c) || !sm_make_map(ahat_o, (long long)B * N, C, &tm_o)) return MST_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sim_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int nb = N / 128;
  sim_tile_kernel<<<tiles, SM_THREADS, SM_SMEM_BYTES, (cudaStream_t)stream>>>(tm_c, tm_o, inv_cs_c, inv_cs_o, partials, N, C, squared, nb * (nb + 1) / 2);
  return (int)cudaGetLastError();
}

extern "C" int mst_sim_finalize(const float* partials0, int n0, double count0, const float* partials1, int n1, double count1, float* out, void* stream) {
  if (!partials0 || n0 <= 0 || count0 <= 0 || !out || (n1 > 0 && (!partials1 || count1 <= 0))) return MST_ERR_BAD_ARG;
  sim_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials0, n0, 1.0 / count0, partials1, n1 > 0 ? n1 : 0, n1 > 0 ? 1.0 / count1 : 0.0, out);
  return (int)cudaGetLastError();
}
