// 3x3 convolution as an im2col-free implicit GEMM on tcgen05 (sm_100a).
//
// A CTA takes a "band": R output rows of one image, full width.  The R+2 input rows it needs (zero / reflect
// padding and an optional nearest x2 upsample applied while loading) are staged ONCE in shared memory as
//     halo[channel chunk of 8][padded pixel]   (16 bytes per entry, pitch Wp = W + 2 pixels per row)
// which is exactly the SWIZZLE_NONE K-major UMMA layout: for one 16-byte channel chunk, 8 consecutive pixels
// are a contiguous 128-byte core matrix, the next 8 pixels are SBO = 128 B further, the next channel chunk is
// LBO = NPX*16 B further.  Output positions are enumerated over the PADDED raster (Wp per row, 2 of Wp are
// discarded), so tap (ky,kx) of 128 consecutive positions is the same operand shifted by ky*Wp + kx pixels:
// nine taps = nine descriptor start addresses, no data movement.  Each input pixel is read from L2/HBM
// (R+2)/R times instead of nine.  Weight tiles stream through a small ring with one bulk copy per k-block
// (same pre-swizzled packing as gemm_tc.cu).  All NS = ceil(R*Wp/128) strips of the band accumulate in TMEM
// (NS*BN <= 512 columns) while the k-blocks stream once.
#include "../../include/mst_b200.h"
#include "common.cuh"
#include <cuda.h>
#include <string.h>
#include <stdlib.h>

namespace mst {

constexpr int CB_THREADS = 256;
constexpr int CB_MAX_BIAS = 1024;

struct BandGeom {
  int R;          // output rows per band
  int NS;         // 128-position strips per band
  int npx;        // allocated halo pixels per channel chunk
  int Wp;         // padded row pitch (W + 2)
  int bands_per_img;
  int n_tiles;    // N / BN
  int wst;        // weight ring stages
  int tmem_cols;  // TMEM columns allocated by the CTA (256 when two CTAs share an SM, else 512)
  int lead;       // halo pixel index of (row 0, padded column 0): 1, or 8 when the halo is filled by tensor copies (128-byte alignment)
  int tma;        // 1: halo rows fetched with cp.async.bulk.tensor (one per 8-channel plane and row)
};

MST_DEVINL void cb_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
MST_DEVINL void cb_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// K-major SWIZZLE_NONE descriptor: 8-row x 16-byte core matrices, LBO between the two K chunks of a K=16 step,
// SBO between consecutive 8-row groups (cute::UMMA::SmemDescriptor, layout type 0).
MST_DEVINL uint64_t umma_desc_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

template <int BN>
__global__ void __launch_bounds__(CB_THREADS, 2) conv_band_kernel(const GemmCore p, const BandGeom g, const int total_units,
                                                                  const __grid_constant__ CUtensorMap tm_halo) {
  constexpr int MAXST = 4;
  constexpr int B_STAGE_BYTES = BN * 128;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAXST];
  __shared__ uint64_t empty_bar[MAXST];
  __shared__ uint64_t accum_bar;
  __shared__ uint64_t halo_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[CB_MAX_BIAS];

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform for ptxas
  const uint32_t ring_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t halo_base = ring_base + g.wst * B_STAGE_BYTES;
  const int cpp = p.Cin >> 3;            // 16-byte chunks per pixel
  const int kpt = p.Cin >> 4;            // K=16 steps per tap
  const int total_ks = 9 * kpt;
  const int nkb = p.k_pad / 64;
  const int Hs = p.upsample ? (p.H >> 1) : p.H;
  const int Ws = p.upsample ? (p.W >> 1) : p.W;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.wst; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    mbar_init(smem_u32(&halo_bar), 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < p.N; i += CB_THREADS) bias_s[i] = p.bias ? p.bias[i] : 0.f;
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const bf16* Abase = reinterpret_cast<const bf16*>(p.A);
  const uint8_t* Wbase = reinterpret_cast<const uint8_t*>(p.Wt);
  constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
#ifdef MST_BAND_PROF
  long long t_fill = 0, t_mma = 0, t_epi = 0, t_mark = clock64();
#endif
  int wit = 0;  // running weight k-block counter (ring position), identical in producer and MMA warps
  int ucount = 0;

  // Issue (do not wait for) the cp.async gathers that stage the (R+2)-row halo of `unit`, with zero / reflect
  // padding and the nearest-x2 upsample applied on the fly.
  auto issue_halo = [&](int unit) {
    const int band = unit / g.n_tiles;
    const int b = band / g.bands_per_img;
    const int y0 = (band - b * g.bands_per_img) * g.R;
    if (g.tma) {
      // one tensor copy per (halo row, 8-channel plane): box = 8 channels x Wp pixels starting at x = -1; out-of-image
      // pixels / rows are zero-filled by the hardware (zero padding), reflect rows use the reflected row coordinate and the
      // two reflect columns are patched from the neighbouring pixels once the copies have landed
      if (threadIdx.x == 0) {
        cb_arrive_expect_tx(smem_u32(&halo_bar), (uint32_t)((g.R + 2) * g.Wp * 16 * cpp));
        for (int hr = 0; hr < g.R + 2; ++hr) {
          int yy = y0 + hr - 1;
          if (p.pad_mode == 1) {
            yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
            yy = min(max(yy, 0), p.H - 1);
          }
          const uint32_t rowdst = halo_base + (uint32_t)(hr * g.Wp + g.lead) * 16u;
          for (int c = 0; c < cpp; ++c)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                             rowdst + (uint32_t)(c * g.npx) * 16u),
                         "l"(&tm_halo), "r"(c * 8), "r"(-1), "r"(yy), "r"(b), "r"(smem_u32(&halo_bar))
                         : "memory");
        }
      }
      return;
    }
    const bf16* img = Abase + (long long)b * Hs * Ws * p.Cin;
    const int row_chunks = g.Wp * cpp;
    for (int hr = 0; hr < g.R + 2; ++hr) {
      int yy = y0 + hr - 1;
      bool vrow = true;
      if (p.pad_mode == 1) {  // reflect (no edge repeat)
        yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
        yy = min(max(yy, 0), p.H - 1);  // rows past the image in a partial last band (their outputs are discarded)
      } else {
        vrow = (unsigned)yy < (unsigned)p.H;
      }
      if (p.upsample) yy >>= 1;
      const bf16* rowsrc = img + (long long)yy * Ws * p.Cin;
      const uint32_t rowdst = halo_base + (uint32_t)(hr * g.Wp + g.lead) * 16u;
      for (int idx = threadIdx.x; idx < row_chunks; idx += CB_THREADS) {
        const int col = idx / cpp;
        const int c = idx - col * cpp;
        int xx = col - 1;
        bool valid = vrow;
        if (p.pad_mode == 1) xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
        else valid = valid && (unsigned)xx < (unsigned)p.W;
        if (p.upsample) xx >>= 1;
        cp_async16(rowdst + (uint32_t)(c * g.npx + col) * 16u, valid ? rowsrc + (xx * p.Cin + c * 8) : Abase, valid);
      }
    }
  };
  if ((int)blockIdx.x < total_units) issue_halo(blockIdx.x);

  for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++ucount) {
    const int n_tile = unit % g.n_tiles;
    const int band = unit / g.n_tiles;
    const int b = band / g.bands_per_img;
    const int y0 = (band - b * g.bands_per_img) * g.R;

    // ---------------- phase 1: the halo of this band was issued ahead (before the previous epilogue); land it ----------------
    if (g.tma) {
      mbar_wait(smem_u32(&halo_bar), ucount & 1);
      if (p.pad_mode == 1) {  // reflect columns: x = -1 <- x = 1, x = W <- x = W - 2 (16-byte pixels of every row and plane)
        const int n = (g.R + 2) * cpp * 2;
        for (int i = threadIdx.x; i < n; i += CB_THREADS) {
          const int side = i & 1, c = (i >> 1) % cpp, hr = (i >> 1) / cpp;
          const uint32_t rowb = halo_base + (uint32_t)(c * g.npx + hr * g.Wp + g.lead) * 16u;
          const uint32_t src = rowb + (uint32_t)(side ? p.W - 1 : 2) * 16u, dst = rowb + (uint32_t)(side ? p.W + 1 : 0) * 16u;
          uint32_t v0, v1, v2, v3;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(src) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
        }
      }
    } else {
      cp_async_wait_all();
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
    __syncthreads();
#ifdef MST_BAND_PROF
    { long long n = clock64(); t_fill += n - t_mark; t_mark = n; }
#endif

    // ---------------- phase 2: stream weight k-blocks, accumulate every strip of the band in TMEM ----------------
    if (warp == 0) {
      if (lane == 0) {
        const uint8_t* wtile = Wbase + (size_t)n_tile * nkb * B_STAGE_BYTES;
        int it = wit;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % g.wst;
          if (it >= g.wst) mbar_wait(smem_u32(&empty_bar[s]), ((it / g.wst) - 1) & 1);
          cb_arrive_expect_tx(smem_u32(&full_bar[s]), B_STAGE_BYTES);
          cb_bulk_g2s(ring_base + s * B_STAGE_BYTES, wtile + (size_t)kb * B_STAGE_BYTES, B_STAGE_BYTES, smem_u32(&full_bar[s]));
        }
      }
    } else if (warp == 1) {
      int it = wit;
      const uint32_t lbo = (uint32_t)g.npx * 16u;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % g.wst;
        mbar_wait(smem_u32(&full_bar[s]), (it / g.wst) & 1);
        tc_fence_after();
        {  // whole warp, warp-uniform values, one lane elected inside the asm (uniform-register issue loop)
          const uint32_t b_stage = ring_base + s * B_STAGE_BYTES;
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int ks = kb * 4 + j;
            if (ks >= total_ks) break;
            const int tap = ks / kpt;
            const int chunk0 = (ks - tap * kpt) * 2;  // first of the two 16-byte channel chunks of this K=16 step
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint64_t bd = umma_desc_sw128(b_stage + j * 32);
            const uint32_t a0 = halo_base + (uint32_t)(chunk0 * g.npx + ky * g.Wp + kx + g.lead - 1) * 16u;
            for (int st = 0; st < g.NS; ++st)
              umma_bf16_pred(tmem_base + st * BN, umma_desc_none(a0 + (uint32_t)st * 2048u, lbo, 128u), bd, idesc, ks != 0);
          }
          umma_commit_pred(smem_u32(&empty_bar[s]));
          if (kb == nkb - 1) umma_commit_pred(smem_u32(&accum_bar));
        }
      }
    }
    wit += nkb;

    // ---------------- phase 3: epilogue (all 8 warps) ----------------
    if (lane == 0) mbar_wait(smem_u32(&accum_bar), ucount & 1);
    __syncwarp();
    tc_fence_after();
#ifdef MST_BAND_PROF
    { long long n = clock64(); t_mma += n - t_mark; t_mark = n; }
#endif
    // every MMA of this band has retired, so the halo buffer is free: start fetching the next band's halo now and
    // let it land while the epilogue drains TMEM
    __syncthreads();
    if (unit + (int)gridDim.x < total_units) issue_halo(unit + gridDim.x);
    {
      constexpr int CH = 16;
      constexpr int NCC = BN / CH;
      const int quad = warp & 3, half = warp >> 2;
      const int nbase = n_tile * BN;
      const long long hw = (long long)p.H * p.W;
#pragma unroll 1
      for (int item = half; item < g.NS * NCC; item += 2) {
        const int st = item / NCC, cc = item - st * NCC;
        uint32_t v[CH];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + st * BN + cc * CH;
        tmem_ld16(taddr, v);
        tmem_wait_ld();
        const int q = st * 128 + quad * 32 + lane;
        const int r = q / g.Wp;
        const int pxl = q - r * g.Wp;
        const int y = y0 + r, x = pxl - 1;
        if (r >= g.R || y >= p.H || x < 0 || x >= p.W) continue;
        const int n = nbase + cc * CH;
        float xv[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          xv[j] = __uint_as_float(v[j]) + bias_s[n + j];
          if (p.act == MST_ACT_RELU) xv[j] = fmaxf(xv[j], 0.0f);
          else if (p.act == MST_ACT_GELU) xv[j] = gelu_erf(xv[j]);
        }
        const long long pix = (long long)b * hw + (long long)y * p.W + x;
        if (p.out_nchw) {
          const long long base = (long long)b * p.n_real * hw + (long long)y * p.W + x;
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (n + j < p.n_real) p.out_f32[base + (long long)(n + j) * hw] = xv[j];
        } else {
          if (p.out_f32) {
            float4* o4 = reinterpret_cast<float4*>(p.out_f32 + pix * p.ld_out32 + n);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) o4[j] = make_float4(xv[4 * j], xv[4 * j + 1], xv[4 * j + 2], xv[4 * j + 3]);
          }
          if (p.out_bf16 && CH >= 16 && ((reinterpret_cast<uintptr_t>(p.out_bf16) | (uintptr_t)(p.ld_out16 * 2)) & 31) == 0) {
            bf16* op = reinterpret_cast<bf16*>(p.out_bf16) + pix * p.ld_out16 + n;  // one 32-byte sector per lane and store
#pragma unroll
            for (int j = 0; j < CH / 16; ++j) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(xv[16 * j + 2 * e], xv[16 * j + 2 * e + 1]);
                pk[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              st_global_256(op + 16 * j, pk);
            }
          } else if (p.out_bf16) {
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out_bf16) + pix * p.ld_out16 + n);
#pragma unroll
            for (int j = 0; j < CH / 8; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(xv[8 * j + 2 * e], xv[8 * j + 2 * e + 1]);
                pk[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              o4[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM and the halo are free for the next band
    tc_fence_after();
#ifdef MST_BAND_PROF
    { long long n = clock64(); t_epi += n - t_mark; t_mark = n; }
#endif
  }
#ifdef MST_BAND_PROF
  if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 64)) printf("band prof thread %d units %d: fill %lld mma %lld epi %lld cycles (R=%d NS=%d npx=%d)\n", threadIdx.x, ucount, t_fill, t_mma, t_epi, g.R, g.NS, g.npx);
#endif
  if (warp == 1) tmem_dealloc(tmem_base, g.tmem_cols);
}

static int cb_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// Choose rows per band: the halo must fit beside the weight ring, all strips must fit in TMEM; maximise
// (useful MMA rows / issued MMA rows) * (halo reuse R/(R+2)).
static bool band_tma_ok(const MstGemm& g) {  // halo by tensor copies: dense NHWC input (no folded upsample), box width <= 256
  return !g.upsample && g.Cin % 8 == 0 && g.W + 2 <= 248 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && getenv("MST_BAND_TMA") == nullptr;
}

static bool plan_band_with(const MstGemm& g, int BN, long long smem_total, int tmem_cols, BandGeom& out, double& best) {
  const bool tma = band_tma_ok(g);
  // tensor copies need 128-byte aligned destinations: row pitch and lead offset in multiples of 8 pixels (16 B each)
  const int Wp = tma ? (g.W + 2 + 7) / 8 * 8 : g.W + 2;
  const int lead = tma ? 8 : 1;
  const int wst = BN >= 256 ? 2 : 4;
  const long long budget = smem_total - 1024 /*align*/ - (long long)wst * BN * 128;
  best = 0.0;
  bool ok = false;
  for (int R = 1; R <= g.H && R <= 32; ++R) {
    const int NS = (R * Wp + 127) / 128;
    if (NS * BN > tmem_cols) break;
    const int npx = (NS * 128 + 2 * Wp + 8 + lead + 7) / 8 * 8;
    const long long halo = (long long)g.Cin * 2 * npx;
    if (halo > budget) break;
    if (npx > 0x3FFF) break;  // LBO field
    const int bands = (g.H + R - 1) / R;
    const double util = (double)g.H * g.W / ((double)bands * NS * 128.0);
    const double score = util * R / (R + 2.0);
    if (score > best) {
      best = score;
      ok = true;
      out.R = R; out.NS = NS; out.npx = npx; out.Wp = Wp; out.bands_per_img = bands; out.n_tiles = g.N / BN; out.wst = wst;
      out.tmem_cols = tmem_cols;
      out.lead = lead; out.tma = tma ? 1 : 0;
    }
  }
  return ok;
}

// Prefer a plan that lets two CTAs share an SM (half the shared memory and TMEM each): their load / MMA / epilogue
// phases then overlap.  Fall back to one big CTA per SM when that would cost too much halo re-reading.
static bool plan_band(const MstGemm& g, int BN, BandGeom& out) {
  BandGeom two, one;
  double s2 = 0.0, s1 = 0.0;
  const bool ok2 = plan_band_with(g, BN, 110 * 1024, 256, two, s2);
  const bool ok1 = plan_band_with(g, BN, 220 * 1024, 512, one, s1);
  // measured (tools/gemm_bench.py band): the extra halo re-reads of the small plan cost more than the overlap gains
  if (ok1) { out = one; return true; }
  if (ok2) { out = two; return true; }
  return false;
}

template <int BN>
static int launch_band(const MstGemm& g, cudaStream_t st) {
  BandGeom geo;
  if (!plan_band(g, BN, geo)) return MST_ERR_UNSUPPORTED;
  const size_t smem = 1024 + (size_t)geo.wst * BN * 128 + (size_t)g.Cin * 2 * geo.npx;
  static size_t attr_set = 0;
  if (smem > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_band_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = 227 * 1024;
  }
  const int B = g.M / (g.H * g.W);
  const long long units = (long long)B * geo.bands_per_img * geo.n_tiles;
  const unsigned grid = (unsigned)(units < cb_num_sms() ? units : cb_num_sms());
  GemmCore core;
  memcpy(&core, &g, sizeof(GemmCore));
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (geo.tma) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
      void* fp = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        enc = reinterpret_cast<EncodeTiledFn>(fp);
    }
    const cuuint64_t gdim[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)B};
    const cuuint64_t gstride[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.W * g.Cin * 2, (cuuint64_t)g.H * g.W * g.Cin * 2};
    const cuuint32_t box[4] = {8, (cuuint32_t)geo.Wp, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!enc || enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(reinterpret_cast<const void*>(g.A)), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MST_ERR_UNSUPPORTED;
  }
  conv_band_kernel<BN><<<grid, CB_THREADS, smem, st>>>(core, geo, (int)units, tmap);
  return (int)cudaGetLastError();
}


// =====================================================================================================
// Row-streaming variant for the thin layers (Cin <= 64, one n-tile, W a multiple of 128).
//
// The band kernel above runs load -> MMA -> epilogue one after the other per band; for the thin layers of the
// CNN decoder (64->64, 64->32 at 128^2; 32->32, 32->3 at 256^2) and VGG conv1_2 that serial chain, not HBM or
// the tensor pipe, set the time (ncu: tensor pipe 9 %, dram 6 %).  Here the image is a STREAM of padded input
// rows through a shared-memory ring and the three phases run concurrently, warp-specialised:
//
//   warps 0-7   epilogue   : TMEM -> bias / activation -> global, one finished output row at a time
//   warps 8-14  producers  : one padded input row per ring slot with cp.async (reflect / zero padding and the
//                            nearest-x2 upsample folded into the source address), asynchronous mbarrier arrival
//   warp  15    MMA issuer : output row y = 9 taps x Cin/16 k-steps x W/128 strips of tcgen05.mma reading ring
//                            rows y-1, y, y+1 through ROW-SHIFTED swizzled descriptors; accumulators rotate
//                            through NACC TMEM buffers
//
// The whole [N x 9 Cin] weight matrix stays resident in shared memory (fetched once per CTA).  Each CTA owns a
// contiguous run of output rows (over all images), so every input row is fetched from L2/HBM once per CTA that
// needs it (3 extra rows per run) and each strip is exactly one 128-pixel half/whole row: no padded-raster waste.
// Ring layout: a slot is one padded row, pixel-major, one pixel = one K-major operand row of Cin * 2 bytes (64 B ->
// SWIZZLE_64B, 128 B -> SWIZZLE_128B), the 16-byte chunks swizzled by the address bits above them; slot pitch = a whole
// number of 8-row swizzle atoms.  Tap (ky, kx) of a strip is the SAME bytes read through a descriptor whose start is
// moved by kx rows: the swizzle is a function of the absolute shared-memory address (tools/micro/umma_rowshift.cu
// checks every shift 0..9 on the device), so a start that is not atom-aligned is fine, base-offset field 0.  The
// earlier SWIZZLE_NONE layout (planes of 16-byte chunks, pixel shift = 16 B) made two of three taps read core
// matrices that straddle 128-byte lines: ~58 clk per MMA whatever N.
constexpr int RS_EPI_WARPS = 8;
constexpr int RS_PROD_WARPS = 7;
constexpr int RS_MMA_WARPS = 1;
constexpr int RS_THREADS = (RS_EPI_WARPS + RS_PROD_WARPS + RS_MMA_WARPS) * 32;
constexpr int RS_MAX_RING = 16;

struct RowsGeom {
  int ring;         // ring slots (padded input rows resident)
  int spr;          // 128-pixel strips per output row (W / 128)
  int nacc;         // TMEM accumulator buffers
  int tmem_cols;    // allocated TMEM columns (power of two >= nacc * spr * BN)
  int rows_total;   // B * H output rows
  int rows_per_cta;
  int row_pitch;    // bytes per ring slot: (W + 2) pixels rounded up to whole 8-pixel swizzle atoms
  int mode;         // experiment switch (MST_ROWS_MODE)
};

// G > 1: BANDED weights.  One MMA group produces G consecutive output rows of a strip: N = G * BN accumulator columns
// (column dy * BN + co = output row y + dy, channel co), K runs over the G + 2 input rows the group touches
// (k = ((j * 3 + kx) * CIN + c), j = input row y - 1 + j), and the B operand is the weight matrix laid out as a band:
// B[dy * BN + co][j, kx, c] = W[co][ky = j - dy][kx][c] for 0 <= j - dy <= 2, zero elsewhere.  Two thirds (G = 2) or half
// (G = 4) of the MACs multiply zeros -- but the thin layers are bound by the shared-memory read of the PIXEL operand (4 KB per
// MMA whatever N), and a banded MMA reads each staged pixel row once for all G output rows: (G + 2) * 3 * CIN / 16 MMAs per G
// rows instead of 9 * CIN / 16 per row.  The band is built once per CTA in shared memory from the ordinary packed weights.
template <int BN, int CIN, int G>
__global__ void __launch_bounds__(RS_THREADS, 1) conv_rows_kernel(const GemmCore p, const RowsGeom g) {
  constexpr int CPP = CIN / 8;    // 16-byte chunks per pixel
  constexpr int KPT = CIN / 16;   // K=16 steps per tap
  constexpr int NB = G * BN;      // MMA N: G output rows x BN channels
  constexpr int JR = G + 2;       // input rows per group
  constexpr int TOTAL_KS = JR * 3 * KPT;
  constexpr int NKB = (JR * 3 * CIN + 63) / 64;
  constexpr int NKB_SRC = (9 * CIN + 63) / 64;  // k-blocks of the packed (un-banded) weights
  constexpr int B_STAGE_BYTES = NB * 128;
  constexpr int CH = BN >= 32 ? 32 : 16;  // columns per tcgen05.ld
  constexpr int NCC = BN / CH;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[RS_MAX_RING];
  __shared__ uint64_t free_bar[RS_MAX_RING];
  __shared__ uint64_t acc_full[4];
  __shared__ uint64_t acc_empty[4];
  __shared__ uint64_t w_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[BN];

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform for ptxas
  const uint32_t w_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring_base = w_base + NKB * B_STAGE_BYTES;
  const int Wp = p.W + 2;
  const int r_begin = blockIdx.x * g.rows_per_cta;
  const int r_end = min(r_begin + g.rows_per_cta, g.rows_total);

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.ring; ++s) {
      mbar_init(smem_u32(&full_bar[s]), RS_PROD_WARPS * 32);
      mbar_init(smem_u32(&free_bar[s]), RS_MMA_WARPS);
    }
    for (int b = 0; b < g.nacc; ++b) {
      mbar_init(smem_u32(&acc_full[b]), RS_MMA_WARPS);
      mbar_init(smem_u32(&acc_empty[b]), RS_EPI_WARPS);
    }
    mbar_init(smem_u32(&w_bar), G == 1 ? 1 : RS_PROD_WARPS * 32);
    mbar_fence_init();
  }
  if (threadIdx.x < BN) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (warp == RS_EPI_WARPS + RS_PROD_WARPS) {
    tmem_alloc(smem_u32(&tmem_base_slot), g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int acc_cols = g.spr * NB;

  if (warp >= RS_EPI_WARPS && warp < RS_EPI_WARPS + RS_PROD_WARPS) {
    // =========================== producers ===========================
    const int t = threadIdx.x - RS_EPI_WARPS * 32;
    if constexpr (G == 1) {
      if (t == 0) {  // resident weights, once
        cb_arrive_expect_tx(smem_u32(&w_bar), NKB * B_STAGE_BYTES);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.Wt);
        for (int kb = 0; kb < NKB; ++kb)
          cb_bulk_g2s(w_base + kb * B_STAGE_BYTES, wsrc + (size_t)kb * B_STAGE_BYTES, B_STAGE_BYTES, smem_u32(&w_bar));
      }
    } else {
      // the band, 16 bytes (8 input channels) at a time: destination row n = dy * BN + co, K chunk (j, kx, c8)
      uint8_t* gen = smem_raw + (w_base - smem_u32(smem_raw));
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.Wt);
      constexpr int KCH = NKB * 8;  // 16-byte K chunks per destination row (incl. the zero tail of the last k-block)
      for (int i = t; i < NB * KCH; i += RS_PROD_WARPS * 32) {
        const int n = i / KCH, kc = i - n * KCH;
        const int dy = n / BN, co = n - dy * BN;
        const int kel = kc * 8;                      // first K element of the chunk
        const int jk = kel / CIN, c = kel - jk * CIN;  // jk = j * 3 + kx
        const int j = jk / 3, kx = jk - j * 3;
        const int ky = j - dy;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (jk < JR * 3 && ky >= 0 && ky <= 2) {
          const int ksrc = (ky * 3 + kx) * CIN + c;
          v = *reinterpret_cast<const uint4*>(wsrc + (size_t)(ksrc >> 6) * (BN * 128) + sw128_offset(co, (ksrc & 63) >> 3));
        }
        *reinterpret_cast<uint4*>(gen + (size_t)(kc >> 3) * B_STAGE_BYTES + sw128_offset(n, kc & 7)) = v;
      }
      fence_proxy_async_smem();  // generic-proxy stores, read by the tensor core through the async proxy
      mbar_arrive(smem_u32(&w_bar));
    }
    const bf16* Abase = reinterpret_cast<const bf16*>(p.A);
    const int Hs = p.upsample ? (p.H >> 1) : p.H;
    const int Ws = p.upsample ? (p.W >> 1) : p.W;
    const int row_chunks = Wp * CPP;
    // this thread's copies are the same for every row: chunk idx = t + k * NPROD -> (padded column, channel chunk);
    // source element offset inside the source row (-1 = zero fill) and destination byte offset inside the ring row
    constexpr int NPROD = RS_PROD_WARPS * 32;
    constexpr int MAXK = (514 * CPP + NPROD - 1) / NPROD;  // W <= 512
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
      if (p.upsample) xx >>= 1;
      src_off[k] = valid ? xx * CIN + c * 8 : -1;
      // pixel-major, chunk index XOR the address bits 7.. of the pixel row (CIN 64: col & 7; CIN 32: (col >> 1) & 3)
      const int sw = CIN == 64 ? (col & 7) : ((col >> 1) & 3);
      dst_off[k] = (uint32_t)col * (uint32_t)(CIN * 2) + (uint32_t)((c ^ sw) << 4);
    }
    const int nk = (row_chunks - t + NPROD - 1) / NPROD;  // copies this thread makes per row
    int slot = 0;
    const bool prof = (g.mode & 8) != 0;
    long long t_wait = 0, t_all = prof ? clock64() : 0;
    uint32_t fphase = 1;  // a fresh free_bar passes a wait on parity 1
    int b = r_begin / p.H;
    int y = r_begin - b * p.H;
    for (int r = r_begin; r < r_end; r += G) {
      const int nload = (r == r_begin || y == 0) ? JR : G;
      for (int j = JR - nload; j < JR; ++j) {
        int yy = y + j - 1;  // padded row -1 .. H
        bool vrow = true;
        if (p.pad_mode == 1) yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);
        else vrow = (unsigned)yy < (unsigned)p.H;
        if (p.upsample) yy >>= 1;
        const bf16* rowsrc = Abase + ((long long)b * Hs + (vrow ? yy : 0)) * Ws * CIN;
        long long tp0 = prof ? clock64() : 0;
        mbar_wait(smem_u32(&free_bar[slot]), fphase);
        if (prof) t_wait += clock64() - tp0;
        const uint32_t rowdst = ring_base + (uint32_t)slot * (uint32_t)g.row_pitch;
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
          if (k < nk) {
            const bool valid = vrow && src_off[k] >= 0;
            cp_async16(rowdst + dst_off[k], valid ? rowsrc + src_off[k] : Abase, valid);
          }
        }
        cp_async_mbar_arrive_noinc(smem_u32(&full_bar[slot]));
        if (++slot == g.ring) { slot = 0; fphase ^= 1; }
      }
      y += G;
      if (y == p.H) { y = 0; ++b; }
    }
    cp_async_wait_all();
    if (prof && blockIdx.x == 1 && t == 0) printf("rows prof producer: total %lld wait_free %lld rows %d\n", clock64() - t_all, t_wait, r_end - r_begin);
  } else if (warp == RS_EPI_WARPS + RS_PROD_WARPS) {
    // =========================== MMA issuer ===========================
    // The whole warp runs this loop convergently on warp-uniform values (kernel parameters, blockIdx, loop counters);
    // umma_bf16_pred / umma_commit_pred elect one lane inside the asm.  ptxas then keeps the descriptors in uniform
    // registers and emits back-to-back UTCHMMA with one uniform add in between -- issuing from inside an
    // `if (lane == 0)` branch instead costs ~15 instructions (register -> uniform-register broadcasts in an elect
    // loop) per MMA, more than these small (N <= 64) MMAs take to execute.
    constexpr uint32_t idesc = umma_idesc_bf16(128, NB);
    constexpr uint32_t PIX16 = CIN * 2 / 16;                           // one pixel (operand row) in address-field units
    constexpr uint32_t a_hi = ((8u * CIN * 2u) >> 4) | (1u << 14) | ((CIN == 64 ? 2u : 4u) << 29);  // SBO = 8 pixels, version 1, SWIZZLE_128B / _64B
    constexpr uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
    const uint32_t b_lo0 = ((w_base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t row16 = (uint32_t)g.row_pitch >> 4;                 // one ring slot in address-field units (16 B)
    const uint32_t a_lo_base = ((ring_base & 0x3FFFFu) >> 4) | (1u << 16);
    mbar_wait(smem_u32(&w_bar), 0);
    const bool prof = (g.mode & 8) != 0;
    long long t_full = 0, t_acc = 0, t_all = prof ? clock64() : 0;
    int rcount = 0;                 // row groups issued
    int y = r_begin % p.H;
    int s_new = 0;                  // ring slot of the next padded row to arrive
    uint32_t full_phase = 0;
    int sl[JR];                     // ring slots of padded rows y-1 .. y+G
#pragma unroll
    for (int j = 0; j < JR; ++j) sl[j] = 0;
    int buf = 0;
    uint32_t acc_phase = 1;         // a fresh acc_empty passes a wait on parity 1
    for (int r = r_begin; r < r_end; r += G, ++rcount) {
      const int nload = (r == r_begin || y == 0) ? JR : G;
      long long tm0 = prof ? clock64() : 0;
      for (int j = 0; j < nload; ++j) {
        mbar_wait(smem_u32(&full_bar[s_new]), full_phase);
#pragma unroll
        for (int q = 0; q + 1 < JR; ++q) sl[q] = sl[q + 1];
        sl[JR - 1] = s_new;
        if (++s_new == g.ring) { s_new = 0; full_phase ^= 1; }
      }
      long long tm1 = prof ? clock64() : 0;
      mbar_wait(smem_u32(&acc_empty[buf]), acc_phase);
      if (prof) { t_full += tm1 - tm0; t_acc += clock64() - tm1; }
      fence_proxy_async_smem();
      tc_fence_after();
      y += G;
      if (y == p.H) y = 0;
      const bool last_of_run = (r + G >= r_end || y == 0);
      uint32_t a_row[JR];
#pragma unroll
      for (int j = 0; j < JR; ++j) a_row[j] = a_lo_base + (uint32_t)sl[j] * row16;
      for (int st = 0; st < g.spr; ++st) {
        const uint32_t d_tmem = tmem_base + buf * acc_cols + st * NB;
#pragma unroll
        for (int ks = 0; ks < TOTAL_KS; ++ks) {
          const int tap = ks / KPT, kk = ks - tap * KPT;
          const int j = tap / 3, kx = tap - j * 3;
          // a pixel further = one operand row (PIX16 address units); a K=16 step = 32 bytes inside the swizzled row
          const uint32_t a_lo = a_row[j] + (uint32_t)(st * 128 + kx) * PIX16 + (uint32_t)(kk * 2);
          const uint32_t b_lo = b_lo0 + (uint32_t)(((ks >> 2) * B_STAGE_BYTES + (ks & 3) * 32) >> 4);
          umma_bf16_pred(d_tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, ks != 0);
        }
      }
      umma_commit_pred(smem_u32(&acc_full[buf]));
      // padded rows y-1 .. y+G-2 are not needed again; at the end of an image / of this CTA's run the last two die as well
#pragma unroll
      for (int j = 0; j < G; ++j) umma_commit_pred(smem_u32(&free_bar[sl[j]]));
      if (last_of_run) {
        umma_commit_pred(smem_u32(&free_bar[sl[G]]));
        umma_commit_pred(smem_u32(&free_bar[sl[G + 1]]));
      }
      if (++buf == g.nacc) { buf = 0; acc_phase ^= 1; }
    }
    if (prof && blockIdx.x == 1 && lane == 0) printf("rows prof mma: total %lld wait_full %lld wait_acc_empty %lld\n", clock64() - t_all, t_full, t_acc);
    tc_fence_before();
  } else {
    // =========================== epilogue (warps 0-7) ===========================
    const int quad = warp & 3, half = warp >> 2;
    const bool prof = (g.mode & 8) != 0;
    long long t_wait = 0, t_all = prof ? clock64() : 0;
    const int items = g.spr * G * NCC;  // (strip, output row of the group, column chunk)
    float bias_r[NCC == 1 ? CH : 1];    // one column chunk per row (BN <= 32): the bias lives in registers, not 32 LDS per item
    if constexpr (NCC == 1) {
#pragma unroll
      for (int j = 0; j < CH; ++j) bias_r[j] = bias_s[j];
    }
    const long long hw = (long long)p.H * p.W;
    const bool wide16 = p.out_bf16 && ((reinterpret_cast<uintptr_t>(p.out_bf16) | (uintptr_t)(p.ld_out16 * 2)) & 31) == 0;
    int rcount = 0;
    for (int r0 = r_begin; r0 < r_end; r0 += G, ++rcount) {
      const int buf = rcount % g.nacc;
      long long te0 = prof ? clock64() : 0;
      if (lane == 0) mbar_wait(smem_u32(&acc_full[buf]), ((uint32_t)(rcount / g.nacc)) & 1u);
      __syncwarp();
      if (prof) t_wait += clock64() - te0;
      tc_fence_after();
      const int last_item = half + ((items - 1 - half) / 2) * 2;  // this warp's last item (< half: none)
      if (half >= items) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
        continue;
      }
#pragma unroll 1
      for (int item = half; item < items; item += 2) {
        const int st = item / (G * NCC), rem = item - st * (G * NCC);
        const int dy = rem / NCC, cc = rem - dy * NCC;
        const int r = r0 + dy;  // output row of this item
        uint32_t v[CH];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * acc_cols + st * NB + dy * BN + cc * CH;
        if constexpr (CH == 32) tmem_ld32(taddr, v);
        else tmem_ld16(taddr, v);
        tmem_wait_ld();
        if (item == last_item) {  // accumulator drained by this warp: hand it back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
        }
        const int x = st * 128 + quad * 32 + lane;
        const int n = cc * CH;
        float xv[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          xv[j] = __uint_as_float(v[j]) + (NCC == 1 ? bias_r[j] : bias_s[n + j]);
          if (p.act == MST_ACT_RELU) xv[j] = fmaxf(xv[j], 0.0f);
          else if (p.act == MST_ACT_GELU) xv[j] = gelu_erf(xv[j]);
        }
        const long long pix = (long long)r * p.W + x;
        if (p.out_nchw == MST_OUT_IMAGE_U8) {  // uint8 [B,H,W,n_real] image: np.clip(x * 255, 0, 255).astype(np.uint8) (NaN -> 0)
          uint8_t* o8 = reinterpret_cast<uint8_t*>(p.out_f32) + pix * p.n_real;
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (n + j < p.n_real) o8[n + j] = (uint8_t)(uint32_t)fminf(fmaxf(__fmul_rn(xv[j], 255.0f), 0.0f), 255.0f);
        } else if (p.out_nchw) {
          const int b = r / p.H;
          const long long base = (long long)b * p.n_real * hw + (pix - (long long)b * hw);
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (n + j < p.n_real) p.out_f32[base + (long long)(n + j) * hw] = xv[j];
        } else {
          if (p.out_f32) {
            float4* o4 = reinterpret_cast<float4*>(p.out_f32 + pix * p.ld_out32 + n);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) o4[j] = make_float4(xv[4 * j], xv[4 * j + 1], xv[4 * j + 2], xv[4 * j + 3]);
          }
          if (p.out_bf16) {
            bf16* op = reinterpret_cast<bf16*>(p.out_bf16) + pix * p.ld_out16 + n;
            if (wide16) {
#pragma unroll
              for (int j = 0; j < CH / 16; ++j) {
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(xv[16 * j + 2 * e], xv[16 * j + 2 * e + 1]);
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
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(xv[8 * j + 2 * e], xv[8 * j + 2 * e + 1]);
                  pk[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
                o4[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
      }
    }
    if (prof && blockIdx.x == 1 && lane == 0 && (warp == 0 || warp == 7)) printf("rows prof epilogue warp %d: total %lld wait_acc_full %lld\n", warp, clock64() - t_all, t_wait);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == RS_EPI_WARPS + RS_PROD_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
}

// Output rows per MMA group: as many as keep N = G * BN at 64 (the pixel operand's shared-memory read then serves G rows),
// if the image height divides and the band + a ring of G + 3 rows fit.  MST_ROWS_BAND=0 turns banding off (experiments).
static int rows_band(const MstGemm& g, int BN) {
  static int allow = -1;
  if (allow < 0) { const char* e = getenv("MST_ROWS_BAND"); allow = e ? atoi(e) : 1; }
  if (!allow) return 1;
  // measured (tools/rows_prof.py, bench detail): 32 -> 3(16) @256^2: G = 4 starves the ring (a 74 KB band leaves 8 slots for
  // 6 live rows), G = 2 does not; 32 -> 32: G = 2 a little faster (76 -> 70 us); 64 -> 32 @128^2: slower (98 KB band, ring of 7)
  int G = 1;
  if (g.Cin == 32 && (BN == 32 || BN == 16)) G = 2;
  if (G > 1 && g.H % G != 0) G = 1;
  return G;
}

static bool plan_rows(const MstGemm& g, int BN, RowsGeom& out) {
  if (g.W % 128 != 0 || g.W > 512 || g.N != BN || BN > 64) return false;
  if (g.Cin != 32 && g.Cin != 64) return false;
  const int nkb_src = (9 * g.Cin + 63) / 64;
  if (g.k_pad != nkb_src * 64) return false;
  const int G = rows_band(g, BN);
  const int nkb = ((G + 2) * 3 * g.Cin + 63) / 64;
  const long long budget = 220LL * 1024 - 1024 - (long long)nkb * G * BN * 128;
  const long long row_bytes = (long long)((g.W + 2 + 7) / 8 * 8) * g.Cin * 2;
  long long ring = budget / row_bytes;
  if (ring > 12) ring = 12;
  if (ring < G + 3) return false;
  out.ring = (int)ring;
  out.spr = g.W / 128;
  const int acc_cols = out.spr * G * BN;
  if (acc_cols > 256) return false;
  int nacc = 256 / acc_cols;
  if (nacc > 4) nacc = 4;
  if (nacc < 2) nacc = 2;
  out.nacc = nacc;
  int cols = 32;
  while (cols < nacc * acc_cols) cols <<= 1;
  out.tmem_cols = cols;
  out.row_pitch = (int)row_bytes;
  { const char* e = getenv("MST_ROWS_MODE"); out.mode = e ? atoi(e) : 0; }
  const int B = g.M / (g.H * g.W);
  out.rows_total = B * g.H;
  const int sms = cb_num_sms();
  out.rows_per_cta = (out.rows_total + sms - 1) / sms;
  const int min_rows = out.rows_total < 4 ? out.rows_total : 4;  // a run re-fetches 2 extra input rows: keep runs >= 4 rows
  if (out.rows_per_cta < min_rows) out.rows_per_cta = min_rows;
  out.rows_per_cta = (out.rows_per_cta + G - 1) / G * G;  // runs are whole row groups (H % G == 0, so a group never straddles two images)
  return true;
}

template <int BN, int CIN, int G>
static int launch_rows(const MstGemm& g, const RowsGeom& geo, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)(((G + 2) * 3 * CIN + 63) / 64) * G * BN * 128 + (size_t)geo.ring * geo.row_pitch;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_rows_kernel<BN, CIN, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const unsigned grid = (unsigned)((geo.rows_total + geo.rows_per_cta - 1) / geo.rows_per_cta);
  GemmCore core;
  memcpy(&core, &g, sizeof(GemmCore));
  conv_rows_kernel<BN, CIN, G><<<grid, RS_THREADS, smem, st>>>(core, geo);
  return (int)cudaGetLastError();
}

// Row-streaming kernel when the shape allows it, else the band kernel.
static int try_launch_rows(const MstGemm& g, cudaStream_t st, bool& handled) {
  RowsGeom geo;
  handled = true;
  const int bn = mst_gemm_tile_n(g.N);
  if (!plan_rows(g, bn, geo)) { handled = false; return 0; }
  const int G = rows_band(g, bn);
  if (bn == 64 && g.Cin == 64) return launch_rows<64, 64, 1>(g, geo, st);
  if (bn == 32 && g.Cin == 64) return launch_rows<32, 64, 1>(g, geo, st);
  if (bn == 32 && g.Cin == 32) return G == 2 ? launch_rows<32, 32, 2>(g, geo, st) : launch_rows<32, 32, 1>(g, geo, st);
  if (bn == 16 && g.Cin == 32) return G == 2 ? launch_rows<16, 32, 2>(g, geo, st) : launch_rows<16, 32, 1>(g, geo, st);
  handled = false;
  return 0;
}

}  // namespace mst

using namespace mst;

extern "C" int mst_conv3x3_band_supported(int N, int Cin, int H, int W) {
  if (N <= 0 || N % 16 || N > CB_MAX_BIAS || Cin % 16 || Cin <= 0 || H < 2 || W < 2) return 0;
  MstGemm g{};
  g.N = N; g.Cin = Cin; g.H = H; g.W = W;
  BandGeom geo;
  const int bn = mst_gemm_tile_n(N);
  return plan_band(g, bn, geo) ? 1 : 0;
}

extern "C" int mst_conv3x3_band(const MstGemm* g, void* stream) {
  if (!g || !g->A || !g->Wt) return MST_ERR_BAD_ARG;
  if (g->a_mode != MST_A_CONV3X3 || g->M <= 0 || g->N <= 0 || g->K != 9 * g->Cin || g->k_pad % 64 || g->k_pad < g->K) return MST_ERR_BAD_ARG;
  if (g->N % 16 || g->N > CB_MAX_BIAS || g->Cin % 16) return MST_ERR_UNSUPPORTED;
  if (g->H < 2 || g->W < 2 || g->M % (g->H * g->W) != 0) return MST_ERR_BAD_ARG;
  if (g->upsample && ((g->H | g->W) & 1)) return MST_ERR_BAD_ARG;
  if (g->res || g->mul || g->out_nchw == MST_OUT_IMAGE_U8) return MST_ERR_UNSUPPORTED;
  if (!g->out_f32 && !g->out_bf16) return MST_ERR_BAD_ARG;
  if (g->out_nchw) {
    if (!g->out_f32 || g->n_real <= 0 || g->n_real > g->N) return MST_ERR_BAD_ARG;
  } else {
    if (g->out_f32 && g->ld_out32 % 4) return MST_ERR_BAD_ARG;
    if (g->out_bf16 && g->ld_out16 % 8) return MST_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (mst_gemm_tile_n(g->N)) {
    case 256: return launch_band<256>(*g, st);
    case 128: return launch_band<128>(*g, st);
    case 64: return launch_band<64>(*g, st);
    case 32: return launch_band<32>(*g, st);
    default: return launch_band<16>(*g, st);
  }
}

static bool rows_instantiated(int bn, int Cin) {
  return (bn == 64 && Cin == 64) || (bn == 32 && Cin == 64) || (bn == 32 && Cin == 32) || (bn == 16 && Cin == 32);
}

extern "C" int mst_conv3x3_rows_supported(int N, int Cin, int H, int W) {
  if (N <= 0 || N % 16 || Cin <= 0 || H < 2 || W < 2) return 0;
  MstGemm g{};
  g.N = N; g.Cin = Cin; g.H = H; g.W = W; g.M = H * W;
  g.k_pad = (9 * Cin + 63) / 64 * 64;
  RowsGeom geo;
  const int bn = mst_gemm_tile_n(N);
  return plan_rows(g, bn, geo) && rows_instantiated(bn, Cin) ? 1 : 0;
}

extern "C" int mst_conv3x3_rows(const MstGemm* g, void* stream) {
  if (!g || !g->A || !g->Wt) return MST_ERR_BAD_ARG;
  if (g->a_mode != MST_A_CONV3X3 || g->M <= 0 || g->N <= 0 || g->K != 9 * g->Cin || g->k_pad % 64 || g->k_pad < g->K) return MST_ERR_BAD_ARG;
  if (g->H < 2 || g->W < 2 || g->M % (g->H * g->W) != 0) return MST_ERR_BAD_ARG;
  if (g->upsample && ((g->H | g->W) & 1)) return MST_ERR_BAD_ARG;
  if (g->res || g->mul || g->gate || g->add16 || g->out_pre16 || g->row_scale || g->conv_full) return MST_ERR_UNSUPPORTED;
  if (!g->out_f32 && !g->out_bf16) return MST_ERR_BAD_ARG;
  if (g->out_nchw) {
    if (g->out_nchw < 0 || g->out_nchw > MST_OUT_IMAGE_U8 || !g->out_f32 || g->n_real <= 0 || g->n_real > g->N) return MST_ERR_BAD_ARG;
    if (g->out_nchw == MST_OUT_IMAGE_U8 && g->out_bf16) return MST_ERR_BAD_ARG;
  } else {
    if (g->out_f32 && g->ld_out32 % 4) return MST_ERR_BAD_ARG;
    if (g->out_bf16 && g->ld_out16 % 8) return MST_ERR_BAD_ARG;
  }
  bool handled = false;
  const int rc = try_launch_rows(*g, (cudaStream_t)stream, handled);
  return handled ? rc : MST_ERR_UNSUPPORTED;
}
