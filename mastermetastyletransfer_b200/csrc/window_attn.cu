// Fused shifted-window attention core (codes/style_transformer.py:83-111,127-168 and :544-607).
//
// One warp per (window, head), four warps per CTA, CTAs loop over windows.  Roll + partition live in the load addresses,
// window reverse + roll back in the store addresses.  K, V head slices ([64 x 32] bf16, 7x7 windows padded to
// 64 rows) are staged in shared memory, Q fragments are read straight from global; S = QK^T and O = PV run on the tensor cores (mma.sync m16n8k16,
// bf16 in, fp32 accumulate) 16 query rows at a time, with scale, relative-position bias, the 9-region shift
// mask and the softmax applied to the fp32 accumulator fragments in registers (quad shuffles for the row
// max / sum).  With v2/out2 the same probabilities multiply a second value tensor (sigma/mu attention).
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

constexpr int AT_WARPS = 4;
constexpr int AT_LD = 40;  // bf16 row stride of the staged tiles: 80 B keeps ldmatrix bank-conflict free

MST_DEVINL void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
MST_DEVINL void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
MST_DEVINL void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
MST_DEVINL float exp2f_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
MST_DEVINL uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int WS>
__global__ void __launch_bounds__(AT_WARPS * 32) window_attn_kernel(const MstWindowAttn a, const WinGeom g, const int total_cta_tasks) {
  constexpr int N = WS * WS;   // real tokens per window
  constexpr int NP = 64;       // rows of the staged tiles (N padded to a multiple of 16)
  constexpr int NT = (2 * WS - 1) * (2 * WS - 1);
  constexpr int NTV = (N + 7) / 8;  // 8-wide key tiles holding real keys (7 of 8 for 7x7 windows): the rest is never computed
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int heads = a.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dual = a.v2 != nullptr;
  const int tiles_per_warp = dual ? 3 : 2;  // K, V (, V2): Q fragments come straight from global, outputs go straight back
  bf16* tile0 = reinterpret_cast<bf16*>(smem_raw) + (size_t)warp * tiles_per_warp * NP * AT_LD;
  bf16* Ks = tile0;
  bf16* Vs = Ks + NP * AT_LD;
  bf16* V2s = Vs + NP * AT_LD;
  float* table_s = reinterpret_cast<float*>(reinterpret_cast<bf16*>(smem_raw) + (size_t)AT_WARPS * tiles_per_warp * NP * AT_LD);
  int* src_s = reinterpret_cast<int*>(table_s + NT * heads);
  int* lab_s = src_s + NP;

  // The softmax runs in base 2: scores, bias and mask are pre-multiplied by log2(e), so exp(s - max) is one FADD + one
  // MUFU.EX2 per element (mathematically identical to the reference's exp).
  constexpr float LOG2E = 1.4426950408889634f;
  // bias table transposed to [heads][NT] so a head's lookups are contiguous
  for (int i = threadIdx.x; i < NT * heads; i += blockDim.x) {
    const int idx = i / heads, hh = i - idx * heads;
    table_s[hh * NT + idx] = a.bias_table[i] * LOG2E;
  }
  const int gq = lane >> 2;        // fragment row within the 16-row tile (and +8)
  const int cq = (lane & 3) * 2;   // fragment column pair within an 8-column tile
  // this lane's 16 score columns j = nt*8 + cq + e: relative-position part yj*(2WS-1)+xj (window independent)
  int colpart[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = nt * 8 + cq + e;
      const int yj = j / WS, xj = j - yj * WS;
      colpart[nt][e] = yj * (2 * WS - 1) + xj;
    }

  // Each CTA walks over (window, group of four heads) tasks: the bias table and the column bookkeeping above are set up
  // once per CTA instead of once per window.
#pragma unroll 1
  for (int cta_task = blockIdx.x; cta_task < total_cta_tasks; cta_task += gridDim.x) {
  const int task0 = cta_task * AT_WARPS;  // the CTA's four tasks share one window (AT_WARPS divides heads)
  const int win_global = task0 / heads;
  const int h = task0 - win_global * heads + warp;
  const int b = win_global / g.nW;
  const int win = win_global - b * g.nW;
  __syncthreads();  // every warp is done with the previous window's src_s / lab_s (and the table is loaded)
  const int lab0 = win_label(g, win, 0);
  int differs = 0;
  for (int i = threadIdx.x; i < NP; i += blockDim.x) {
    int s = -2, l = 0;  // -2: row beyond the window (7x7 padded to 64 rows)
    if (i < N) {
      int y, x;
      win_source(g, win, i, y, x);
      s = (y < g.H && x < g.W) ? (b * g.H + y) * g.W + x : -1;  // -1: zero-padded token (takes the projection bias)
      l = win_label(g, win, i);
      differs |= (l != lab0);
    }
    src_s[i] = s;
    lab_s[i] = l;
  }
  // only windows that straddle a region boundary of the rolled grid have a non-trivial mask (19 of 100 at 7x7 / 64^2)
  const bool masked = __syncthreads_or(differs) != 0;

  const int c0 = h * 32;
  // ---- stage K, V (V2): 4 lanes x 16 B per token row, 8 rows per pass; real tokens go global -> shared with
  //      cp.async (all 16-24 copies of a lane in flight at once, no register staging) ----
  {
    const int chunk = lane & 3, rsub = lane >> 2;
    const uint32_t ks_ = smem_u32(Ks), vs = smem_u32(Vs), v2s = smem_u32(V2s);
#pragma unroll
    for (int r = rsub; r < NP; r += 8) {
      const int s = src_s[r];
      const uint32_t off = (uint32_t)(r * AT_LD + chunk * 8) * 2u;
      if (s >= 0) {
        cp_async16(ks_ + off, reinterpret_cast<const bf16*>(a.k) + (long long)s * a.ldk + c0 + chunk * 8, true);
        cp_async16(vs + off, reinterpret_cast<const bf16*>(a.v) + (long long)s * a.ldv + c0 + chunk * 8, true);
        if (dual) cp_async16(v2s + off, reinterpret_cast<const bf16*>(a.v2) + (long long)s * a.ldv + c0 + chunk * 8, true);
      } else {
        uint4 kv = make_uint4(0, 0, 0, 0), vv = kv, v2v = kv;
        if (s == -1) {  // zero-padded token: its projections are the biases
          const int cc = c0 + chunk * 8;
          uint32_t* kp = reinterpret_cast<uint32_t*>(&kv);
          uint32_t* vp = reinterpret_cast<uint32_t*>(&vv); uint32_t* wp = reinterpret_cast<uint32_t*>(&v2v);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (a.pad_k) kp[e] = pack_bf16(a.pad_k[(long long)b * a.pad_k_stride + cc + 2 * e], a.pad_k[(long long)b * a.pad_k_stride + cc + 2 * e + 1]);
            if (a.pad_v) vp[e] = pack_bf16(a.pad_v[cc + 2 * e], a.pad_v[cc + 2 * e + 1]);
            if (dual && a.pad_v2) wp[e] = pack_bf16(a.pad_v2[cc + 2 * e], a.pad_v2[cc + 2 * e + 1]);
          }
        }
        *reinterpret_cast<uint4*>(Ks + r * AT_LD + chunk * 8) = kv;
        *reinterpret_cast<uint4*>(Vs + r * AT_LD + chunk * 8) = vv;
        if (dual) *reinterpret_cast<uint4*>(V2s + r * AT_LD + chunk * 8) = v2v;
      }
    }
    cp_async_wait_all();
  }
  __syncwarp();

  const float scale = 0.17677669529663687f * LOG2E;  // 32^-0.5 (the reference scales q before the matmul), base-2 softmax
  const float* tab = table_s + h * NT;
  const uint32_t k_base = smem_u32(Ks), v_base = smem_u32(Vs), v2_base = smem_u32(V2s);
  constexpr int MT = (N + 15) / 16;
  // region labels of this lane's 16 score columns (only read when the window is masked)
  int collab[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = nt * 8 + cq + e;
      collab[nt][e] = (masked && j < N) ? lab_s[j] : -1;
    }

  // Q A-fragments (m16n8k16: rows gq / gq+8, column pairs cq / cq+8 of each 16-wide k-step) straight from global: 4-byte
  // loads, 16 contiguous bytes per row and quad.  Keeping Q (and the output staging) out of shared memory lets a fourth
  // CTA fit on the SM -- the kernel is issue-bound and needs the warps.
  auto load_q = [&](int mt, uint32_t (&qf)[2][4]) {
    const int s0 = src_s[mt * 16 + gq], s1 = src_s[mt * 16 + gq + 8];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = c0 + ks * 16 + hh * 8 + cq;
        uint32_t v0 = 0, v1 = 0;
        if (s0 >= 0) v0 = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const bf16*>(a.q) + (long long)s0 * a.ldq + col);
        else if (s0 == -1 && a.pad_q) v0 = pack_bf16(a.pad_q[col], a.pad_q[col + 1]);  // zero-padded token: q = the bias
        if (s1 >= 0) v1 = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const bf16*>(a.q) + (long long)s1 * a.ldq + col);
        else if (s1 == -1 && a.pad_q) v1 = pack_bf16(a.pad_q[col], a.pad_q[col + 1]);
        qf[ks][hh * 2] = v0;
        qf[ks][hh * 2 + 1] = v1;
      }
  };
  uint32_t qn[2][4];
  load_q(0, qn);
#pragma unroll 1
  for (int mt = 0; mt < MT; ++mt) {
    // ---- Q fragments of this m-tile (requested one m-tile ahead) ----
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int e = 0; e < 4; ++e) qa[ks][e] = qn[ks][e];
    if (mt + 1 < MT) load_q(mt + 1, qn);
    // ---- S = Q K^T : 8 key tiles of 8 ----
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f; }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int krow = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int kcol = ks * 16 + ((lane >> 3) & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(k_base + (krow * AT_LD + kcol) * 2, b0, b1, b2, b3);
        mma_bf16_16816(sc[2 * np], qa[ks], b0, b1);
        if (2 * np + 1 < NTV) mma_bf16_16816(sc[2 * np + 1], qa[ks], b2, b3);
      }
    }
    // ---- scale + relative-position bias + shift mask, row max ----
    const int i0 = mt * 16 + gq, i1 = i0 + 8;
    const int yi0 = i0 / WS, xi0 = i0 - yi0 * WS, yi1 = i1 / WS, xi1 = i1 - yi1 * WS;
    // bias index = (yi-yj+WS-1)*(2WS-1) + (xi-xj+WS-1) = rowpart - colpart; rows past the window read entry 0 (discarded)
    // rows past the window (discarded) use the centre entry so that rp - cp stays inside the table without a select
    constexpr int RP_PAD = (WS - 1) * (2 * WS - 1) + WS - 1;
    const int rp0 = i0 < N ? (yi0 + WS - 1) * (2 * WS - 1) + xi0 + WS - 1 : RP_PAD;
    const int rp1 = i1 < N ? (yi1 + WS - 1) * (2 * WS - 1) + xi1 + WS - 1 : RP_PAD;
    const float* tab0 = tab + rp0;
    const float* tab1 = tab + rp1;
    const int li0 = lab_s[i0], li1 = lab_s[i1];
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NTV; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float s0 = -INFINITY, s1 = -INFINITY;
        if (nt * 8 + cq + e < N) {  // (always true for 8x8 windows; folds away)
          const int cp = colpart[nt][e];
          s0 = fmaf(sc[nt][e], scale, tab0[-cp]);
          s1 = fmaf(sc[nt][2 + e], scale, tab1[-cp]);
          if (masked) {
            if (collab[nt][e] != li0) s0 += -100.0f * LOG2E;
            if (collab[nt][e] != li1) s1 += -100.0f * LOG2E;
          }
        }
        sc[nt][e] = s0;
        sc[nt][2 + e] = s1;
        mx0 = fmaxf(mx0, s0);
        mx1 = fmaxf(mx1, s1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[4][4];  // P as bf16 A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < NTV) {
        const float p00 = exp2f_fast(sc[nt][0] - mx0), p01 = exp2f_fast(sc[nt][1] - mx0);
        const float p10 = exp2f_fast(sc[nt][2] - mx1), p11 = exp2f_fast(sc[nt][3] - mx1);
        sum0 += p00 + p01;
        sum1 += p10 + p11;
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p00, p01);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p10, p11);
      } else {  // keys past the window: probability 0
        pa[nt >> 1][(nt & 1) * 2 + 0] = 0u;
        pa[nt >> 1][(nt & 1) * 2 + 1] = 0u;
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

    // ---- O = P V (and P V2): 4 dim tiles of 8, 4 k-steps of 16 keys ----
#pragma unroll 1
    for (int which = 0; which < (dual ? 2 : 1); ++which) {
      const uint32_t vb = which ? v2_base : v_base;
      float o[4][4];
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {
          const int vrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int vcol = dp * 16 + (lane >> 4) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4_trans(vb + (vrow * AT_LD + vcol) * 2, b0, b1, b2, b3);
          mma_bf16_16816(o[2 * dp], pa[kk], b0, b1);
          mma_bf16_16816(o[2 * dp + 1], pa[kk], b2, b3);
        }
      }
      // ---- outputs straight from the accumulator fragments: 4 bytes per lane, 16 contiguous bytes per row and quad ----
      bf16* outp = reinterpret_cast<bf16*>(which ? a.out2 : a.out);
      const int so0 = src_s[mt * 16 + gq], so1 = src_s[mt * 16 + gq + 8];
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        if (so0 >= 0) *reinterpret_cast<uint32_t*>(outp + (long long)so0 * a.ldo + c0 + dt * 8 + cq) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        if (so1 >= 0) *reinterpret_cast<uint32_t*>(outp + (long long)so1 * a.ldo + c0 + dt * 8 + cq) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
      }
    }
  }
  }  // cta_task
}

__global__ void window_maps_kernel(WinGeom g, int32_t* gather, int32_t* labels, int32_t* relidx) {
  const int N = g.ws * g.ws;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < g.nW * N) {
    const int win = tid / N, tok = tid - win * N;
    int y, x;
    win_source(g, win, tok, y, x);
    if (gather) gather[tid] = (y < g.H && x < g.W) ? y * g.W + x : -1;
    if (labels) labels[tid] = win_label(g, win, tok);
  }
  if (relidx && tid < N * N) relidx[tid] = rel_pos_index(tid / N, tid % N, g.ws);
}

template <int WS>
static int launch_attn(const MstWindowAttn& a, const WinGeom& g, cudaStream_t st) {
  constexpr int NT = (2 * WS - 1) * (2 * WS - 1);
  const size_t smem = (size_t)AT_WARPS * (a.v2 ? 3 : 2) * 64 * AT_LD * sizeof(bf16) + sizeof(float) * NT * a.heads + sizeof(int) * 2 * 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long tasks = (long long)a.B * g.nW * a.heads;
  const long long cta_tasks = tasks / AT_WARPS;
  if (cta_tasks > 0x7fffffffLL) return MST_ERR_BAD_ARG;
  // a few waves of resident CTAs, each looping over its share of the windows
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  long long per_sm = (227 * 1024) / (long long)(smem + 1024);  // CTAs resident per SM by shared memory (registers allow 5)
  if (per_sm > 5) per_sm = 5;
  const long long resident = (long long)sms * per_sm;
  long long grid = cta_tasks;
  if (grid > 4 * resident) grid = 4 * resident;  // >= 4 tasks per CTA keeps the tail short without paying the set-up per window
  window_attn_kernel<WS><<<(unsigned)grid, AT_WARPS * 32, smem, st>>>(a, g, (int)cta_tasks);
  return (int)cudaGetLastError();
}

}  // namespace mst

namespace mst { int attn_core_try(const MstWindowAttn& a, cudaStream_t st, bool& handled); }  // attn_core.cu

extern "C" int mst_window_attention(const MstWindowAttn* a, void* stream) {
  using namespace mst;
  if (!a || !a->q || !a->k || !a->v || !a->out || !a->bias_table) return MST_ERR_BAD_ARG;
  if ((a->v2 == nullptr) != (a->out2 == nullptr)) return MST_ERR_BAD_ARG;
  if (a->B <= 0 || a->H <= 0 || a->W <= 0 || a->heads <= 0 || a->heads > 32) return MST_ERR_BAD_ARG;
  if (a->heads % AT_WARPS != 0) return MST_ERR_UNSUPPORTED;
  if ((a->ldq | a->ldk | a->ldv | a->ldo) % 8 != 0) return MST_ERR_BAD_ARG;
  if (a->shift < 0 || a->shift >= a->ws || a->pad_k_stride < 0) return MST_ERR_BAD_ARG;
  const WinGeom g = make_geom(a->H, a->W, a->ws, a->shift);
  cudaStream_t st = (cudaStream_t)stream;
  {  // the dual (two value tensors, one softmax) passes on maps the windows tile: the tcgen05 kernel of attn_core.cu
    bool handled = false;
    const int rc = attn_core_try(*a, st, handled);
    if (handled) return rc;
  }
  if (a->ws == 8) return launch_attn<8>(*a, g, st);
  if (a->ws == 7) return launch_attn<7>(*a, g, st);
  return MST_ERR_UNSUPPORTED;
}

extern "C" int mst_window_maps(int H, int W, int ws, int shift, int32_t* gather, int32_t* labels, int32_t* relidx,
                               void* stream) {
  using namespace mst;
  if (H <= 0 || W <= 0 || ws <= 0 || shift < 0) return MST_ERR_BAD_ARG;
  const WinGeom g = make_geom(H, W, ws, shift);
  const int N = ws * ws;
  const int total = max(g.nW * N, N * N);
  window_maps_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, gather, labels, relidx);
  return (int)cudaGetLastError();
}
