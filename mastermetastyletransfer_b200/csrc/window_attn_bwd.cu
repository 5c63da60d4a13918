// Backward of the fused shifted-window attention core (adjoint of window_attn.cu; the reference differentiates
// codes/style_transformer.py:127-155 / :544-594 with autograd).  8x8 or 7x7 windows on a map that is a window multiple (the
// training path materialises the reference's zero padding, train_engine._pad16); a 7x7 window occupies 49 of the 64 token
// slots: the other slots load zeros, are masked out of every softmax, carry zero dO and are never stored.
//
// One warp per (window, head), four warps per CTA, all matmuls on mma.sync m16n8k16 (bf16 in, fp32 accumulate):
//   pass 1 (16 query rows at a time):  S = scale*Q K^T + bias + mask,  P = softmax(S),  dP = dO V^T (+ dO2 V2^T),
//            D_i = sum_j P_ij dP_ij,  dS = P*(dP - D),  dQ = scale * dS K,  dbias[idx(i,j)] += dS_ij;
//            row max / 1/sum / D_i are parked in shared memory.
//   pass 2 (16 key rows at a time):    S^T = scale*K Q^T recomputed in the transposed orientation so that P^T and dS^T
//            come out directly as A fragments:  dV = P^T dO (dV2 = P^T dO2),  dK = scale * dS^T Q.
// Nothing of size N x N ever touches shared or global memory; roll / partition are address arithmetic as in the forward.
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

constexpr int AB_WARPS = 4;
constexpr int AB_LD = 40;  // bf16 row stride (80 B): conflict-free ldmatrix
constexpr int AB_N = 64;
constexpr int AB_NT = 225;

MST_DEVINL void ab_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
MST_DEVINL void ab_ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
MST_DEVINL void ab_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
MST_DEVINL uint32_t ab_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// A fragments of rows [row0, row0+16) x 32 dims (two k-steps) of a staged [64 x AB_LD] tile
MST_DEVINL void ab_load_a(uint32_t base, int row0, int lane, uint32_t (&a)[2][4]) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int row = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int col = ks * 16 + (lane >> 4) * 8;
    ab_ldsm_x4(base + (row * AB_LD + col) * 2, a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
  }
}
// acc[16 x 64] += A[16 x 32] * B^T, B = staged [64 x 32] tile (rows = output columns)
MST_DEVINL void ab_mm_nt(float (&acc)[8][4], const uint32_t (&a)[2][4], uint32_t bbase, int lane) {
#pragma unroll
  for (int np = 0; np < 4; ++np) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int brow = np * 16 + (lane & 7) + (lane >> 4) * 8;
      const int bcol = ks * 16 + ((lane >> 3) & 1) * 8;
      uint32_t b0, b1, b2, b3;
      ab_ldsm_x4(bbase + (brow * AB_LD + bcol) * 2, b0, b1, b2, b3);
      ab_mma(acc[2 * np], a[ks], b0, b1);
      ab_mma(acc[2 * np + 1], a[ks], b2, b3);
    }
  }
}
// o[16 x 32] = P[16 x 64] * B, B = staged [64 x 32] tile (rows = reduction index)
MST_DEVINL void ab_mm_nn(float (&o)[4][4], const uint32_t (&pa)[4][4], uint32_t bbase, int lane) {
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int dp = 0; dp < 2; ++dp) {
      const int vrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int vcol = dp * 16 + (lane >> 4) * 8;
      uint32_t b0, b1, b2, b3;
      ab_ldsm_x4_trans(bbase + (vrow * AB_LD + vcol) * 2, b0, b1, b2, b3);
      ab_mma(o[2 * dp], pa[kk], b0, b1);
      ab_mma(o[2 * dp + 1], pa[kk], b2, b3);
    }
  }
}
// fp32 [16 x 64] accumulator fragments -> bf16 A fragments (4 k-steps of 16)
MST_DEVINL void ab_pack_a(const float (&x)[8][4], uint32_t (&pa)[4][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    pa[nt >> 1][(nt & 1) * 2 + 0] = ab_pack(x[nt][0], x[nt][1]);
    pa[nt >> 1][(nt & 1) * 2 + 1] = ab_pack(x[nt][2], x[nt][3]);
  }
}
// stage a [16 x 32] fp32 fragment tile (times `mul`) and store it as 16-byte chunks to rows src_s[row0 + r] of `out`
MST_DEVINL void ab_store_tile(const float (&o)[4][4], float mul, bf16* stg, bf16* out, int ld, int c0, const int* src_s, int row0,
                              int lane) {
  const int gq = lane >> 2, cq = (lane & 3) * 2;
  __syncwarp();
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    *reinterpret_cast<uint32_t*>(stg + gq * AB_LD + dt * 8 + cq) = ab_pack(o[dt][0] * mul, o[dt][1] * mul);
    *reinterpret_cast<uint32_t*>(stg + (gq + 8) * AB_LD + dt * 8 + cq) = ab_pack(o[dt][2] * mul, o[dt][3] * mul);
  }
  __syncwarp();
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int r = rr * 8 + (lane >> 2);
    const int s = src_s[row0 + r];
    const uint4 val = *reinterpret_cast<const uint4*>(stg + r * AB_LD + (lane & 3) * 8);
    if (s >= 0) *reinterpret_cast<uint4*>(out + (long long)s * ld + c0 + (lane & 3) * 8) = val;
  }
  __syncwarp();
}

template <int WS>
__global__ void __launch_bounds__(AB_WARPS * 32) window_attn_bwd_kernel(const MstWindowAttnBwd a, const WinGeom g) {
  constexpr int NTOK = WS * WS;                      // real tokens of a window (<= AB_N slots)
  constexpr int NT = (2 * WS - 1) * (2 * WS - 1);    // relative-position table entries per head (<= AB_NT)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int heads = a.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dual = a.v2 != nullptr;
  const int tiles_per_warp = dual ? 6 : 4;
  // per warp: Q K V dO (V2 dO2) tiles, a 16-row staging tile, 3 x 64 row statistics
  const size_t warp_bytes = (size_t)(tiles_per_warp * AB_N + 16) * AB_LD * sizeof(bf16) + 3 * AB_N * sizeof(float);
  uint8_t* wbase = smem_raw + warp * warp_bytes;
  bf16* Qs = reinterpret_cast<bf16*>(wbase);
  bf16* Ks = Qs + AB_N * AB_LD;
  bf16* Vs = Ks + AB_N * AB_LD;
  bf16* dOs = Vs + AB_N * AB_LD;
  bf16* V2s = dOs + AB_N * AB_LD;
  bf16* dO2s = V2s + AB_N * AB_LD;
  bf16* stg = Qs + tiles_per_warp * AB_N * AB_LD;
  float* mrow = reinterpret_cast<float*>(stg + 16 * AB_LD);
  float* irow = mrow + AB_N;
  float* drow = irow + AB_N;
  float* table_s = reinterpret_cast<float*>(smem_raw + AB_WARPS * warp_bytes);  // [heads][225]
  float* dtab_s = table_s + AB_NT * heads;                                     // [AB_WARPS][225]
  int* src_s = reinterpret_cast<int*>(dtab_s + AB_WARPS * AB_NT);
  int* lab_s = src_s + AB_N;

  const int task0 = blockIdx.x * AB_WARPS;
  const int win_global = task0 / heads;
  const int h = task0 - win_global * heads + warp;
  const int b = win_global / g.nW;
  const int win = win_global - b * g.nW;
  const bool masked = (g.sy + g.sx) > 0;

  for (int i = threadIdx.x; i < NT * heads; i += blockDim.x) {
    const int idx = i / heads, hh = i - idx * heads;
    table_s[hh * AB_NT + idx] = a.bias_table[i];
  }
  for (int i = threadIdx.x; i < AB_WARPS * AB_NT; i += blockDim.x) dtab_s[i] = 0.f;
  for (int i = threadIdx.x; i < AB_N; i += blockDim.x) {
    if (i < NTOK) {
      int y, x;
      win_source(g, win, i, y, x);
      src_s[i] = (b * g.H + y) * g.W + x;
      lab_s[i] = win_label(g, win, i);
    } else {
      src_s[i] = -1;
      lab_s[i] = -1;
    }
  }
  __syncthreads();

  const int c0 = h * 32;
  {
    const int chunk = lane & 3, rsub = lane >> 2;
    const uint32_t qs = smem_u32(Qs), ks_ = smem_u32(Ks), vs = smem_u32(Vs), dos = smem_u32(dOs), v2s = smem_u32(V2s), do2s = smem_u32(dO2s);
#pragma unroll
    for (int r = rsub; r < AB_N; r += 8) {
      const bool real = src_s[r] >= 0;  // empty slots of a 7x7 window: zero fill
      const int s = real ? src_s[r] : 0;
      const uint32_t off = (uint32_t)(r * AB_LD + chunk * 8) * 2u;
      const int cc = c0 + chunk * 8;
      cp_async16(qs + off, reinterpret_cast<const bf16*>(a.q) + (long long)s * a.ldq + cc, real);
      cp_async16(ks_ + off, reinterpret_cast<const bf16*>(a.k) + (long long)s * a.ldk + cc, real);
      cp_async16(vs + off, reinterpret_cast<const bf16*>(a.v) + (long long)s * a.ldv + cc, real);
      cp_async16(dos + off, reinterpret_cast<const bf16*>(a.dout) + (long long)s * a.ldo + cc, real);
      if (dual) {
        cp_async16(v2s + off, reinterpret_cast<const bf16*>(a.v2) + (long long)s * a.ldv + cc, real);
        cp_async16(do2s + off, reinterpret_cast<const bf16*>(a.dout2) + (long long)s * a.ldo + cc, real);
      }
    }
    cp_async_wait_all();
  }
  __syncwarp();

  const float scale = 0.17677669529663687f;
  const float* tab = table_s + h * AB_NT;
  float* dtab = dtab_s + warp * AB_NT;
  const int gq = lane >> 2, cq = (lane & 3) * 2;
  const uint32_t q_base = smem_u32(Qs), k_base = smem_u32(Ks), v_base = smem_u32(Vs), do_base = smem_u32(dOs);
  const uint32_t v2_base = smem_u32(V2s), do2_base = smem_u32(dO2s);
  // per-lane column constants: column c = nt*8 + cq + e
  // (empty slots use token 0's table indices -- in range, and their scores are forced to -inf / their dS is exactly zero)
  int colcp[8][2], colrp[8][2], collab[8][2];
  uint32_t colreal = 0;  // bit nt*2+e: column is a real token
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int jslot = nt * 8 + cq + e;
      const int j = jslot < NTOK ? jslot : 0;
      if (jslot < NTOK) colreal |= 1u << (nt * 2 + e);
      const int yj = j / WS, xj = j - yj * WS;
      colcp[nt][e] = yj * (2 * WS - 1) + xj;
      colrp[nt][e] = (yj + WS - 1) * (2 * WS - 1) + xj + WS - 1;
      collab[nt][e] = lab_s[j];
    }
  (void)colreal;

  // ------------------------------------------------------------------ pass 1: query tiles
#pragma unroll 1
  for (int mt = 0; mt < 4; ++mt) {
    uint32_t af[2][4];
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f; }
    ab_load_a(q_base, mt * 16, lane, af);
    ab_mm_nt(sc, af, k_base, lane);
    const int i0 = mt * 16 + gq, i1 = i0 + 8;
    const int t0 = i0 < NTOK ? i0 : 0, t1 = i1 < NTOK ? i1 : 0;
    const int rp0 = (t0 / WS + WS - 1) * (2 * WS - 1) + (t0 % WS) + WS - 1;
    const int rp1 = (t1 / WS + WS - 1) * (2 * WS - 1) + (t1 % WS) + WS - 1;
    const int li0 = lab_s[t0], li1 = lab_s[t1];
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float s0 = fmaf(sc[nt][e], scale, tab[rp0 - colcp[nt][e]]);
        float s1 = fmaf(sc[nt][2 + e], scale, tab[rp1 - colcp[nt][e]]);
        if (masked) {
          if (collab[nt][e] != li0) s0 += -100.0f;
          if (collab[nt][e] != li1) s1 += -100.0f;
        }
        if (NTOK < AB_N && !((colreal >> (nt * 2 + e)) & 1u)) s0 = s1 = -INFINITY;  // empty key slot
        sc[nt][e] = s0;
        sc[nt][2 + e] = s1;
        mx0 = fmaxf(mx0, s0);
        mx1 = fmaxf(mx1, s1);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      sc[nt][0] = __expf(sc[nt][0] - mx0); sc[nt][1] = __expf(sc[nt][1] - mx0);
      sc[nt][2] = __expf(sc[nt][2] - mx1); sc[nt][3] = __expf(sc[nt][3] - mx1);
      sum0 += sc[nt][0] + sc[nt][1];
      sum1 += sc[nt][2] + sc[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    // dP = dO V^T (+ dO2 V2^T)
    float dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f; }
    ab_load_a(do_base, mt * 16, lane, af);
    ab_mm_nt(dp, af, v_base, lane);
    if (dual) {
      ab_load_a(do2_base, mt * 16, lane, af);
      ab_mm_nt(dp, af, v2_base, lane);
    }
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      sc[nt][0] *= inv0; sc[nt][1] *= inv0; sc[nt][2] *= inv1; sc[nt][3] *= inv1;
      d0 += sc[nt][0] * dp[nt][0] + sc[nt][1] * dp[nt][1];
      d1 += sc[nt][2] * dp[nt][2] + sc[nt][3] * dp[nt][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    if ((lane & 3) == 0) {
      mrow[i0] = mx0; irow[i0] = inv0; drow[i0] = d0;
      mrow[i1] = mx1; irow[i1] = inv1; drow[i1] = d1;
    }
    // dS = P * (dP - D); bias-table gradient; dQ = scale * dS K
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float ds0 = sc[nt][e] * (dp[nt][e] - d0);
        const float ds1 = sc[nt][2 + e] * (dp[nt][2 + e] - d1);
        dp[nt][e] = ds0;
        dp[nt][2 + e] = ds1;
        if (a.dbias_table) {
          atomicAdd(&dtab[rp0 - colcp[nt][e]], ds0);
          atomicAdd(&dtab[rp1 - colcp[nt][e]], ds1);
        }
      }
    uint32_t pa[4][4];
    ab_pack_a(dp, pa);
    float o[4][4];
    ab_mm_nn(o, pa, k_base, lane);
    ab_store_tile(o, scale, stg, reinterpret_cast<bf16*>(a.dq), a.lddq, c0, src_s, mt * 16, lane);
  }
  __syncwarp();

  // ------------------------------------------------------------------ pass 2: key tiles (transposed orientation)
#pragma unroll 1
  for (int jt = 0; jt < 4; ++jt) {
    uint32_t af[2][4];
    float st[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f; }
    ab_load_a(k_base, jt * 16, lane, af);
    ab_mm_nt(st, af, q_base, lane);  // rows = keys j, columns = queries i
    const int j0 = jt * 16 + gq, j1 = j0 + 8;
    const int u0 = j0 < NTOK ? j0 : 0, u1 = j1 < NTOK ? j1 : 0;
    const int cp0 = (u0 / WS) * (2 * WS - 1) + (u0 % WS);
    const int cp1 = (u1 / WS) * (2 * WS - 1) + (u1 % WS);
    const int lj0 = lab_s[u0], lj1 = lab_s[u1];
    float dpt[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f; }
    ab_load_a(v_base, jt * 16, lane, af);
    ab_mm_nt(dpt, af, do_base, lane);
    if (dual) {
      ab_load_a(v2_base, jt * 16, lane, af);
      ab_mm_nt(dpt, af, do2_base, lane);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = nt * 8 + cq + e;
        const float mi = mrow[i], ii = irow[i], di = drow[i];
        float s0 = fmaf(st[nt][e], scale, tab[colrp[nt][e] - cp0]);
        float s1 = fmaf(st[nt][2 + e], scale, tab[colrp[nt][e] - cp1]);
        if (masked) {
          if (collab[nt][e] != lj0) s0 += -100.0f;
          if (collab[nt][e] != lj1) s1 += -100.0f;
        }
        if (NTOK < AB_N) {  // empty key slots (rows here): P = 0
          if (j0 >= NTOK) s0 = -INFINITY;
          if (j1 >= NTOK) s1 = -INFINITY;
        }
        const float p0 = __expf(s0 - mi) * ii, p1 = __expf(s1 - mi) * ii;
        st[nt][e] = p0;
        st[nt][2 + e] = p1;
        dpt[nt][e] = p0 * (dpt[nt][e] - di);
        dpt[nt][2 + e] = p1 * (dpt[nt][2 + e] - di);
      }
    uint32_t pa[4][4];
    float o[4][4];
    ab_pack_a(st, pa);
    ab_mm_nn(o, pa, do_base, lane);
    ab_store_tile(o, 1.0f, stg, reinterpret_cast<bf16*>(a.dv), a.lddv, c0, src_s, jt * 16, lane);
    if (dual) {
      ab_mm_nn(o, pa, do2_base, lane);
      ab_store_tile(o, 1.0f, stg, reinterpret_cast<bf16*>(a.dv2), a.lddv, c0, src_s, jt * 16, lane);
    }
    ab_pack_a(dpt, pa);
    ab_mm_nn(o, pa, q_base, lane);
    ab_store_tile(o, scale, stg, reinterpret_cast<bf16*>(a.dk), a.lddk, c0, src_s, jt * 16, lane);
  }

  if (a.dbias_table) {
    __syncthreads();
    for (int i = threadIdx.x; i < AB_WARPS * AB_NT; i += blockDim.x) {
      const int w = i / AB_NT, idx = i - w * AB_NT;
      const int hh = task0 - win_global * heads + w;
      const float v = dtab_s[i];
      if (v != 0.f) atomicAdd(a.dbias_table + idx * heads + hh, v);
    }
  }
}

}  // namespace mst

extern "C" int mst_window_attention_bwd(const MstWindowAttnBwd* a, void* stream) {
  using namespace mst;
  if (!a || !a->q || !a->k || !a->v || !a->dout || !a->dq || !a->dk || !a->dv || !a->bias_table) return MST_ERR_BAD_ARG;
  const bool dual = a->v2 != nullptr;
  if (dual != (a->dout2 != nullptr) || dual != (a->dv2 != nullptr)) return MST_ERR_BAD_ARG;
  if (a->B <= 0 || a->H <= 0 || a->W <= 0 || a->heads <= 0 || a->heads > 32) return MST_ERR_BAD_ARG;
  if ((a->ws != 8 && a->ws != 7) || a->H % a->ws != 0 || a->W % a->ws != 0 || a->heads % AB_WARPS != 0) return MST_ERR_UNSUPPORTED;
  if ((a->ldq | a->ldk | a->ldv | a->ldo | a->lddq | a->lddk | a->lddv) % 8 != 0) return MST_ERR_BAD_ARG;
  if (a->shift < 0 || a->shift >= a->ws) return MST_ERR_BAD_ARG;
  const WinGeom g = make_geom(a->H, a->W, a->ws, a->shift);
  const size_t warp_bytes = (size_t)((dual ? 6 : 4) * AB_N + 16) * AB_LD * sizeof(bf16) + 3 * AB_N * sizeof(float);
  const size_t smem = AB_WARPS * warp_bytes + sizeof(float) * AB_NT * a->heads + sizeof(float) * AB_WARPS * AB_NT + sizeof(int) * 2 * AB_N;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attn_bwd_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (smem > 160 * 1024) return MST_ERR_UNSUPPORTED;
  const long long tasks = (long long)a->B * g.nW * a->heads;
  if (a->ws == 8)
    window_attn_bwd_kernel<8><<<(unsigned)(tasks / AB_WARPS), AB_WARPS * 32, smem, (cudaStream_t)stream>>>(*a, g);
  else
    window_attn_bwd_kernel<7><<<(unsigned)(tasks / AB_WARPS), AB_WARPS * 32, smem, (cudaStream_t)stream>>>(*a, g);
  return (int)cudaGetLastError();
}
