// Shared device helpers for the sm_100a kernels: mbarrier / cp.async / tcgen05 / TMEM PTX wrappers
// and the window-index arithmetic that every kernel (and the map-export test hook) uses.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mst {

typedef __nv_bfloat16 bf16;

#define MST_DEVINL __device__ __forceinline__

MST_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
MST_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
MST_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MST_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
MST_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking test of a phase (try_wait may park the thread for a system-dependent time; an event loop that polls several
// barriers must not be held by one of them).
MST_DEVINL bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware may park the thread (it is woken when the phase completes) for up to `ns`
// nanoseconds before returning false, instead of returning at once.
MST_DEVINL bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (reported as a launch failure) instead of a hung GPU.  The polling loop is kept
// short (the clock is read every 256th poll only): ncu showed the spin loops of idle producer / epilogue / streamer warps as
// more than half of all executed instructions of the tensor-core kernels.  (A try_wait with a suspend-time hint was measured
// too: it removes the spinning but wakes late -- the GEMMs got 4 % slower.)
MST_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++polls & 255u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000LL) {
        printf("mst: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------- cp.async (LDGSTS) with zero fill
MST_DEVINL void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
// L1-allocating variant: 3x3-conv taps re-read each input pixel up to nine times, mostly from the same SM
MST_DEVINL void cp_async16_ca(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
MST_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
MST_DEVINL void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// mbarrier arrival that fires when all cp.async previously issued by this thread have completed
MST_DEVINL void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
MST_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
MST_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
MST_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MST_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
MST_DEVINL void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
MST_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
MST_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread
MST_DEVINL void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, for a warp-uniform issue loop: every lane executes the (uniform) descriptor arithmetic, only the lane with
// leader != 0 issues.  Keeping the loop free of divergent control flow lets ptxas hold descriptors in uniform registers
// instead of broadcasting them out of a divergent branch before every MMA.
MST_DEVINL void umma_bf16_pred(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in TENSOR MEMORY (lane = row of A, 8 columns of two 16-bit values per K = 16 step): no shared-memory read for A.
// tools/micro/umma_a_tmem.cu checks the form against the shared-memory one.  Same election scheme as umma_bf16_pred.
MST_DEVINL void umma_ts_pred(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One [128 rows x 32 bytes] K step of a swizzled K-major shared-memory tile (same descriptor as the MMA's) -> 8 TMEM columns.
// Copies and MMAs issued by one thread execute in issue order.
MST_DEVINL void tmem_cp_128x256b_pred(uint32_t taddr, uint64_t s_desc) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.cp.cta_group::1.128x256b [%0], %1;\n"
      "}\n" ::"r"(taddr),
      "l"(s_desc)
      : "memory");
}
MST_DEVINL void umma_commit_pred(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar)
      : "memory");
}
// arrive on an mbarrier when all tcgen05 ops previously issued by this thread have completed
MST_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
MST_DEVINL void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
MST_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

MST_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// registers -> TMEM: thread i of the warp writes 32 consecutive fp32 columns of lane (base_lane + i)
MST_DEVINL void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
MST_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(v[0]),
      "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
      "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
MST_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4 = 1024 B between
// 8-row groups, [46,48) version=1 (sm_100), [61,64) layout type 2 = SWIZZLE_128B.
MST_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 (bit 4), a=b=bf16 (bits 7,10),
// both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a [rows x 128 B] SW128 K-major tile
MST_DEVINL uint32_t sw128_offset(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

// ---------------------------------------------------------------- 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256)
// One lane moves a whole 32-byte sector, so a row-per-lane epilogue issues half as many L1 requests as with 128-bit
// accesses.  Addresses must be 32-byte aligned.
MST_DEVINL void st_global_256(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
               "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
MST_DEVINL void st_global_256f(void* p, const float* r) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]),
               "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}
MST_DEVINL void ld_global_256(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
MST_DEVINL void ld_global_256f(const void* p, float* r) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}

// ---------------------------------------------------------------- kernel-side views of MstGemm
// The C-ABI struct is passed to kernels BY VALUE in two pieces: GemmCore is the inference part (layout-identical to the
// leading fields of MstGemm), GemmExt the training-step extensions.  The split matters: the tensor-core kernels run at
// their register cap, and growing the by-value parameter of the inference instantiation by 48 bytes was measured to push
// ptxas into spilling in the producer / epilogue loops (16.3 -> 25.9 us on the 32768x256x256 projection).
struct GemmCore {
  const void* A; const void* Wt; const float* bias; const float* res; const float* mul;
  float* out_f32; void* out_bf16;
  int M, N, K, k_pad, lda, ld_res, ld_out32, ld_out16, a_mode, act, H, W, Cin, pad_mode, upsample, out_nchw, n_real;
};
struct GemmExt {
  const void* gate; const void* add16; void* out_pre16; const float* row_scale;
  int gate_mode, ld_gate, rows_per_scale, conv_full;
};
struct GemmNoExt {};

// ---------------------------------------------------------------- window index arithmetic (integer, bit-exact)
// Follows codes/style_transformer.py:77-111 (pad -> clamp shift -> roll(-s) -> partition) and :134-147 (mask labels).
struct WinGeom {
  int H, W;      // un-padded feature map
  int Hp, Wp;    // padded to a multiple of ws
  int ws;        // window side
  int sy, sx;    // effective shift (0 when the window covers the padded axis)
  int nwx, nW;   // windows per row / per image
};
__host__ __device__ inline WinGeom make_geom(int H, int W, int ws, int shift) {
  WinGeom g;
  g.H = H; g.W = W; g.ws = ws;
  g.Hp = H + (ws - H % ws) % ws;
  g.Wp = W + (ws - W % ws) % ws;
  g.sy = ws >= g.Hp ? 0 : shift;
  g.sx = ws >= g.Wp ? 0 : shift;
  g.nwx = g.Wp / ws;
  g.nW = (g.Hp / ws) * g.nwx;
  return g;
}
// slot (win, tok) -> source (y, x) on the padded un-rolled grid; y >= H or x >= W means zero padding
__host__ __device__ inline void win_source(const WinGeom& g, int win, int tok, int& y, int& x) {
  const int wy = win / g.nwx, wx = win - wy * g.nwx;
  const int iy = tok / g.ws, ix = tok - iy * g.ws;
  y = wy * g.ws + iy + g.sy; if (y >= g.Hp) y -= g.Hp;
  x = wx * g.ws + ix + g.sx; if (x >= g.Wp) x -= g.Wp;
}
__host__ __device__ inline int win_band(int p, int size, int ws, int s) {
  if (s == 0) return 2;
  if (p >= size - s) return 2;
  return p >= size - ws ? 1 : 0;
}
// region label of slot (win, tok) on the rolled grid; only differences of labels are ever used
__host__ __device__ inline int win_label(const WinGeom& g, int win, int tok) {
  const int wy = win / g.nwx, wx = win - wy * g.nwx;
  const int iy = tok / g.ws, ix = tok - iy * g.ws;
  return 3 * win_band(wy * g.ws + iy, g.Hp, g.ws, g.sy) + win_band(wx * g.ws + ix, g.Wp, g.ws, g.sx);
}
__host__ __device__ inline int rel_pos_index(int i, int j, int ws) {
  const int yi = i / ws, xi = i - yi * ws, yj = j / ws, xj = j - yj * ws;
  return (yi - yj + ws - 1) * (2 * ws - 1) + (xi - xj + ws - 1);
}

// Exact (erf) GELU with ONE MUFU op and 9 ALU ops per element:  1 - erf(a/sqrt2) = 2^(-a*R(a)) for a = |x| >= 0, R a
// degree-4 minimax fit (tools/micro/gelu_fit.py; R is positive and increasing, so large |x| underflow to the right limits
// without a clamp), and  gelu(x) = max(x, 0) - |x/2| * 2^(-a R(a))  for either sign.  Max |difference| to
// 0.5*x*(1+erf(x/sqrt2)) over [-8,8]: 5.4e-7, with a relative-accurate negative tail.  It replaces an Abramowitz-Stegun
// 7.1.26 evaluation (rcp + ex2, 16 ALU ops, 5.9e-7) -- the GELU epilogue of the fused MLP is issue-bound.
// A one-MUFU tanh form, 0.5x(1+tanh(x*Q(x^2))) with MUFU.TANH, was measured too: 2 ALU ops fewer, but its 2^-11 relative
// error on tanh pushed the image max-error of one parity case over the 2e-2 gate (tools/debug/path_err.py).
MST_DEVINL float gelu_erf(float x) {
  const float a = fabsf(x);
  float r = fmaf(0.0004881656787f, a, -0.007198856212f);
  r = fmaf(r, a, 0.05214627460f);
  r = fmaf(r, a, 0.4595968127f);
  r = fmaf(r, a, 1.151000023f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a * r));
  return fmaf(-fabsf(0.5f * x), e, fmaxf(x, 0.0f));
}

// The same GELU on TWO hidden units at a time in packed fp16 (HFMA2 / HMNMX2 / MUFU.EX2.F16x2): 11 instructions per pair instead of
// 11 per element.  The fused MLP's GELU phase is issue-bound (65 k elements per 128-token tile at C = 128: ~6 k cycles of fp32
// issue against 4.6 k cycles of tensor-core time for the whole tile), and its result is the fp16 A operand of the second GEMM
// anyway.  Accuracy: the polynomial and the product a*R(a) carry fp16's 2^-11 relative error into the exponent, i.e. a relative
// error of |t|*2^-11*ln2 on the SMALL term |x/2|*2^t only (absolute error < 1e-4 everywhere, < 5e-4 relative on gelu(x) for
// x > 0.25); the result is rounded to fp16 (2^-11), four times finer than the bf16 rounding it replaces.  |x| is clamped to 64 for
// the exponent (2^(-64*R(64)) is 0 in any format), so an fp16 overflow of the polynomial cannot meet a 0 * inf.
MST_DEVINL uint32_t gelu_erf_h2(float x0, float x1) {
  uint32_t x, a, r, e, m, out;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x) : "f"(x1), "f"(x0));
  asm("abs.f16x2 %0, %1;" : "=r"(a) : "r"(x));
  asm("min.f16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(0x54005400u));          // min(|x|, 64)
  // r = -R(m): Horner with the negated coefficients of gelu_erf (fp16 constants)
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(0x90009000u), "r"(m), "r"(0x1f5f1f5fu));   // -0.00048828 * m + 0.0071983
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(r), "r"(m), "r"(0xaaadaaadu));             //  ... - 0.052146
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(r), "r"(m), "r"(0xb75bb75bu));             //  ... - 0.45972
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(r), "r"(m), "r"(0xbc9bbc9bu));             //  ... - 1.1510
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(m));                                   // t = -m * R(m)
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(e) : "r"(r));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(0xb800b800u));                         // -|x| / 2
  asm("max.f16x2 %0, %1, %2;" : "=r"(x) : "r"(x), "r"(0u));                                     // max(x, 0)
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(out) : "r"(a), "r"(e), "r"(x));
  return out;
}

// d/dx of the exact GELU: Phi(x) + x*phi(x), same erf approximation (e = exp(-x^2/2) is shared by both terms)
MST_DEVINL float gelu_erf_grad(float x) {
  const float z = x * 0.70710678118654752440f, az = fabsf(z);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-az * az * 1.4426950408889634f));
  const float er = copysignf(fmaf(-p * t, e, 1.0f), z);
  return fmaf(x * 0.3989422804014327f, e, 0.5f * (1.0f + er));
}

MST_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
MST_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace mst
