/* mst_b200 -- C ABI of the B200-native Master stylization hot path.
 *
 * Plain pointers, sizes and a cudaStream_t (passed as void*); no torch types.  Every function
 * enqueues work on the given stream, never synchronises the host, never allocates persistent
 * memory, keeps no global mutable state, and returns
 *      0  = OK,   <0 = bad argument / unsupported shape,   >0 = a cudaError_t value.
 *
 * The reference (uozyurt/MasterMetaStyleTransfer) has no FFI: its boundary is the Python
 * nn.Module API.  Each entry point therefore cites the reference
 * function whose arithmetic it replaces; INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add.
 *
 * Layout conventions: activations are token-major [B,H,W,C] ("BHWC", exactly the layout
 * codes/style_transformer.py works in); bf16 = IEEE bfloat16 stored as uint16_t.
 */
#ifndef MST_B200_H
#define MST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t mst_bf16;

#define MST_OK 0
#define MST_ERR_BAD_ARG (-1)
#define MST_ERR_UNSUPPORTED (-2)

/* library version (major*10000 + minor*100 + patch) and the SM architecture it was compiled for */
int mst_version(void);
int mst_sm_arch(void);
const char* mst_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * Weight packing.  nn.Linear weight [N,K] fp32 (codes/style_transformer.py:206-213) and
 * nn.Conv2d weight [N,Cin,3,3] fp32 (codes/decoder.py:23-55, torchvision vgg19) are converted
 * to the bf16 [n_pad, k_pad] K-contiguous operand the tensor-core GEMM streams; conv k index is
 * (ky*3+kx)*Cin + ci.  Rows >= N and columns >= K are zero.  The operand is stored tile-blocked and
 * pre-swizzled: block (n_tile, k_block) is the exact 128B-swizzled shared-memory image of a
 * [tile_n x 64] K-major UMMA operand, tile_n = mst_gemm_tile_n(n_pad), so the GEMM fetches it with one
 * bulk copy.  n_pad % 16 == 0, k_pad % 64 == 0.
 * ------------------------------------------------------------------------------------------ */
int mst_gemm_tile_n(int n_pad);
int mst_pack_linear_weight(const float* w, int N, int K, mst_bf16* dst, int n_pad, int k_pad, void* stream);
int mst_pack_conv3x3_weight(const float* w, int N, int Cin, mst_bf16* dst, int n_pad, int k_pad, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core GEMM with gathered A operand and fused epilogue (tcgen05.mma, fp32 accum in TMEM).
 *     acc[m,n] = sum_k A(m,k) * Wt[n,k]
 *     x = act(acc + bias[n]);  x = mul ? res*mul + x : (res ? res + x : x);  store fp32 / bf16
 * a_mode MST_A_PLAIN:   A(m,k) = A[m*lda + k]                       (F.linear, style_transformer.py:114-116,156)
 * a_mode MST_A_CONV3X3: A(m,k) = in[b, pad(y+ky-1), pad(x+kx-1), ci] (Conv2d 3x3 stride 1;
 *                       pad_mode 0 = zeros (vgg19), 1 = reflect (codes/decoder.py:24);
 *                       upsample=1 folds nn.Upsample(2,'nearest') (decoder.py:27) into the read:
 *                       the stored input is [B,H/2,W/2,Cin]).  m = (b*H + y)*W + x.
 * ------------------------------------------------------------------------------------------ */
enum { MST_A_PLAIN = 0, MST_A_CONV3X3 = 1 };
enum { MST_ACT_NONE = 0, MST_ACT_RELU = 1, MST_ACT_GELU = 2 };
enum { MST_GATE_NONE = 0, MST_GATE_RELU = 1, MST_GATE_GELU = 2 };
enum { MST_OUT_TOKENS = 0, MST_OUT_NCHW_F32 = 1, MST_OUT_IMAGE_U8 = 2 }; /* MstGemm.out_nchw */

typedef struct MstGemm {
  const mst_bf16* A; /* bf16 activations */
  const mst_bf16* Wt; /* packed weights [n_pad, k_pad] */
  const float* bias;  /* [N] or NULL */
  const float* res;   /* fp32 [M, ld_res] or NULL (may alias out_f32) */
  const float* mul;   /* fp32 [M, ld_res] or NULL: out = res*mul + x (Query*sigma+mu, style_transformer.py:1123) */
  float* out_f32;     /* fp32 [M, ld_out32] or NULL */
  mst_bf16* out_bf16; /* bf16 [M, ld_out16] or NULL */
  int M, N, K;        /* N = the n_pad the weight was packed with (<= 1024); K = un-padded reduction length */
  int k_pad;          /* row stride of Wt, multiple of 64 */
  int lda, ld_res, ld_out32, ld_out16;
  int a_mode, act;
  int H, W, Cin, pad_mode, upsample; /* MST_A_CONV3X3 geometry (H,W = output = padded-input size) */
  int out_nchw;       /* MST_OUT_NCHW_F32: out_f32 is [B, n_real, H, W] (final decoder conv, decoder.py:54).
                       * MST_OUT_IMAGE_U8 (mst_conv3x3_rows only, others: MST_ERR_UNSUPPORTED): out_f32 points at a uint8_t
                       * [B, H, W, n_real] image and the epilogue stores (uint8) clip(x * 255, 0, 255) -- test_model.py:207's
                       * `np.clip(img * 255, 0, 255).astype(np.uint8)` on the NHWC result, without the fp32 image in HBM. */
  int n_real;         /* channels actually stored when out_nchw (<= N) */
  /* ---- training-step extensions (all optional; zero = the inference behaviour above) ----
   * epilogue order:  x = acc + bias;  out_pre16 <- x;  x = act(x);  gate;  x += add16;  x *= row_scale[m / rows_per_scale];
   *                  res / mul;  stores.
   * gate_mode MST_GATE_RELU: x = gate[m,n] > 0 ? x : 0      (ReLU backward: gate = the forward output)
   *           MST_GATE_GELU: x *= GELU'(gate[m,n])          (GELU backward: gate = the forward pre-activation)
   * conv_full = 1 (MST_A_CONV3X3, zero padding): "full" correlation used for the data gradient of a reflect-padded conv:
   *   H, W are the OUTPUT grid, the stored input is [B, H-2, W-2, Cin] and tap (ky,kx) reads in[y+ky-2, x+kx-2]. */
  const mst_bf16* gate;  /* bf16 [M, ld_gate] or NULL */
  const mst_bf16* add16; /* bf16 [M, ld_gate] or NULL */
  mst_bf16* out_pre16;   /* bf16 [M, ld_out16] or NULL: pre-activation copy kept for the backward pass */
  const float* row_scale; /* fp32 [ceil(M / rows_per_scale)] or NULL: per-sample stochastic-depth factor (tv StochasticDepth "row") */
  int gate_mode, ld_gate, rows_per_scale, conv_full;
} MstGemm;

int mst_gemm(const MstGemm* g, void* stream);

/* Same operation for a_mode == MST_A_CONV3X3 (no res/mul), im2col-free: a CTA stages R+2 input rows of a band
 * once in shared memory and the tensor core reads all nine taps from it through shifted SWIZZLE_NONE descriptors
 * (conv_band.cu).  Cin % 16 == 0.  mst_conv3x3_band_supported() says whether a band plan exists for a shape. */
int mst_conv3x3_band(const MstGemm* g, void* stream);
int mst_conv3x3_band_supported(int N, int Cin, int H, int W);

/* Row-streaming variant for the thin layers (Cin in {32,64}, N <= 64 in one tile, W % 128 == 0; decoder.py:39-54 and
 * VGG conv1_2): the padded input rows stream through a shared-memory ring while producers, the MMA issuer and the
 * epilogue run concurrently (warp-specialised), the [N x 9 Cin] weights stay resident in shared memory. */
int mst_conv3x3_rows(const MstGemm* g, void* stream);
int mst_conv3x3_rows_supported(int N, int Cin, int H, int W);

/* Channel-major variant for the wide layers (N % 128 == 0 or N == 64, Cin % 64 == 0, W in {32, 64, 128}, H % (256 / W) == 0; decoder.py:25-37
 * and VGG conv2_x .. conv4_x, loss.py:23-37): output channels are the MMA's M dimension, the pixels of ONE image row its N
 * dimension, each padded input row is staged once and all nine taps read it through row-shifted descriptors, one streamed
 * weight k-block feeds 256 pixels (conv_cm.cu).  bf16 output only (the folded nearest-x2 upsample is supported); no residual /
 * training-step extensions. */
int mst_conv3x3_cm(const MstGemm* g, void* stream);
int mst_conv3x3_cm_supported(int N, int Cin, int H, int W);

/* ------------------------------------------------------------------------------------------
 * Fused transformer MLP (torchvision ops/misc.py:264-306 MLP as used at style_transformer.py:366,839-841,991
 * and in the tv Swin blocks):  out = res + fc2(GELU_erf(fc1(A) + b1)) + b2  with the [M x 4C] hidden activation
 * kept on chip (TMEM -> GELU -> swizzled shared memory -> second tcgen05 GEMM).  C in {128, 256}.
 * Both weight matrices are packed once into one linear stream of 32 KB stages in consumption order.
 * ------------------------------------------------------------------------------------------ */
typedef struct MstMlp {
  const mst_bf16* A;       /* bf16 [M, lda] */
  const mst_bf16* Wstream; /* mst_pack_mlp_weights() output, mst_mlp_stream_bytes(C) bytes */
  const float* b1;         /* [4C] */
  const float* b2;         /* [C] */
  const float* res;        /* fp32 [M, ld_res] or NULL (may alias out_f32) */
  float* out_f32;          /* fp32 [M, ld_out32] or NULL */
  mst_bf16* out_bf16;      /* bf16 [M, ld_out16] or NULL */
  int M, C, lda, ld_res, ld_out32, ld_out16;
  /* ---- optional attention-output stage in front of the MLP (pre != 0; all zero = the plain MLP above) ----
   *   x1  = res + A . Wpre^T + bpre            (attention projection + residual: style_transformer.py:156,383-386 / tv swin block)
   *      or res * mul + A . Wpre^T + bpre      (mul != NULL: Query*sigma + mu, style_transformer.py:1123; A.Wpre^T+bpre is mu)
   *   X   = LayerNorm(x1; ln_g, ln_b) eps 1e-5, or x1 when ln_g == NULL     (bf16, never leaves shared memory)
   *   out = x1 + fc2(GELU(fc1(X) + b1)) + b2
   * res is required (it may alias out_f32); x1 itself stays on chip (pre-loaded into the fc2 accumulator); Wstream must come from
   * mst_pack_mlp_weights_pre (the [C, C] projection weight is prepended to every tile's weight stream). */
  const float* bpre;       /* [C] projection bias */
  const float* mul;        /* fp32 [M, ld_res] or NULL */
  const float* ln_g;       /* [C] or NULL */
  const float* ln_b;       /* [C] or NULL */
  int pre;
  /* ---- optional LayerNorm of the result for the NEXT block (pre != 0, else MST_ERR_UNSUPPORTED) ----
   *   out_bf16 rows < lnn_rows = LayerNorm(out; lnn_g, lnn_b) eps 1e-5 instead of the plain bf16 copy of out (rows >= lnn_rows keep
   *   the copy; out_f32 is unchanged): the following block's norm1 (tv swin_transformer.py SwinTransformerBlock.forward;
   *   style_transformer.py StyleDecoder norm1 on the content half of the encoded batch) without a launch of its own.
   *   lnn_g / lnn_b: both or neither; lnn_rows == 0 means all M rows. */
  const float* lnn_g;      /* [C] or NULL */
  const float* lnn_b;      /* [C] or NULL */
  int lnn_rows;
} MstMlp;
size_t mst_mlp_stream_bytes(int C);
size_t mst_mlp_stream_bytes_pre(int C);
int mst_pack_mlp_weights(const float* w1 /*[4C,C]*/, const float* w2 /*[C,4C]*/, mst_bf16* dst, int C, void* stream);
int mst_pack_mlp_weights_pre(const float* wpre /*[C,C]*/, const float* w1, const float* w2, mst_bf16* dst, int C, void* stream);
int mst_mlp_fused(const MstMlp* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused shifted-window attention core: roll + window partition folded into the loads, QK^T,
 * relative-position bias, 9-region shift mask, softmax, PV, window reverse + roll back folded into
 * the store (codes/style_transformer.py:83-111,127-168; with v2/out2 set it is the shared-softmax
 * sigma/mu core of :544-607).  q,k,v are the already projected bf16 token-major tensors (row
 * strides ldq/ldk/ldv elements, so slices of a fused qkv buffer work); q is scaled by
 * head_dim^-0.5 inside.  pad_q/pad_k/pad_v ([C] fp32 or NULL) are the values a zero-padded
 * token takes after its projection (= the bias, :83-85,114-116); needed when H or W is not a
 * multiple of ws.  head_dim must be 32; ws in {7, 8}.
 * ------------------------------------------------------------------------------------------ */
typedef struct MstWindowAttn {
  const mst_bf16 *q, *k, *v, *v2;
  mst_bf16 *out, *out2;
  const float* bias_table; /* [(2ws-1)^2, heads] fp32 */
  const float *pad_q, *pad_k, *pad_v, *pad_v2;
  int B, H, W, heads, ws, shift;
  int ldq, ldk, ldv, ldo;
  int pad_k_stride; /* 0: pad_k is [C]; C: pad_k is [B, C], one vector per image (the instance-normalised Wk bias of the
                       sigma/mu attention on a padded map, style_transformer.py:520-530) */
} MstWindowAttn;

int mst_window_attention(const MstWindowAttn* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused self-attention half of a window transformer block (csrc/attn_fused.cu): the three projections
 * q = x Wq^T + bq, k = x Wk^T + bk, v = x Wv^T + bv AND the shifted-window attention core above in one
 * kernel -- q, k, v never touch HBM.  Replaces F.linear x3 + shifted_window_attention up to (not
 * including) the output projection, codes/style_transformer.py:77-155 with q_in = k_in = v_in = x, and
 * the attention of a torchvision Swin block (swin_transformer.py:116-220, fused qkv weight).
 * x: bf16 token-major [B*H*W, ldx] (already LayerNorm'd where the block has a norm1); out: bf16
 * [B*H*W, ldo], heads side by side exactly like mst_window_attention's `out`.  Window partition and
 * cyclic shift are the coordinates of the TMA box that loads a window; zero-padded tokens of maps that
 * are not a multiple of ws are zero rows of x (their q/k/v = the biases, as pad-then-linear makes
 * them).  wqkv / bqkv come from mst_pack_attn_qkv.  C in {128, 256}, head_dim 32, ws in {7, 8}.
 * dbg_qkv (tests only, else NULL): bf16 [B*H*W, 3C] receives the projected q | k | v rows.
 * ------------------------------------------------------------------------------------------ */
typedef struct MstAttnBlock {
  const mst_bf16* x;
  const mst_bf16* wqkv;     /* mst_attn_qkv_packed_bytes(C, heads) bytes */
  const float* bqkv;        /* [3C] fp32, packed order */
  const float* bias_table;  /* [(2ws-1)^2, heads] fp32 */
  mst_bf16* out;
  mst_bf16* dbg_qkv;
  int B, H, W, C, heads, ws, shift;
  int ldx, ldo;
} MstAttnBlock;
size_t mst_attn_qkv_packed_bytes(int C, int heads);
/* wq, wk, wv: fp32 [C, C] row-major (nn.Linear weights; slices of a fused [3C, C] qkv weight work); bq, bk, bv: [C] or NULL */
int mst_pack_attn_qkv(const float* wq, const float* wk, const float* wv, const float* bq, const float* bk, const float* bv,
                      mst_bf16* dst_w, float* dst_b, int C, int heads, void* stream);
int mst_attn_block(const MstAttnBlock* a, void* stream);

/* integer maps the attention kernel uses, exported for the bit-exact parity tests
 * (gather: [nW, ws*ws] int32 = y*W+x or -1 for padding; labels: [nW, ws*ws] int32;
 *  relidx: [ws^4] int32).  Device pointers. */
int mst_window_maps(int H, int W, int ws, int shift, int32_t* gather, int32_t* labels, int32_t* relidx, void* stream);

/* ------------------------------------------------------------------------------------------
 * Similarity loss of the paper (codes/utils.py:105-133, codes/loss.py:137-146,321-336) without materialising the B x N x N
 * cosine self-similarity maps (csrc/similarity.cu).  For a tap tensor feat [B, N, C] bf16 token-major:
 *   mst_sim_prepare : ahat = feat / max(|feat row|, 1e-8) (bf16, [B, N, C]), svec [B, C] = per-image sum of the ahat rows,
 *                     inv_cs [B, N] = 1 / (column sum of the cosine map + 1e-6)   (the map is symmetric: column sum j = ahat_j . svec)
 *   mst_sim_tiles   : per 128x128 tile of the strict lower triangle, both maps by tcgen05.mma and
 *                     sum |D_c[i][j]*inv_c[j] - D_o[i][j]*inv_o[j]| (or squares) -> partials[mst_sim_num_tiles(B, N)]
 *   mst_sim_finalize: out[0] = sum(partials0)/count0 + sum(partials1)/count1 (two taps; count = B*N*N, the mean runs over the
 *                     whole map as torch.mean over the tril'ed tensors does)
 * N % 128 == 0, C % 256 == 0.
 * ------------------------------------------------------------------------------------------ */
int mst_sim_num_tiles(int B, int N);
int mst_sim_prepare(const mst_bf16* feat, int B, int N, int C, mst_bf16* ahat, float* svec, float* inv_cs, void* stream);
int mst_sim_tiles(const mst_bf16* ahat_c, const float* inv_cs_c, const mst_bf16* ahat_o, const float* inv_cs_o, int B, int N, int C,
                  int squared, float* partials, int n_partials, void* stream);
int mst_sim_finalize(const float* partials0, int n0, double count0, const float* partials1, int n1, double count1, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Normalisations.
 * mst_layernorm: nn.LayerNorm(C) eps 1e-5 over the last dim, fp32 in -> bf16 out
 *   (style_transformer.py:340-343,390-392; tv swin blocks).  C in {128,256,512}.
 * mst_patch_merge_layernorm: tv _patch_merging_pad + LayerNorm(4C): x [B,H,W,C] fp32 ->
 *   [B,H/2,W/2,4C] bf16 (H, W even).
 * mst_instnorm_stats: nn.InstanceNorm2d(C, affine=False) statistics of x [B,T,C] fp32 over T
 *   (style_transformer.py:468,526,1056-1057): mean[B,C], rstd[B,C].  twice=1 returns the
 *   combined scale of IN(IN(x)) (the reference normalises the query twice, :1056 then :468).
 * mst_instnorm_apply: y = (x-mean)*rstd -> bf16 and/or fp32.
 * ------------------------------------------------------------------------------------------ */
int mst_layernorm(const float* x, const float* gamma, const float* beta, mst_bf16* y, int rows, int C, void* stream);
int mst_patch_merge_layernorm(const float* x, const float* gamma, const float* beta, mst_bf16* y, int B, int H, int W,
                              int C, void* stream);
int mst_instnorm_stats(const float* x, float* mean, float* rstd, int B, int T, int C, int twice, void* stream);
/* Affine InstanceNorm2d (decoder_use_instance_norm_with_affine, style_transformer.py:982-984): gamma [C] (or NULL = 1) is folded
 * into rstd -- once: r*gamma; twice=1 (the same affine module applied to its own output): gamma^2*r/sqrt(gamma^2*var*r^2+eps) --
 * and mst_instnorm_apply_affine adds beta [C] (or NULL): y = (x-mean)*rstd + beta.  n_pad / pad_val / pad_norm as below
 * (pad_norm then includes beta); twice and n_pad > 0 together are not supported. */
int mst_instnorm_stats_affine(const float* x, float* mean, float* rstd, int B, int T, int C, int twice, int n_pad,
                              const float* pad_val, float* pad_norm, const float* gamma, const float* beta, void* stream);
int mst_instnorm_apply_affine(const float* x, const float* mean, const float* rstd, const float* beta, mst_bf16* y16, float* y32,
                              int B, int T, int C, void* stream);
/* mst_instnorm_stats_affine + mst_instnorm_apply_affine(y16) as ONE call: mean / rstd [B,C] (and pad_norm) as the former, y16 bf16
 * [B,T,C] = (x - mean) * rstd + beta as the latter, bit-identical to the sequence.  When a (image, 32-channel) slice of x fits in
 * shared memory (mst_instnorm_fused_supported: T <= 1600, T a multiple of a box height in 8..256) it runs as one kernel that
 * fetches the slice once with TMA tensor copies (instnorm_fused.cu); other shapes run the two kernels.  gamma / beta / pad_val /
 * pad_norm may be NULL (n_pad == 0 without pad_val). */
int mst_instnorm(const float* x, float* mean, float* rstd, mst_bf16* y16, int B, int T, int C, int twice, int n_pad,
                 const float* pad_val, float* pad_norm, const float* gamma, const float* beta, void* stream);
int mst_instnorm_fused_supported(int T, int C);
/* Statistics over (T, C) JOINTLY per image -- what nn.InstanceNorm2d computes when the regular-MHA decoder variant feeds it
 * [B, C, T] tensors, read as one unbatched image (style_transformer.py:1063-1119).  mean / rstd [B, C] receive the per-image
 * scalars replicated over C, so mst_instnorm_apply applies them. */
int mst_jointnorm_stats(const float* x, float* mean, float* rstd, int B, int T, int C, void* stream);
/* P[r, :] = softmax(scale * S[r, :]): fp32 scores [rows, n] -> bf16 probabilities (the single-head global attention of the
 * regular-MHA decoder variant, style_transformer.py:1100-1106); n % 4 == 0, scale > 0. */
int mst_softmax_rows(const float* S, mst_bf16* P, int rows, int n, float scale, void* stream);
/* Tile-blocked, pre-swizzled B-operand image (the format of mst_pack_linear_weight) of a bf16 matrix already on the device:
 * trans = 0: W[n][k] = src[n*ld + k]; trans = 1: W[n][k] = src[k*ld + n].  Used where an ACTIVATION is the second operand of a
 * GEMM (keys and values of the regular-MHA variant's global attention). */
int mst_pack_bf16_matrix(const mst_bf16* src, int N, int K, int ld, int trans, mst_bf16* dst, int n_pad, int k_pad, void* stream);
/* Same statistics over a map that additionally holds n_pad tokens whose value is pad_val[c] (the zero-padded positions of a
 * window-padded map after a Linear: value = its bias; style_transformer.py:520-530 normalises Wk.K over the PADDED map).
 * pad_norm [B, C] (optional) receives the normalised padding value (pad_val - mean) * rstd. */
int mst_instnorm_stats_padded(const float* x, float* mean, float* rstd, int B, int T, int C, int n_pad, const float* pad_val,
                              float* pad_norm, void* stream);
int mst_instnorm_apply(const float* x, const float* mean, const float* rstd, mst_bf16* y16, float* y32, int B, int T,
                       int C, void* stream);

/* Swin patch embedding: Conv2d(3,128,4,stride 4) + permute + LayerNorm(128) (tv swin features[0]):
 * img [B,3,S,S] fp32 NCHW -> x [B,S/4,S/4,128] fp32. */
int mst_patch_embed(const float* img, const float* w, const float* b, const float* gamma, const float* beta, float* x,
                    int B, int S, void* stream);
/* Same, with the first Swin block's norm1 (LayerNorm(128), gamma1/beta1) fused: y16 [B*(S/4)^2, 128] bf16 receives
 * LN1(x) (NULL = not wanted).  S % 64 == 0 runs on the tensor cores (mma.sync bf16, image split hi+lo, bf16 weights)
 * unless exact != 0; other sizes use the fp32 SIMT kernel followed by mst_layernorm. */
int mst_patch_embed_ln(const float* img, const float* w, const float* b, const float* gamma, const float* beta, float* x,
                       const float* gamma1, const float* beta1, mst_bf16* y16, int B, int S, int exact, void* stream);


/* uint8 image boundary either side of the path (mean3 / std3 are HOST pointers to three floats, or both NULL).
 * mst_images_u8_to_nchw: uint8 [B,H,W,3] (a decoded, resized image batch) -> fp32 [B,3,H,W] = ((u8 / 255) - mean_c) / std_c in the
 *   fp32 operation order of transforms.ToTensor() + transforms.Normalize (test_model.py:39-48, :111-125; get_dataloader.py:37-38)
 *   -- bit-exact; NULL mean3/std3 stops after the /255 (use_imagenet_normalization_for_swin = False).
 * mst_images_nchw_to_u8: fp32 [B,3,H,W] -> uint8 [B,H,W,3] = (uint8) clip(x * 255, 0, 255), test_model.py:207's
 *   `np.clip(img * 255, 0, 255).astype(np.uint8)` (truncation) -- bit-exact.  W % 4 == 0. */
int mst_images_u8_to_nchw(const uint8_t* src, float* dst, int B, int H, int W, const float* mean3, const float* std3, void* stream);
int mst_images_nchw_to_u8(const float* src, uint8_t* dst, int B, int H, int W, void* stream);
/* mst_patch_embed_ln with mst_images_u8_to_nchw folded into its image loads: img_u8 uint8 [B,S,S,3]; mean3 / std3 as above.  The
 * results equal mst_images_u8_to_nchw followed by mst_patch_embed_ln(exact = 0) bit for bit (same tcgen05 kernel, same fp32 image
 * values), without the fp32 NCHW image in HBM.  S % 16 == 0 (mst_patch_embed_ln_u8_supported), else MST_ERR_UNSUPPORTED. */
int mst_patch_embed_ln_u8(const uint8_t* img_u8, const float* mean3, const float* std3, const float* w, const float* b,
                          const float* gamma, const float* beta, float* x, const float* gamma1, const float* beta1, mst_bf16* y16,
                          int B, int S, void* stream);
int mst_patch_embed_ln_u8_supported(int S);
/* The reference's training-image transform (codes/get_dataloader.py:30-36: ToPILImage -> Resize((512,512)) -> RandomCrop((256,256))
 * -> ToTensor -> Normalize) on a decoded uint8 [H, W, 3] image, one kernel: out fp32 [3, ch, cw] = the (top, left) crop of Pillow's
 * antialiased bilinear resize, /255, normalised (mean3 / std3 host pointers, NULL = ToTensor only).  xmin / xcnt [out_w], xk
 * [out_w, ksx] (and the y arrays, [out_h]) are Pillow's resampling windows and 22-bit fixed-point coefficients for (W -> out_w) and
 * (H -> out_h), device pointers (mastermetastyletransfer_b200.data.pil_resize_coeffs).  Bit-exact against the torchvision pipeline. */
int mst_resize_crop_normalize(const uint8_t* img, int H, int W, const int32_t* xmin, const int32_t* xcnt, const int32_t* xk, int ksx,
                              const int32_t* ymin, const int32_t* ycnt, const int32_t* yk, int ksy, int top, int left, int ch, int cw,
                              const float* mean3, const float* std3, float* out, void* stream);

/* fp32 [rows, C] -> bf16 copy (A operands of the first projections) */
/* nn.Upsample(scale_factor=2, mode='nearest') on a bf16 NHWC tensor: x [B,H,W,C] -> y [B,2H,2W,C] (decoder.py:27), C % 8 == 0. */
int mst_upsample2x_nhwc(const mst_bf16* x, mst_bf16* y, int B, int H, int W, int C, void* stream);
int mst_cast_bf16(const float* x, mst_bf16* y, size_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * VGG-19 perceptual loss (codes/loss.py).  Activations are bf16 NHWC; the twelve 64..512-channel
 * convolutions go through mst_gemm (MST_A_CONV3X3, zero padding, ReLU epilogue).
 * mst_conv3x3_first: features[0] Conv2d(3,64,3,pad 1) (+ReLU features[1]) on fp32 NCHW images
 *   (loss.py:23-25) -> bf16 [B,H,W,64].
 * mst_maxpool2x2: nn.MaxPool2d(2) on bf16 [B,H,W,C] (features[4,9,18,27]).
 * mst_tap_stats: per (image, channel) mean and biased variance over T=H*W of a tap [B,T,C]; T is split over CTAs and
 *   the per-slab partial sums go through caller-provided scratch (mst_tap_stats_scratch_floats floats), combined in a
 *   fixed order so the result is deterministic
 *   (the statistics both loss terms need: loss.py:102-105 InstanceNorm2d, :122-130 mean/std).
 * mst_content_term: sum over all elements of |IN(Fc)-IN(Fcs)| (squared=0, "euclidian", loss.py:114-116)
 *   or its square (loss.py:110-112) as n_partials deterministic per-CTA partial sums.
 * mst_loss_finalize: content = sum_taps partial_sum/(B*T*C); style = sum_taps mean|mu_s-mu_o| +
 *   mean|std_s-std_o| with the unbiased std torch.std uses (loss.py:122-130,293-315);
 *   out3 = {content + lambda*style, content, style} (loss.py:243).
 * ------------------------------------------------------------------------------------------ */
int mst_conv3x3_first(const float* img, const float* w, const float* b, mst_bf16* out, int B, int H, int W, int relu,
                      void* stream);
/* BatchNorm2d (+ ReLU) of the VGG-19-BN loss variant (codes/loss.py:41-63) on a token-major [M, C] activation:
 * y = max(0, (x - mean[c]) * gamma[c] * inv_std[c] + beta[c]) -> bf16.  x = x32 (the convolution's fp32 output; the normalisation
 * must see it before any rounding to bf16) or, with x32 == NULL, y itself in place.  var_is_rstd = 1: `var` already holds
 * 1/sqrt(var + eps) (mst_instnorm_stats with B = 1, T = M: the batch statistics of train mode, which is what the reference's
 * scripts run); 0: `var` is a variance (running_var in eval mode, or mst_tap_stats output).  C % 8 == 0, C / 8 divides 256. */
int mst_bn_relu(const float* x32, mst_bf16* y, const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                size_t M, int C, int relu, int var_is_rstd, void* stream);
int mst_maxpool2x2(const mst_bf16* x, mst_bf16* y, int B, int H, int W, int C, void* stream);
size_t mst_tap_stats_scratch_floats(int B, int T, int C);
int mst_tap_stats(const mst_bf16* x, float* mean, float* var, int B, int T, int C, float* scratch, size_t scratch_floats,
                  void* stream);
int mst_content_term(const mst_bf16* fc, const mst_bf16* fo, const float* mean_c, const float* var_c, const float* mean_o,
                     const float* var_o, int B, int T, int C, int squared, float* partials, int n_partials, void* stream);

typedef struct MstLossTap {
  const float* partials; /* content-term partial sums of this tap */
  const float *mean_s, *var_s, *mean_o, *var_o; /* [B,C] statistics of the style / output taps */
  int n_partials, B, T, C;
} MstLossTap;
typedef struct MstLossTaps {
  MstLossTap tap[4];
  int n_taps;
} MstLossTaps;
int mst_loss_finalize(const MstLossTaps* taps, float lambda, int squared_style, float* out3, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training-step glue: multi-tensor optimiser updates, one launch for all parameters.
 * MstTensorTable describes up to any number of fp32 tensors through DEVICE arrays (built once per parameter
 * list): chunk_start[t] = first CTA-chunk of tensor t (chunks of mst_opt_chunk_elems() elements, prefix sums),
 * numel[t], flat_offset[t] = offset of tensor t in a flat fp32 buffer, a..d = per-tensor base pointers.
 * mst_adam_step: torch.optim.Adam (no amsgrad) as used at train_only_inner_loop.py:468-478,573-575:
 *   a = param, b = grad, c = exp_avg, d = exp_avg_sq; step is 1-based.
 * mst_reptile_delta: flat[off_t + i] = omega_t[i] - theta_t[i]  (a = theta, b = omega)  -- the tensor that is
 *   all-reduced across GPUs when each rank trained omega on its own style task.
 * mst_reptile_apply: theta_t[i] += scale * flat[off_t + i]  (a = theta); with scale = outer_lr (/ world size)
 *   this is train.py:524-534's theta += outer_lr * (omega - theta).
 * ------------------------------------------------------------------------------------------ */
typedef struct MstTensorTable {
  const int* chunk_start;
  const long long* numel;
  const long long* flat_offset;
  void* const* a;
  void* const* b;
  void* const* c;
  void* const* d;
  int n_tensors, n_chunks;
} MstTensorTable;
int mst_opt_chunk_elems(void);
int mst_adam_step(const MstTensorTable* tb, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  void* stream);
/* CUDA-graph-capturable Adam: learning rate and step count live on the device (dev_state = {float lr; int step}), so a captured
 * training step replays with the right bias correction; advance=1 increments the step first (once per optimiser step, i.e. on
 * the first parameter group). */
int mst_adam_step_dev(const MstTensorTable* tb, float beta1, float beta2, float eps, float weight_decay, void* dev_state, int advance,
                      void* stream);
int mst_reptile_delta(const MstTensorTable* tb, float* flat, void* stream);
int mst_reptile_apply(const MstTensorTable* tb, float* flat, float scale, void* stream);

/* ==========================================================================================
 * Backward kernels of the training step (train.py:515-517 loss.backward()).
 * The reference differentiates with autograd; each entry point cites the forward it is the adjoint of.
 * Data gradients of nn.Linear / Conv2d reuse mst_gemm with transposed (flipped) packed weights.
 * ========================================================================================== */

/* Weight gradient:  dW[n,k'] += sum_m dY[m,n] * X(m,k')  (fp32 atomics: shared weights accumulate over their uses).
 * x_mode MST_A_PLAIN: nn.Linear, dW is [N,K] row-major.  MST_A_CONV3X3: Conv2d 3x3, dW is [N,Cin,3,3], X is the
 * conv INPUT image [B,Hs,Ws,Cin] with the same pad_mode / upsample meaning as MstGemm.  n_real (0 = N) limits the
 * dW rows written (a 3-channel dY is stored padded to 8 channels). */
typedef struct MstWgrad {
  const mst_bf16* dY; /* bf16 [M, ld_dy] */
  const mst_bf16* X;  /* bf16 [M, ld_x] or conv input image */
  float* dW;
  int M, N, K;        /* K = X channels (linear) or 9*Cin (conv) */
  int ld_dy, ld_x, x_mode;
  int H, W, Cin, pad_mode, upsample;
  int n_real;
} MstWgrad;
int mst_wgrad(const MstWgrad* g, void* stream);
/* bias gradient: out[n] += sum_m dY[m,n] */
int mst_colsum(const mst_bf16* dY, int M, int N, int ld, float* out, void* stream);

/* Adjoint of mst_window_attention (8x8 windows, no padding): recomputes the probabilities from q,k and returns
 * dq (already multiplied by head_dim^-0.5), dk, dv (dv2) in the token-major layout of their forward tensors, and
 * accumulates the relative-position-bias-table gradient (+=).  dout2/v2/dv2 set = the shared-softmax sigma/mu core. */
typedef struct MstWindowAttnBwd {
  const mst_bf16 *q, *k, *v, *v2;
  const mst_bf16 *dout, *dout2;
  mst_bf16 *dq, *dk, *dv, *dv2;
  const float* bias_table;
  float* dbias_table; /* [(2ws-1)^2, heads] fp32, += */
  int B, H, W, heads, ws, shift;
  int ldq, ldk, ldv, ldo, lddq, lddk, lddv;
} MstWindowAttnBwd;
int mst_window_attention_bwd(const MstWindowAttnBwd* a, void* stream);

/* LayerNorm backward: dx_accum[r,:] += d/dx of LN(x)[r,:] . dy[r,:];  dgamma += sum_r dy*xhat;  dbeta += sum_r dy.
 * x fp32 (the forward input; statistics are recomputed), dy bf16. */
int mst_layernorm_bwd(const float* x, const float* gamma, const mst_bf16* dy, float* dx_accum, float* dgamma, float* dbeta,
                      int rows, int C, void* stream);
/* InstanceNorm2d(affine=False) backward over T for x [B,T,C] fp32 (twice=1: adjoint of IN(IN(x)), the double
 * normalisation of the decoder query).  Step 1 reduces per (b,c): coef[b,c,0..2] such that
 * dx = coef0*(dy - coef1) + coef2*(x - mean); step 2 applies it: dx_accum += dx (fp32) and/or dx16 = dx. */
int mst_instnorm_bwd_stats(const float* x, const void* dy, int dy_is_f32, float* coef /* [B,C,4] */, int B, int T, int C, int twice,
                           void* stream);
int mst_instnorm_bwd_apply(const float* x, const void* dy, int dy_is_f32, const float* coef, float* dx_accum, mst_bf16* dx16,
                           int B, int T, int C, void* stream);
/* y = query*sigma + mu (style_transformer.py:1123):  gquery = gy*sigma (fp32), gsigma16 = gy*query, gmu16 = gy (bf16) */
int mst_blend_bwd(const float* gy, const float* sigma, const float* query, float* gquery, mst_bf16* gsigma16, mst_bf16* gmu16,
                  size_t n, void* stream);
/* out[i] = a[i] (+ b[i]) -> fp32 and/or bf16 (gradient-stream bookkeeping) */
int mst_add_cast(const float* a, const float* b, float* out32, mst_bf16* out16, size_t n, void* stream);

/* Token-map pad / crop for window attention on zero-padded feature maps (style_transformer.py:77-87 `F.pad(x, (0,0,0,pad_r,0,pad_b))`
 * and the final `x[:, :H, :W, :]` crop, :230-232; the sigma/mu attention pads its four inputs at :476-479): src [B,Hs,Ws,token_bytes]
 * -> dst [B,Hd,Wd,token_bytes]; destination tokens outside the source map are zero-filled, source tokens outside the
 * destination map are dropped.  accumulate_f32 != 0: tokens are fp32 and dst += src (the adjoint of the pad: gradient streams
 * on the unpadded map accumulate the cropped data gradient).  token_bytes % 16 == 0, 16-byte aligned buffers. */
int mst_token_map_copy(const void* src, void* dst, int B, int Hs, int Ws, int Hd, int Wd, int token_bytes, int accumulate_f32,
                       void* stream);

/* Reflect-pad fold (adjoint of F.pad(mode='reflect') + optional nn.Upsample(2,'nearest'), decoder.py:24-27):
 * dxp bf16 [B,H+2,W+2,C] (the conv_full data gradient on the padded grid) -> dx bf16 [B,H,W,C], or [B,H/2,W/2,C]
 * summing 2x2 blocks when upsample=1; gate (bf16, shape of dx) applies the previous ReLU's mask (gate > 0). */
int mst_reflect_fold(const mst_bf16* dxp, const mst_bf16* gate, mst_bf16* dx, int B, int H, int W, int C, int upsample, void* stream);
/* MaxPool2d(2) backward fused with the preceding ReLU's mask: dx[b,y,x,c] = dy[b,y/2,x/2,c] if x is the (first) maximum
 * of its 2x2 block and > 0, else 0.  x bf16 [B,H,W,C] = pool input. */
int mst_maxpool2x2_bwd(const mst_bf16* x, const mst_bf16* dy, mst_bf16* dx, int B, int H, int W, int C, void* stream);
/* fp32 NCHW [B,3,H,W] gradient image -> bf16 NHWC [B,H,W,8] (channels 3..7 zero) */
int mst_nchw3_to_nhwc8(const float* g, mst_bf16* out, int B, int H, int W, void* stream);

/* Loss backward (adjoint of mst_content_term + mst_loss_finalize for the OUTPUT image's taps, loss.py:110-130).
 * mst_loss_bwd_stats: per (b,c) sums s[b,c,0] = sum_t g, s[b,c,1] = sum_t g*n_o with g = -phi'(IN(Fc)-IN(Fo)).
 * mst_loss_bwd_apply: dFo = w_c/(B*T*C) * rstd_o*(g - s0/T - n_o*s1/T) + w_s/(B*C) * (-psi'(dmu)/T - psi'(dstd)*(Fo-mu_o)/((T-1)*std_o)),
 *   masked by Fo > 0 (the tap is a ReLU output), bf16.  w = {w_content, w_style} is a DEVICE float[2]
 *   (= {g_total + g_content, lambda*g_total + g_style} of the incoming gradient). */
int mst_loss_bwd_stats(const mst_bf16* fc, const mst_bf16* fo, const float* mean_c, const float* var_c, const float* mean_o,
                       const float* var_o, int B, int T, int C, int squared, float* s, void* stream);
int mst_loss_bwd_apply(const mst_bf16* fc, const mst_bf16* fo, const float* mean_c, const float* var_c, const float* mean_o,
                       const float* var_o, const float* mean_s, const float* var_s, const float* s, const float* w, int B, int T,
                       int C, int squared_content, int squared_style, mst_bf16* dfo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MST_B200_H */
