"""Forward engine of the stylization hot path: sequences the C-ABI kernels over pre-allocated
HBM buffers.  The nn.Module mirrors in this package (full_model.py, style_transformer.py,
decoder.py) own the parameters; this file only reads them (packed once per parameter version).

Data layout in HBM: every activation is token-major [B*H*W, C]; the residual
stream of each sub-network is fp32, every tensor-core A operand is a bf16 copy written by the
producing kernel's epilogue.  Window partition / cyclic shift are never materialised: they are
address arithmetic inside the attention kernel.
"""
from __future__ import annotations

from typing import Dict, Optional

import os

import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, PAD_REFLECT, PackedMatrix


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class Workspace:
    """Named, lazily created, shape-checked device buffers (PyTorch's caching allocator owns the memory)."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}
        # Buffers superseded by a larger request stay alive: a CUDA graph captured earlier (GraphedStylizer,
        # GraphedTrainStep) has their addresses baked in and still writes to them on replay -- handing the memory back to the
        # caching allocator would let another tensor reuse it.
        self.retired = []

    def get(self, name: str, shape, dtype) -> torch.Tensor:
        t = self.bufs.get(name)
        n = 1
        for s in shape:
            n *= s
        if t is None or t.numel() < n or t.dtype != dtype:
            if t is not None:
                self.retired.append(t)
            t = torch.empty(n, dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t[:n].view(*shape)

    def bf16(self, name, *shape):
        return self.get(name, shape, torch.bfloat16)

    def f32(self, name, *shape):
        return self.get(name, shape, torch.float32)


# --------------------------------------------------------------------------------------------
# Swin-B first two stages (reference: codes/utils.py:59-102 slice of torchvision swin_b;
# arithmetic tv swin_transformer.py:116-220,35-87,401-456)
# --------------------------------------------------------------------------------------------


# MST_FUSE_PROJ_MLP=0 falls back to the separate projection GEMM / LayerNorm / MLP kernels (A/B measurements)
FUSE_PROJ_MLP = os.environ.get("MST_FUSE_PROJ_MLP", "1") != "0"
# MST_FUSE_ATTN=0 falls back to the stand-alone QKV GEMM + window-attention kernel for the self-attentions (A/B measurements);
# default: LN1(x) -> ONE kernel (three projections + shifted-window attention, q/k/v stay on chip, csrc/attn_fused.cu)
FUSE_ATTN = os.environ.get("MST_FUSE_ATTN", "1") != "0"


def _can_fuse_attn(C: int, heads: int, ws: int) -> bool:
    return FUSE_ATTN and C in (128, 256) and heads * 32 == C and ws in (7, 8)


class SwinEncoderWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        g = lambda k: _f32(sd[prefix + k])
        self.pe_w, self.pe_b = g("0.0.weight"), g("0.0.bias")
        self.pe_g, self.pe_beta = g("0.2.weight"), g("0.2.bias")
        self.blocks = {}
        for name, heads in (("1.0", 4), ("1.1", 4), ("3.0", 8), ("3.1", 8)):
            p = name + "."
            qkv_b = g(p + "attn.qkv.bias")
            C = qkv_b.numel() // 3
            self.blocks[name] = dict(
                heads=heads, C=C,
                n1w=g(p + "norm1.weight"), n1b=g(p + "norm1.bias"), n2w=g(p + "norm2.weight"), n2b=g(p + "norm2.bias"),
                qkv=ops.pack_linear(g(p + "attn.qkv.weight"), qkv_b),
                attn=(lambda w_, b_, C_=C, h_=heads: ops.pack_attn_qkv(w_[:C_], w_[C_:2 * C_], w_[2 * C_:], b_[:C_], b_[C_:2 * C_], b_[2 * C_:], h_))(
                    g(p + "attn.qkv.weight"), qkv_b) if _can_fuse_attn(C, heads, 7) else None,
                pad_q=qkv_b[:C].contiguous(), pad_k=qkv_b[C:2 * C].contiguous(), pad_v=qkv_b[2 * C:].contiguous(),
                proj=ops.pack_linear(g(p + "attn.proj.weight"), g(p + "attn.proj.bias")),
                table=g(p + "attn.relative_position_bias_table"),
                mlp=ops.pack_mlp(g(p + "mlp.0.weight"), g(p + "mlp.0.bias"), g(p + "mlp.3.weight"), g(p + "mlp.3.bias")),
                # attention projection + residual + norm2 + MLP + residual as ONE kernel (MstMlp::pre)
                proj_mlp=ops.pack_mlp(g(p + "mlp.0.weight"), g(p + "mlp.0.bias"), g(p + "mlp.3.weight"), g(p + "mlp.3.bias"),
                                      wpre=g(p + "attn.proj.weight"), bpre=g(p + "attn.proj.bias")),
            )
        self.pm_g, self.pm_b = g("2.norm.weight"), g("2.norm.bias")
        self.pm_red = ops.pack_linear(g("2.reduction.weight"), None)


def _swin_block(bw: dict, x32: torch.Tensor, ws_: Workspace, Bt: int, H: int, W: int, shift: int, tag: str,
                ln1_done: bool = False, out16: Optional[torch.Tensor] = None, next_ln=None) -> bool:
    """x += attn(LN1(x)); x += mlp(LN2(x)) on the fp32 residual stream x32 [T, C] (in place).
    ln1_done: the producer of x32 already wrote LN1(x) into the block's `ln` buffer (patch embedding, or the previous block).
    next_ln: (gamma, beta) of the FOLLOWING block's norm1 (same tag, so the same `ln` buffer): when the MLP kernel can apply it in
    its output epilogue it does, and the block returns True -- call the next block with ln1_done=True.  (Asked for at C = 128 only:
    the C = 256 instantiation has no idle on-chip memory for the row-statistics exchange, every warp re-reads the whole row from TMEM,
    and that cost more (+19 us per launch) than the LayerNorm launch it saves (14-17 us) -- DESIGN.md section 8.)"""
    C, heads = bw["C"], bw["heads"]
    T = Bt * H * W
    ln = ws_.bf16(tag + "ln", T, C)
    qkv = ws_.bf16(tag + "qkv", T, 3 * C)
    o = ws_.bf16(tag + "o", T, C)
    if not ln1_done:
        ops.layernorm(x32, bw["n1w"], bw["n1b"], ln, T, C)
    if bw["attn"] is not None:  # q | k | v projections + window attention in one kernel
        ops.attn_block(ln, bw["attn"], bw["table"], o, Bt, H, W, 7, shift)
    else:
        ops.gemm(ln, bw["qkv"], T, out_bf16=qkv)
        ops.window_attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o, bw["table"], Bt, H, W, heads, 7, shift,
                             3 * C, 3 * C, 3 * C, C, pad_q=bw["pad_q"], pad_k=bw["pad_k"], pad_v=bw["pad_v"])
    if FUSE_PROJ_MLP:  # x1 = x + proj(o); x = x1 + mlp(LN2(x1)): one kernel, LN2(x1) and the hidden activation stay on chip
        fuse_next = next_ln is not None and out16 is None and ops.mlp_next_ln_supported(C)
        ops.mlp_fused(o, bw["proj_mlp"], T, res=x32, out_f32=x32, out_bf16=ln if fuse_next else out16, pre=True, ln_g=bw["n2w"], ln_b=bw["n2b"],
                      next_ln=next_ln if fuse_next else None)
        return fuse_next
    ops.gemm(o, bw["proj"], T, res=x32, out_f32=x32)
    ops.layernorm(x32, bw["n2w"], bw["n2b"], ln, T, C)
    ops.mlp_fused(ln, bw["mlp"], T, res=x32, out_f32=x32, out_bf16=out16)  # fc1 + GELU + fc2 + residual, hidden kept on chip
    return False


def swin_encode(w: SwinEncoderWeights, imgs, ws_: Workspace, S: int, out32: torch.Tensor, out16: Optional[torch.Tensor],
                u8_norm=None):
    """imgs: list of [B,3,S,S] fp32 NCHW tensors (content, style) encoded as one batch -- or of uint8 [B,S,S,3] images, converted
    inside the patch-embedding kernel (u8_norm = (mean, std) of transforms.Normalize, None: ToTensor only).
    out32: fp32 [sum(B), S/8, S/8, 256]; out16: optional bf16 copy."""
    u8_mean, u8_std = u8_norm if u8_norm is not None else (None, None)
    Bt = sum(int(i.shape[0]) for i in imgs)
    P = S // 4
    x1 = ws_.f32("sw_x1", Bt * P * P, 128)
    ln1 = ws_.bf16("sw1_ln", Bt * P * P, 128)  # the first block's LN1 output comes out of the patch-embedding kernel
    b10 = w.blocks["1.0"]
    off = 0
    for img in imgs:
        b = int(img.shape[0])
        ops.patch_embed(img, w.pe_w, w.pe_b, w.pe_g, w.pe_beta, x1[off * P * P:], b, S,
                        gamma1=b10["n1w"], beta1=b10["n1b"], y16=ln1[off * P * P:], u8_mean=u8_mean, u8_std=u8_std)
        off += b
    b11 = w.blocks["1.1"]
    done = _swin_block(b10, x1, ws_, Bt, P, P, 0, "sw1_", ln1_done=True, next_ln=(b11["n1w"], b11["n1b"]))
    _swin_block(b11, x1, ws_, Bt, P, P, 3, "sw1_", ln1_done=done)
    P2 = P // 2
    T2 = Bt * P2 * P2
    pm = ws_.bf16("sw_pm", T2, 512)
    ops.patch_merge_layernorm(x1, w.pm_g, w.pm_b, pm, Bt, P, P, 128)
    x2 = out32.view(T2, 256)
    ops.gemm(pm, w.pm_red, T2, out_f32=x2)
    _swin_block(w.blocks["3.0"], x2, ws_, Bt, P2, P2, 0, "sw2_")
    # the bf16 copy of the features (the style transformer's first operands) comes out of the last block's MLP epilogue
    _swin_block(w.blocks["3.1"], x2, ws_, Bt, P2, P2, 3, "sw2_", out16=out16.view(T2, 256) if out16 is not None else None)


# --------------------------------------------------------------------------------------------
# Style transformer (codes/style_transformer.py:777-1245, default flags)
# --------------------------------------------------------------------------------------------


class StyleTransformerWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        g = lambda k: _f32(sd[prefix + k])

        def mlp(p, proj=None):
            if proj is None:
                return ops.pack_mlp(g(p + "0.weight"), g(p + "0.bias"), g(p + "3.weight"), g(p + "3.bias"))
            return ops.pack_mlp(g(p + "0.weight"), g(p + "0.bias"), g(p + "3.weight"), g(p + "3.bias"),
                                wpre=g(proj + "weight"), bpre=g(proj + "bias"))

        e = "encoder.shared_MHA_without_MLP.attn."
        wq, wk, wv = g(e + "Wq.weight"), g(e + "Wk.weight"), g(e + "Wv.weight")
        bq, bk, bv = g(e + "Wq.bias"), g(e + "Wk.bias"), g(e + "Wv.bias")
        self.C = wq.shape[0]
        self.enc_qkv = ops.pack_linear(torch.cat([wq, wk, wv], 0), torch.cat([bq, bk, bv], 0))
        self.heads = int(g(e + "relative_position_bias_table").shape[1])
        self.enc_attn = ops.pack_attn_qkv(wq, wk, wv, bq, bk, bv, self.heads) if _can_fuse_attn(self.C, self.heads, 8) else None
        self.enc_qk = ops.pack_linear(torch.cat([wq, wk], 0), torch.cat([bq, bk], 0))
        self.enc_v = ops.pack_linear(wv, bv)
        self.enc_pad = (bq, bk, bv)
        self.enc_proj = ops.pack_linear(g(e + "proj.weight"), g(e + "proj.bias"))
        self.enc_table = g(e + "relative_position_bias_table")
        self.mlp_key = mlp("encoder.encoder_MLP_Key.")
        self.mlp_scale = mlp("encoder.encoder_MLP_Scale.")
        self.mlp_shift = mlp("encoder.encoder_MLP_Shift.")
        # fused attention-output halves: shared projection + the private MLP that follows it
        self.pm_key = mlp("encoder.encoder_MLP_Key.", e + "proj.")
        self.pm_scale = mlp("encoder.encoder_MLP_Scale.", e + "proj.")
        self.pm_shift = mlp("encoder.encoder_MLP_Shift.", e + "proj.")
        d = "decoder.MHA_self_attn."
        self.n1 = (g(d + "norm1.weight"), g(d + "norm1.bias"))
        # decoder_exclude_MLP_after_Fcs_self_MHA=True builds the block without norm2 / mlp (reference :339-343,365)
        self.has_dec_mlp = (prefix + d + "mlp.0.weight") in sd
        self.n2 = (g(d + "norm2.weight"), g(d + "norm2.bias")) if self.has_dec_mlp else None
        a = d + "attn."
        dbq, dbk, dbv = g(a + "Wq.bias"), g(a + "Wk.bias"), g(a + "Wv.bias")
        self.dec_qkv = ops.pack_linear(torch.cat([g(a + "Wq.weight"), g(a + "Wk.weight"), g(a + "Wv.weight")], 0),
                                       torch.cat([dbq, dbk, dbv], 0))
        self.dec_pad = (dbq, dbk, dbv)
        self.dec_attn = (ops.pack_attn_qkv(g(a + "Wq.weight"), g(a + "Wk.weight"), g(a + "Wv.weight"), dbq, dbk, dbv, self.heads)
                         if _can_fuse_attn(self.C, self.heads, 8) else None)
        self.dec_proj = ops.pack_linear(g(a + "proj.weight"), g(a + "proj.bias"))
        self.dec_table = g(a + "relative_position_bias_table")
        self.dec_mlp = mlp(d + "mlp.") if self.has_dec_mlp else None
        self.pm_dec = mlp(d + "mlp.", a + "proj.") if self.has_dec_mlp else None
        # decoder_use_instance_norm_with_affine (:982-984): separate affine InstanceNorm2d modules for Query and Key
        has = lambda k: (prefix + k) in sd
        self.affine_q = (g("decoder.instance_norm_Query.weight"), g("decoder.instance_norm_Query.bias")) if has("decoder.instance_norm_Query.weight") else None
        self.affine_k = (g("decoder.instance_norm_Key.weight"), g("decoder.instance_norm_Key.bias")) if has("decoder.instance_norm_Key.weight") else None
        # decoder_use_regular_MHA_instead_of_Swin_at_the_end (:1063-1119): five plain Linears instead of the windowed sigma/mu attention
        self.regular_mha = has("decoder.linear_transformation_Key.weight")
        self.last_mlp = mlp("decoder.last_MLP.")
        if self.regular_mha:
            lin = lambda n: ops.pack_linear(g(f"decoder.{n}.weight"), g(f"decoder.{n}.bias"))
            self.rg_k, self.rg_vs, self.rg_vh = lin("linear_transformation_Key"), lin("linear_transformation_Scale"), lin("linear_transformation_Shift")
            self.rg_proj_sigma, self.rg_proj_mu = lin("proj_sigma"), lin("proj_mu")
            self.pm_last = mlp("decoder.last_MLP.", "decoder.proj_mu.")
            return
        m = "decoder.decoder_MHA_for_sigma_and_mu."
        self.sm_k = ops.pack_linear(g(m + "Wk.weight"), g(m + "Wk.bias"))
        self.sm_vs = ops.pack_linear(g(m + "Wv_scale.weight"), g(m + "Wv_scale.bias"))
        self.sm_vh = ops.pack_linear(g(m + "Wv_shift.weight"), g(m + "Wv_shift.bias"))
        self.sm_pad = (g(m + "Wk.bias"), g(m + "Wv_scale.bias"), g(m + "Wv_shift.bias"))
        self.sm_proj = ops.pack_linear(g(m + "proj.weight"), g(m + "proj.bias"))
        self.sm_table = g(m + "relative_position_bias_table")
        self.pm_last = mlp("decoder.last_MLP.", m + "proj.")


def _mlp_residual(x16, x32, fc, T, ws_: Workspace, out16):
    """x32 += fc2(gelu(fc1(x16))) in one fused kernel; optionally refresh the bf16 copy."""
    ops.mlp_fused(x16, fc, T, res=x32, out_f32=x32, out_bf16=out16)


def style_transformer_forward(w: StyleTransformerWeights, fc32: torch.Tensor, fs32: torch.Tensor, k: int, ws_: Workspace,
                              B: int, H: int, W: int, win: int, shift: int, heads: int,
                              out32: torch.Tensor, out16: Optional[torch.Tensor] = None, *, processed_key: bool = True,
                              key_in_after_linear: bool = True, exclude_mlp: bool = False, fs16_in: Optional[torch.Tensor] = None):
    """Fc, Fs fp32 [B,H,W,C] -> out32 fp32 [B,H,W,C] (+ bf16 copy for the CNN decoder).
    Follows StyleTransformer.forward (:1229-1245) -> StyleEncoder.forward (:855-882) ->
    StyleDecoder.forward (:1045-1059,1123-1128).
    Alternate orderings of the same ops:
      processed_key=False        Scale / Shift attend with the layer's INPUT Key, the Key pass runs last (:883-909);
      key_in_after_linear=False  Key is instance-normalised twice BEFORE Wk (:1057 then :470-472, on the unpadded map) and
                                 Wk.Key is used as it is -- no InstanceNorm over the padded map, padded keys = bk (:520);
      exclude_mlp=True           the decoder's self-attention block has no norm2 / MLP (:389-392)."""
    if exclude_mlp == w.has_dec_mlp:
        raise ValueError("decoder_exclude_MLP_after_Fcs_self_MHA does not match the packed state_dict (decoder.MHA_self_attn.mlp.*)")
    # Feature maps that are not a multiple of the window (the reference CLI's default 7x7 windows on 32^2 / 64^2 maps,
    # train.py:703-711) are zero-padded at the bottom/right inside the attention (style_transformer.py:77-87): a padded
    # token's projection is the bias, and the sigma/mu attention normalises Wk.K over the PADDED map (:520-530).
    padded = bool(H % win or W % win)
    Hp, Wp = -(-H // win) * win, -(-W // win) * win
    n_pad = Hp * Wp - H * W
    C = w.C
    T = B * H * W
    key32, scale32, shift32 = ws_.f32("st_key32", T, C), ws_.f32("st_scale32", T, C), ws_.f32("st_shift32", T, C)
    key16 = ws_.bf16("st_key16", T, C)
    ss16 = ws_.bf16("st_scale_shift16", 2 * T, C)  # Scale | Shift back to back: the shared Wv projects both in one GEMM
    scale16, shift16 = ss16[:T], ss16[T:]
    x32 = out32.view(T, C)
    x16 = out16.view(T, C) if out16 is not None else ws_.bf16("st_x16", T, C)
    qkv = ws_.bf16("st_qkv", T, 3 * C)
    vsh16 = ws_.bf16("st_vs_vh", 2 * T, C)
    vs16, vh16 = vsh16[:T], vsh16[T:]
    o16, o2_16 = ws_.bf16("st_o", T, C), ws_.bf16("st_o2", T, C)
    ln16 = ws_.bf16("st_ln", T, C)
    qhat16, khat16 = ws_.bf16("st_qhat", T, C), ws_.bf16("st_khat", T, C)
    kk32, sigma32 = ws_.f32("st_kk32", T, C), ws_.f32("st_sigma32", T, C)
    mean, rstd = ws_.f32("st_mean", B, C), ws_.f32("st_rstd", B, C)
    kpad = ws_.f32("st_kpad", B, C)

    fuse_attn = w.enc_attn is not None and w.dec_attn is not None and win in (7, 8) and heads == w.heads
    # Scale = Shift = Key = Fs and Query = Fc at the first layer (:1233-1236).  No copies: the first layer reads the INPUT tensors
    # (as residual sources and operands) and writes the work buffers; later layers read what the previous one wrote.  (Four fp32
    # and two bf16 copies of a [T, C] map used to open every call: ~70 us of a 2.3 ms step at batch 32.)
    fc_in, fs_in = fc32.reshape(T, C), fs32.reshape(T, C)
    if fs16_in is not None:  # the caller already holds bf16(Fs) (the Swin encoder writes it from its last MLP epilogue)
        fs16 = fs16_in.reshape(T, C)
    else:
        fs16 = ws_.bf16("st_fs16", T, C)
        ops.cast_bf16(fs_in, fs16)
    x_src, key_src, scale_src, shift_src = fc_in, fs_in, fs_in, fs_in   # fp32 residual sources of the next update of each stream
    key16_cur = fs16                                                    # bf16 Key as the attentions read it
    first = True                                                        # Scale16 = Shift16 = fs16 until the first Scale / Shift pass

    for _ in range(k):
        # ---------------- StyleEncoder: shared MHA, three private MLPs ----------------
        def key_pass():
            nonlocal key_src, key16_cur
            if fuse_attn:
                ops.attn_block(key16_cur, w.enc_attn, w.enc_table, o16, B, H, W, win, shift)
            else:
                ops.gemm(key16_cur, w.enc_qkv, T, out_bf16=qkv)
                ops.window_attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o16, w.enc_table, B, H, W, heads, win, shift, 3 * C, 3 * C, 3 * C, C,
                                     pad_q=w.enc_pad[0], pad_k=w.enc_pad[1], pad_v=w.enc_pad[2])
            if FUSE_PROJ_MLP:  # Key' = Key + proj(o); Key' += MLP_K(Key')
                ops.mlp_fused(o16, w.pm_key, T, res=key_src, out_f32=key32, out_bf16=key16, pre=True)
            else:
                ops.gemm(o16, w.enc_proj, T, res=key_src, out_f32=key32, out_bf16=key16)
                _mlp_residual(key16, key32, w.mlp_key, T, ws_, key16)
            key_src, key16_cur = key32, key16

        def scale_shift_passes():
            # q = k = Key (one softmax for both), v = Scale | Shift, residual from v
            nonlocal scale_src, shift_src, first
            ops.gemm(key16_cur, w.enc_qk, T, out_bf16=qkv, ld_out16=3 * C)
            if first:  # Scale = Shift = Fs: one projection Wv.Fs serves both value tensors (fs16 is the bf16 copy of Fs)
                ops.gemm(fs16, w.enc_v, T, out_bf16=vs16)
                v2_16 = vs16
            else:
                ops.gemm(ss16, w.enc_v, 2 * T, out_bf16=vsh16)  # v_scale = Wv.Scale, v_shift = Wv.Shift (same weight: one launch)
                v2_16 = vh16
            ops.window_attention(qkv, qkv[:, C:], vs16, o16, w.enc_table, B, H, W, heads, win, shift, 3 * C, 3 * C, C, C,
                                 v2=v2_16, out2=o2_16, pad_q=w.enc_pad[0], pad_k=w.enc_pad[1], pad_v=w.enc_pad[2], pad_v2=w.enc_pad[2])
            if FUSE_PROJ_MLP:
                ops.mlp_fused(o16, w.pm_scale, T, res=scale_src, out_f32=scale32, out_bf16=scale16, pre=True)
                ops.mlp_fused(o2_16, w.pm_shift, T, res=shift_src, out_f32=shift32, out_bf16=shift16, pre=True)
            else:
                ops.gemm(o16, w.enc_proj, T, res=scale_src, out_f32=scale32, out_bf16=scale16)
                _mlp_residual(scale16, scale32, w.mlp_scale, T, ws_, scale16)
                ops.gemm(o2_16, w.enc_proj, T, res=shift_src, out_f32=shift32, out_bf16=shift16)
                _mlp_residual(shift16, shift32, w.mlp_shift, T, ws_, shift16)
            scale_src, shift_src, first = scale32, shift32, False

        if processed_key:  # default (:857-882): Scale / Shift attend with the processed Key
            key_pass()
            scale_shift_passes()
        else:  # (:883-909): Scale / Shift attend with this layer's input Key, the Key pass runs last
            scale_shift_passes()
            key_pass()

        # ---------------- StyleDecoder ----------------
        ops.layernorm(x_src, w.n1[0], w.n1[1], ln16, T, C)
        if fuse_attn:
            ops.attn_block(ln16, w.dec_attn, w.dec_table, o16, B, H, W, win, shift)
        else:
            ops.gemm(ln16, w.dec_qkv, T, out_bf16=qkv)
            ops.window_attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o16, w.dec_table, B, H, W, heads, win, shift, 3 * C, 3 * C, 3 * C, C,
                                 pad_q=w.dec_pad[0], pad_k=w.dec_pad[1], pad_v=w.dec_pad[2])
        if exclude_mlp:  # Query = Fcs + proj(attention) only
            ops.gemm(o16, w.dec_proj, T, res=x_src, out_f32=x32)
        elif FUSE_PROJ_MLP:
            ops.mlp_fused(o16, w.pm_dec, T, res=x_src, out_f32=x32, pre=True, ln_g=w.n2[0], ln_b=w.n2[1])  # x32 = Query
        else:
            ops.gemm(o16, w.dec_proj, T, res=x_src, out_f32=x32)
            ops.layernorm(x32, w.n2[0], w.n2[1], ln16, T, C)
            _mlp_residual(ln16, x32, w.dec_mlp, T, ws_, None)  # x32 = Query
        x_src = x32
        if w.regular_mha:
            _regular_mha_tail(w, x32, key32, key16, scale16, shift16, x16, ws_, B, H * W, C, key_in_after_linear)
            continue
        # Query is instance-normalised twice (:1056 then :468); Key once before Wk and once after (:1057, :520-530).  With
        # decoder_use_instance_norm_with_affine both normalisations of a tensor go through the SAME affine module (:982-984,1052-1054)
        gq, bq_ = w.affine_q if w.affine_q is not None else (None, None)
        gk, bk_ = w.affine_k if w.affine_k is not None else (None, None)
        ops.instnorm(x32, mean, rstd, qhat16, B, H * W, C, twice=True, gamma=gq, beta=bq_)
        ops.instnorm(key32, mean, rstd, ln16, B, H * W, C, twice=not key_in_after_linear, gamma=gk, beta=bk_)
        if not key_in_after_linear:  # IN(IN(Key)) on the unpadded map, then k = Wk.Key + bk as it is
            ops.gemm(ln16, w.sm_k, T, out_bf16=khat16)
        else:
            ops.gemm(ln16, w.sm_k, T, out_f32=kk32)
            if padded:  # statistics over the padded map: its n_pad extra tokens all hold Wk.0 + bk = bk
                ops.instnorm(kk32, mean, rstd, khat16, B, H * W, C, gamma=gk, beta=bk_, n_pad=n_pad, pad_val=w.sm_pad[0], pad_norm=kpad)
            else:
                ops.instnorm(kk32, mean, rstd, khat16, B, H * W, C, gamma=gk, beta=bk_)
        ops.gemm(scale16, w.sm_vs, T, out_bf16=vs16)
        ops.gemm(shift16, w.sm_vh, T, out_bf16=vh16)
        # padded tokens: q = IN(0) (no Q projection, :511-514: zero without the affine bias), k = the normalised bias (per image),
        # v = the value biases
        ops.window_attention(qhat16, khat16, vs16, o16, w.sm_table, B, H, W, heads, win, shift, C, C, C, C, v2=vh16, out2=o2_16,
                             pad_k=(kpad if key_in_after_linear else w.sm_pad[0]) if padded else None,
                             pad_v=w.sm_pad[1], pad_v2=w.sm_pad[2], pad_k_per_image=padded and key_in_after_linear)
        ops.gemm(o16, w.sm_proj, T, out_f32=sigma32)
        if FUSE_PROJ_MLP:  # Query = Query*sigma + mu (:1123); Query += last_MLP(Query)
            ops.mlp_fused(o2_16, w.pm_last, T, res=x32, mul=sigma32, out_f32=x32, out_bf16=x16, pre=True)
        else:
            ops.gemm(o2_16, w.sm_proj, T, res=x32, mul=sigma32, out_f32=x32, out_bf16=x16)  # Query*sigma + mu (:1123)
            _mlp_residual(x16, x32, w.last_mlp, T, ws_, x16)


def _regular_mha_tail(w: StyleTransformerWeights, x32, key32, key16, scale16, shift16, x16, ws_: Workspace, B: int, Ti: int, C: int,
                      key_in_after_linear: bool):
    """decoder_use_regular_MHA_instead_of_Swin_at_the_end (codes/style_transformer.py:1063-1119): ONE head over all Ti = H*W tokens
    of an image.  q = IN(Query) * C^-0.5 with no projection, k / v_scale / v_shift through linear_transformation_*, separate
    proj_sigma / proj_mu; the reference's InstanceNorm2d reads the [B, C, T] tensors as one unbatched image, i.e. normalises over
    (C, T) jointly (mst_jointnorm_stats).  Sequenced from the tensor-core GEMM with the image's keys / values packed as the B
    operand: S = Q K^T (fp32, [Ti, Ti] per image), row softmax, O = P [V_scale | V_shift]; then sigma = proj_sigma(O_s) and
    Query*sigma + proj_mu(O_h) + last_MLP through the fused projection + MLP kernel.  x32 holds Query on entry, the layer
    output on exit."""
    T = B * Ti
    mean, rstd = ws_.f32("st_mean", B, C), ws_.f32("st_rstd", B, C)
    qhat16, khat16 = ws_.bf16("st_qhat", T, C), ws_.bf16("st_khat", T, C)
    ln16 = ws_.bf16("st_ln", T, C)
    kk32, sigma32 = ws_.f32("st_kk32", T, C), ws_.f32("st_sigma32", T, C)
    vcat = ws_.bf16("rg_vcat", T, 2 * C)   # [v_scale | v_shift] per token
    ocat = ws_.bf16("rg_ocat", T, 2 * C)   # [P v_scale | P v_shift]
    S = ws_.f32("rg_scores", Ti, Ti)
    P = ws_.bf16("rg_probs", Ti, Ti)
    NC = min(Ti, 1024)  # keys per score GEMM (mst_gemm takes N <= 1024)
    kpack = ws_.bf16("rg_kpack", NC, ops.round_up(C, 64))
    vpack = ws_.bf16("rg_vpack", ops.n_pad_of(2 * C), ops.round_up(Ti, 64))
    ops.jointnorm_stats(x32, mean, rstd, B, Ti, C)
    ops.instnorm_apply(x32, mean, rstd, B, Ti, C, y16=qhat16)
    if key_in_after_linear:
        ops.gemm(key16, w.rg_k, T, out_f32=kk32)
        ops.jointnorm_stats(kk32, mean, rstd, B, Ti, C)
        ops.instnorm_apply(kk32, mean, rstd, B, Ti, C, y16=khat16)
    else:
        ops.jointnorm_stats(key32, mean, rstd, B, Ti, C)
        ops.instnorm_apply(key32, mean, rstd, B, Ti, C, y16=ln16)
        ops.gemm(ln16, w.rg_k, T, out_bf16=khat16)
    ops.gemm(scale16, w.rg_vs, T, out_bf16=vcat, ld_out16=2 * C)
    ops.gemm(shift16, w.rg_vh, T, out_bf16=vcat[:, C:], ld_out16=2 * C)
    for b in range(B):
        rows = slice(b * Ti, (b + 1) * Ti)
        for j0 in range(0, Ti, NC):
            nc = min(NC, Ti - j0)
            pk = ops.pack_bf16_matrix(khat16[b * Ti + j0: b * Ti + j0 + nc], nc, C, C, dst=kpack)
            ops.gemm(qhat16[rows], pk, Ti, out_f32=S[:, j0:], ld_out32=Ti)
        ops.softmax_rows(S, P, Ti, Ti, float(C) ** -0.5)
        pv = ops.pack_bf16_matrix(vcat[rows], 2 * C, Ti, 2 * C, trans=True, dst=vpack)
        ops.gemm(P, pv, Ti, out_bf16=ocat[rows])
    ops.gemm(ocat, w.rg_proj_sigma, T, lda=2 * C, out_f32=sigma32)
    if FUSE_PROJ_MLP:
        ops.mlp_fused(ocat[:, C:], w.pm_last, T, lda=2 * C, res=x32, mul=sigma32, out_f32=x32, out_bf16=x16, pre=True)
    else:
        ops.gemm(ocat[:, C:], w.rg_proj_mu, T, lda=2 * C, res=x32, mul=sigma32, out_f32=x32, out_bf16=x16)
        _mlp_residual(x16, x32, w.last_mlp, T, ws_, x16)


# --------------------------------------------------------------------------------------------
# CNN decoder (codes/decoder.py:23-55)
# --------------------------------------------------------------------------------------------

CNN_LAYOUT = [  # (sequential index, upsample folded into this conv's read, relu)
    (0, False, True), (3, True, True), (5, False, True), (7, False, True), (9, False, True),
    (12, True, True), (14, False, True), (17, True, True), (19, False, False)]


class CnnDecoderWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = "decoder."):
        self.convs = []
        for idx, up, relu in CNN_LAYOUT:
            wt = _f32(sd[f"{prefix}{idx}.weight"])
            self.convs.append((ops.pack_conv3x3(wt, _f32(sd[f"{prefix}{idx}.bias"])), int(wt.shape[1]), up, relu))


def decoder_u8_supported(w: "CnnDecoderWeights", H: int, W: int) -> bool:
    """Whether cnn_decoder_forward can write the uint8 image itself (the last conv runs on the row-streaming kernel)."""
    pm, cin, _, _ = w.convs[-1]
    return ops.rows_supported(pm.N, cin, 8 * H, 8 * W)


def cnn_decoder_forward(w: CnnDecoderWeights, x16: torch.Tensor, ws_: Workspace, B: int, H: int, W: int, out: torch.Tensor):
    """x16 bf16 [B,H,W,256] token-major -> out fp32 [B,3,8H,8W] NCHW, or -- out uint8 [B,8H,8W,3], see decoder_u8_supported --
    the image test_model.py:207 saves, np.clip(out * 255, 0, 255).astype(np.uint8), straight from the last conv's epilogue."""
    cur = x16
    h, wd = H, W
    last = len(w.convs) - 1
    for i, (pm, cin, up, relu) in enumerate(w.convs):
        if up:
            if cin >= 128 and cin % 64 == 0 and not ops.cm_supported(pm.N, cin, 2 * h, 2 * wd):
                # wide layer on the gathered GEMM: materialise the nearest-x2 upsample so the conv can be fed by tensor copies (TMA
                # cannot repeat pixels); the channel-major kernel folds it into its row fetch
                big = ws_.bf16("cnn_up", B * 4 * h * wd, cin)
                ops.upsample2x_nhwc(cur, big, B, h, wd, cin)
                cur, up = big, False
            h, wd = 2 * h, 2 * wd
        M = B * h * wd
        conv = dict(H=h, W=wd, Cin=cin, pad_mode=PAD_REFLECT, upsample=up)
        if i == last and out.dtype == torch.uint8:
            conv.update(n_real=pm.N)
            ops.gemm(cur, pm, M, act=ACT_NONE, out_u8=out, conv=conv)
        elif i == last:
            conv.update(out_nchw=True, n_real=pm.N)
            ops.gemm(cur, pm, M, act=ACT_NONE, out_f32=out, conv=conv)
        else:
            nxt = ws_.bf16(f"cnn_{i % 2}", M, pm.n_pad)
            ops.gemm(cur, pm, M, act=ACT_RELU if relu else ACT_NONE, out_bf16=nxt, conv=conv)
            cur = nxt


# --------------------------------------------------------------------------------------------
# VGG-19 taps + content/style loss (codes/loss.py:15-37,71-336)
# --------------------------------------------------------------------------------------------

VGG_CONVS = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28]  # Conv2d indices inside features[:30]
VGG_POOL_BEFORE = {5, 10, 19, 28}                              # MaxPool2d(2) sits right before these convs
VGG_TAP_AFTER = {5: 0, 10: 1, 19: 2, 28: 3}                    # relu2_1, relu3_1, relu4_1, relu5_1


class VggWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = "features."):
        self.first_w = _f32(sd[prefix + "0.weight"])
        self.first_b = _f32(sd[prefix + "0.bias"])
        self.convs = {}
        for idx in VGG_CONVS[1:]:
            wt = _f32(sd[f"{prefix}{idx}.weight"])
            self.convs[idx] = (ops.pack_conv3x3(wt, _f32(sd[f"{prefix}{idx}.bias"])), int(wt.shape[1]))


def vgg_taps_forward(w: VggWeights, imgs: torch.Tensor, ws_: Workspace, tag: str):
    """imgs fp32 [N,3,H,W] NCHW -> four bf16 NHWC taps [N,h,w,C] (views into the workspace, valid until the next call
    with the same tag)."""
    N, _, H, W = imgs.shape
    cur = ws_.bf16(tag + "a0", N * H * W, 64)
    ops.conv3x3_first(imgs, w.first_w, w.first_b, cur, N, H, W, relu=True)
    h, wd, c = H, W, 64
    taps = [None] * 4
    flip = 1
    for idx in VGG_CONVS[1:]:
        pm, cin = w.convs[idx]
        if idx in VGG_POOL_BEFORE:
            pooled = ws_.bf16(tag + f"p{idx}", N * (h // 2) * (wd // 2), c)
            ops.maxpool2x2(cur, pooled, N, h, wd, c)
            cur, h, wd = pooled, h // 2, wd // 2
        name = tag + (f"tap{VGG_TAP_AFTER[idx]}" if idx in VGG_TAP_AFTER else f"a{flip}")
        out = ws_.bf16(name, N * h * wd, pm.n_pad)
        ops.gemm(cur, pm, N * h * wd, act=ACT_RELU, out_bf16=out, conv=dict(H=h, W=wd, Cin=cin, pad_mode=0, upsample=False))
        cur, c = out, pm.n_pad
        if idx in VGG_TAP_AFTER:
            taps[VGG_TAP_AFTER[idx]] = (out, h, wd, c)
        else:
            flip ^= 1
    return taps


# ---- VGG-19-BN variant (use_vgg19_with_batchnorm, codes/loss.py:41-63; torchvision vgg19_bn.features[:43]) ----
VGG_BN_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512]  # up to relu5_1 (index 42)
VGG_BN_TAP_AFTER = {9: 0, 16: 1, 29: 2, 42: 3}  # the ReLU that closes features[:10], [10:17], [17:30], [30:43]


class VggBnWeights:
    """features.{i}.weight/bias of the convolutions, features.{i+1}.{weight,bias,running_mean,running_var} of their BatchNorm2d."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = "features."):
        self.layers = []  # ("pool",) | ("conv", packed / first-conv weights, cin, bn tensors, index of the ReLU)
        i = 0
        for v in VGG_BN_CFG:
            if v == "M":
                self.layers.append(("pool",))
                i += 1
                continue
            wt, bias = _f32(sd[f"{prefix}{i}.weight"]), _f32(sd[f"{prefix}{i}.bias"])
            bn = tuple(_f32(sd[f"{prefix}{i + 1}.{k}"]) for k in ("weight", "bias", "running_mean", "running_var"))
            conv = (wt, bias) if i == 0 else ops.pack_conv3x3(wt, bias)
            self.layers.append(("conv", conv, int(wt.shape[1]), bn, i + 2))
            i += 3


def vgg_bn_taps_forward(w: VggBnWeights, imgs: torch.Tensor, ws_: Workspace, tag: str, training: bool, tap_out=None, tap_row0: int = 0,
                        eps: float = 1e-5):
    """imgs fp32 [N,3,H,W] -> the four bf16 NHWC taps.  Every convolution is followed by BatchNorm2d -- in train mode with the
    statistics of THIS batch over (N,H,W) (biased variance; mst_tap_stats with B = 1), in eval mode with the running ones -- and
    ReLU, applied in place by one elementwise kernel.  tap_out: optional list of four [rows, C] buffers; the taps are written
    there starting at row tap_row0 * (tap tokens per image) (the loss stacks content | style | output taps)."""
    N, _, H, W = imgs.shape
    h, wd, c = H, W, 3
    cur = None
    taps = [None] * 4
    mean, var = ws_.f32(tag + "bn_mean", 512), ws_.f32(tag + "bn_var", 512)
    flip = 0
    for layer in w.layers:
        if layer[0] == "pool":
            pooled = ws_.bf16(tag + f"p{flip}", N * (h // 2) * (wd // 2), c)
            ops.maxpool2x2(cur, pooled, N, h, wd, c)
            cur, h, wd = pooled, h // 2, wd // 2
            continue
        _, conv, cin, (g, b, rm, rv), relu_idx = layer
        M = N * h * wd
        slot = VGG_BN_TAP_AFTER.get(relu_idx)
        pre = None
        if isinstance(conv, tuple):  # first convolution: fp32 NCHW image in, 64 channels out (bf16; normalised in place)
            cout = 64
            out = ws_.bf16(tag + "a0", M, cout)
            ops.conv3x3_first(imgs, conv[0], conv[1], out, N, h, wd, relu=False)
        else:
            cout = conv.n_pad
            if slot is not None and tap_out is not None:
                out = tap_out[slot][tap_row0 * h * wd: tap_row0 * h * wd + M]
            else:
                out = ws_.bf16(tag + (f"tap{slot}" if slot is not None else f"a{1 + flip}"), M, cout)
            # the convolution's output stays fp32 until it is normalised: a channel with |mean| >> std must not meet bf16 first
            pre = ws_.f32(tag + "pre", M, cout)
            ops.gemm(cur, conv, M, act=ACT_NONE, out_f32=pre, conv=dict(H=h, W=wd, Cin=cin, pad_mode=0, upsample=False))
        if not training:
            ops.bn_relu(out, rm, rv, g, b, eps, M, cout, x32=pre)
        elif pre is None:
            ops.tap_stats(out, mean[:cout].view(1, cout), var[:cout].view(1, cout), 1, M, cout)
            ops.bn_relu(out, mean[:cout], var[:cout], g, b, eps, M, cout)
        else:  # batch statistics over (N, H, W): the InstanceNorm statistics kernel on the batch read as ONE image (eps = 1e-5 = BN's)
            ops.instnorm_stats(pre, mean[:cout].view(1, cout), var[:cout].view(1, cout), 1, M, cout)
            ops.bn_relu(out, mean[:cout], var[:cout], g, b, eps, M, cout, x32=pre, var_is_rstd=True)
        cur, c = out, cout
        flip ^= 1
        if slot is not None:
            taps[slot] = (out, h, wd, cout)
    return taps


def perceptual_loss_forward_bn(w: VggBnWeights, content, style, output, lam: float, squared_content: bool, squared_style: bool,
                               ws_: Workspace, training: bool) -> torch.Tensor:
    """get_overall_loss (loss.py:201-262) with the VGG-19-BN extractor: the content, style and output batches go through the
    network in three separate passes (loss.py:223-225), each normalised with its own batch statistics in train mode."""
    B, _, H, W = content.shape
    shapes = [(H // 2, W // 2, 128), (H // 4, W // 4, 256), (H // 8, W // 8, 512), (H // 16, W // 16, 512)]
    stacked = [ws_.bf16(f"vggbn_tap{i}", 3 * B * h * wd, c) for i, (h, wd, c) in enumerate(shapes)]
    for k, img in enumerate((content, style, output)):
        vgg_bn_taps_forward(w, img.float().contiguous(), ws_, "vggbn_", training, tap_out=stacked, tap_row0=k * B)
    descs = []
    for i, (h, wd, c) in enumerate(shapes):
        T = h * wd
        mean, var = ws_.f32(f"loss_mean{i}", 3 * B, c), ws_.f32(f"loss_var{i}", 3 * B, c)
        ops.tap_stats(stacked[i], mean, var, 3 * B, T, c)
        partials = ws_.f32(f"loss_part{i}", 592)
        tv = stacked[i].view(3 * B, T * c)
        ops.content_term(tv[:B], tv[2 * B:], mean[:B], var[:B], mean[2 * B:], var[2 * B:], B, T, c, squared_content, partials)
        descs.append(dict(partials=partials, mean_s=mean[B:2 * B], var_s=var[B:2 * B], mean_o=mean[2 * B:], var_o=var[2 * B:], B=B, T=T, C=c))
    out3 = torch.empty(3, dtype=torch.float32, device=content.device)
    ops.loss_finalize(descs, lam, squared_style, out3)
    return out3


def similarity_loss_forward(taps, B: int, squared: bool, ws_: Workspace) -> torch.Tensor:
    """The paper's similarity loss between the CONTENT and the OUTPUT image's relu3_1 / relu4_1 taps (codes/loss.py:137-146,
    321-336 with the arguments the paper means; codes/utils.py:105-133): column-normalised cosine self-similarity maps, strict
    lower triangle, mean |difference| (or squared) over the whole B x N x N map, summed over the two taps.  taps: the list
    vgg_taps_forward returns for the stacked (content | style | output) batch.  Returns a device fp32 scalar tensor."""
    parts = []
    for i in (1, 2):
        t, h, wd, c = taps[i]
        N = h * wd
        tiles = ops.sim_num_tiles(B, N)
        rows = t.view(3 * B * N, c)
        ah_c, ah_o = ws_.bf16(f"sim_ahat_c{i}", B * N, c), ws_.bf16(f"sim_ahat_o{i}", B * N, c)
        sv = ws_.f32(f"sim_svec{i}", 2, B, c)
        inv = ws_.f32(f"sim_inv{i}", 2, B, N)
        part = ws_.f32(f"sim_part{i}", tiles)
        ops.sim_prepare(rows[:B * N], B, N, c, ah_c, sv[0], inv[0])
        ops.sim_prepare(rows[2 * B * N:], B, N, c, ah_o, sv[1], inv[1])
        ops.sim_tiles(ah_c, inv[0], ah_o, inv[1], B, N, c, squared, part)
        parts.append((part, float(B) * N * N))
    out = torch.empty(1, dtype=torch.float32, device=taps[1][0].device)
    ops.sim_finalize(parts[0][0], parts[0][1], parts[1][0], parts[1][1], out)
    return out[0]


def perceptual_loss_forward(w: VggWeights, content: torch.Tensor, style: torch.Tensor, output: torch.Tensor, lam: float,
                            squared_content: bool, squared_style: bool, ws_: Workspace, similarity: bool = False):
    """Returns a device fp32 tensor [3] = (total, content, style) following get_overall_loss (loss.py:201-262); with
    similarity=True a pair (that tensor, the content-vs-output similarity loss scalar)."""
    B = int(content.shape[0])
    imgs = ws_.f32("loss_imgs", 3 * B, 3, content.shape[2], content.shape[3])
    imgs[:B].copy_(content)
    imgs[B:2 * B].copy_(style)
    imgs[2 * B:].copy_(output)
    taps = vgg_taps_forward(w, imgs, ws_, "vgg_")
    descs = []
    for i, (t, h, wd, c) in enumerate(taps):
        T = h * wd
        mean, var = ws_.f32(f"loss_mean{i}", 3 * B, c), ws_.f32(f"loss_var{i}", 3 * B, c)
        ops.tap_stats(t, mean, var, 3 * B, T, c)
        partials = ws_.f32(f"loss_part{i}", 592)
        tv = t.view(3 * B, T * c)
        ops.content_term(tv[:B], tv[2 * B:], mean[:B], var[:B], mean[2 * B:], var[2 * B:], B, T, c, squared_content, partials)
        descs.append(dict(partials=partials, mean_s=mean[B:2 * B], var_s=var[B:2 * B], mean_o=mean[2 * B:], var_o=var[2 * B:], B=B, T=T, C=c))
    out3 = torch.empty(3, dtype=torch.float32, device=content.device)
    ops.loss_finalize(descs, lam, squared_style, out3)
    if similarity:
        return out3, similarity_loss_forward(taps, B, squared_style, ws_)
    return out3
