"""CPU stand-ins for the C-ABI kernels the inference style transformer launches -- TEST INFRASTRUCTURE ONLY.

`install(monkeypatch)` replaces the wrappers in `mastermetastyletransfer_b200.ops` that the inference engine (Swin encoder,
style transformer, CNN decoder, VGG-19 loss: `engine.swin_encode`, `engine.style_transformer_forward`,
`engine.cnn_decoder_forward`, `engine.perceptual_loss_forward` and their weight holders) calls with torch-CPU restatements of each kernel's documented contract
(include/mst_b200.h): same arguments, same in-place buffer semantics, bf16 tensors really stored as bf16 (so the rounding
points of the device path are reproduced: bf16 operands and weights, fp32 accumulation and residual streams).  What this
checks is the HOST logic -- which kernel runs on which buffer in which order for each configuration -- on a machine
without a GPU; the kernels themselves are checked on the B200 (`-m gpu`).  Nothing in the product imports this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from mastermetastyletransfer_b200 import ops
from oracle import master_oracle as O


class _Mlp:
    def __init__(self, w1, b1, w2, b2, wpre=None, bpre=None):
        r = lambda w: None if w is None else w.detach().bfloat16().float()
        self.w1, self.w2, self.wpre = r(w1), w2.detach().half().float(), r(wpre)  # fc2 runs in fp16 (hidden activation and W2)
        self.b1, self.b2 = b1.detach().float(), b2.detach().float()
        self.bpre = None if bpre is None else bpre.detach().float()
        self.C = w1.shape[1]


def pack_linear(weight, bias=None):
    N, K = weight.shape
    return ops.PackedMatrix(weight.detach().bfloat16(), None if bias is None else bias.detach().float(), N, K, N, K)


def pack_bf16_matrix(src, N, K, ld, trans=False, dst=None):
    m = _rows(src, K, N, ld).t() if trans else _rows(src, N, K, ld)
    return ops.PackedMatrix(m.clone(), None, N, K, ops.n_pad_of(N), ops.round_up(K, 64))


def softmax_rows(S, P, rows, n, scale):
    _rows(P, rows, n, n).copy_(torch.softmax(_rows(S, rows, n, n) * scale, dim=-1))


def pack_mlp(w1, b1, w2, b2, wpre=None, bpre=None):
    return _Mlp(w1, b1, w2, b2, wpre, bpre)


def _rows(t, M, N, ld):
    """[M, N] view with row stride ld of a token-major buffer."""
    return t.as_strided((M, N), (ld, 1), t.storage_offset())


def cast_bf16(x, y):
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16
    y.copy_(x)


def pack_conv3x3(weight, bias=None):
    N, Cin = weight.shape[:2]
    n_pad = ops.n_pad_of(N)
    return ops.PackedMatrix(weight.detach().bfloat16(), None if bias is None else bias.detach().float(), N, 9 * Cin, n_pad, 9 * Cin)


def _conv3x3(A, pm, M, conv, act, out_f32, out_bf16):
    """3x3 convolution on a bf16 NHWC tensor: optional nearest x2 upsample folded into the read, reflect or zero padding,
    bias, ReLU; bf16 NHWC [M, n_pad] or fp32 NCHW [B, n_real, H, W] output (include/mst_b200.h, MstGemm a_mode CONV3X3)."""
    H, W, Cin, up = conv["H"], conv["W"], conv["Cin"], bool(conv.get("upsample", False))
    B = M // (H * W)
    hin, win = (H // 2, W // 2) if up else (H, W)
    x = A.reshape(-1)[: B * hin * win * Cin].view(B, hin, win, Cin).float().permute(0, 3, 1, 2)
    if up:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.pad(x, (1, 1, 1, 1), mode="reflect" if conv.get("pad_mode", ops.PAD_ZERO) == ops.PAD_REFLECT else "constant")
    y = F.conv2d(x, pm.w.float(), pm.bias)
    if act == ops.ACT_RELU:
        y = torch.relu(y)
    if conv.get("out_nchw"):
        out_f32.copy_(y[:, : conv.get("n_real", pm.N)])
    else:
        for dst in (out_bf16, out_f32):
            if dst is None:
                continue
            out = dst.reshape(-1)[: M * pm.n_pad].view(M, pm.n_pad)
            out.zero_()
            out[:, : pm.N].copy_(y.permute(0, 2, 3, 1).reshape(M, pm.N))


def upsample2x_nhwc(x, y, B, H, W, C_):
    src = x.reshape(-1)[: B * H * W * C_].view(B, H, W, C_)
    y.reshape(-1)[: B * 4 * H * W * C_].view(B, 2 * H, 2 * W, C_).copy_(src.repeat_interleave(2, 1).repeat_interleave(2, 2))


def patch_embed(img, w, b, gamma, beta, x, B, S, gamma1=None, beta1=None, y16=None, exact=False, u8_mean=None, u8_std=None):
    """tv swin features[0]: Conv2d(3, 128, 4, stride 4) -> BHWC -> LayerNorm; optional fused norm1 of the first block.
    uint8 [B,S,S,3] images: transforms.ToTensor() (+ Normalize(u8_mean, u8_std)) first."""
    P = S // 4
    if img.dtype == torch.uint8:
        img = img.permute(0, 3, 1, 2).float().div(255)
        if u8_mean is not None:
            img = (img - torch.tensor(u8_mean).view(1, 3, 1, 1)) / torch.tensor(u8_std).view(1, 3, 1, 1)
    t = F.layer_norm(F.conv2d(img[:B], w, b, stride=4).permute(0, 2, 3, 1), (128,), gamma, beta).reshape(B * P * P, 128)
    x[: B * P * P].copy_(t)
    if y16 is not None:
        y16[: B * P * P].copy_(F.layer_norm(t, (128,), gamma1, beta1))


def patch_merge_layernorm(x, gamma, beta, y, B, H, W, Cdim):
    """tv PatchMerging (swin_transformer.py:35-87) up to the reduction: 2x2 neighbourhood concat [x00, x10, x01, x11] -> LN(4C)."""
    t = x[: B * H * W].view(B, H, W, Cdim)
    cat = torch.cat([t[:, 0::2, 0::2], t[:, 1::2, 0::2], t[:, 0::2, 1::2], t[:, 1::2, 1::2]], -1)
    y[: B * (H // 2) * (W // 2)].copy_(F.layer_norm(cat, (4 * Cdim,), gamma, beta).reshape(-1, 4 * Cdim))


def gemm(A, pm, M, *, lda=None, act=ops.ACT_NONE, res=None, mul=None, out_f32=None, out_bf16=None, ld_out32=None, ld_out16=None,
         ld_res=None, conv=None, **kw):
    assert not kw, kw
    if conv is not None:
        assert res is None and mul is None
        return _conv3x3(A, pm, M, conv, act, out_f32, out_bf16)
    assert act == ops.ACT_NONE
    assert A.dtype == torch.bfloat16 and (out_f32 is not None or out_bf16 is not None)
    N, K = pm.N, pm.K
    x = _rows(A, M, K, K if lda is None else lda).float() @ pm.w.float().t()
    if pm.bias is not None:
        x = x + pm.bias
    if res is not None:
        r = _rows(res, M, N, N if ld_res is None else ld_res)
        x = r * _rows(mul, M, N, N if ld_res is None else ld_res) + x if mul is not None else r + x
    else:
        assert mul is None
    if out_f32 is not None:
        _rows(out_f32, M, N, N if ld_out32 is None else ld_out32).copy_(x)
    if out_bf16 is not None:
        _rows(out_bf16, M, N, N if ld_out16 is None else ld_out16).copy_(x)


def mlp_next_ln_supported(C_):
    return C_ in (128, 256)


def mlp_fused(A, pm, M, *, lda=None, res=None, out_f32=None, out_bf16=None, pre=False, mul=None, ln_g=None, ln_b=None, next_ln=None):
    C = pm.C
    assert next_ln is None or (pre and out_bf16 is not None)
    a = _rows(A, M, C, C if lda is None else lda).float()
    if pre:
        assert pm.wpre is not None
        x1 = a @ pm.wpre.t() + pm.bpre
        x1 = res[:M] * mul[:M] + x1 if mul is not None else res[:M] + x1
        xin = F.layer_norm(x1, (C,), ln_g, ln_b) if ln_g is not None else x1
        xin = xin.bfloat16().float()
    else:
        assert pm.wpre is None and mul is None and ln_g is None
        x1 = res[:M] if res is not None else 0.0
        xin = a
    h = F.gelu(xin @ pm.w1.t() + pm.b1).half().float()
    y = x1 + h @ pm.w2.t() + pm.b2
    if out_f32 is not None:
        out_f32[:M].copy_(y)
    if out_bf16 is not None:
        out_bf16[:M].copy_(y)
        if next_ln is not None:
            rows = next_ln[2] if len(next_ln) > 2 and next_ln[2] else M
            out_bf16[:rows].copy_(F.layer_norm(y[:rows], (C,), next_ln[0], next_ln[1]))


def layernorm(x, gamma, beta, y, rows, Cdim):
    y[:rows].copy_(F.layer_norm(x[:rows], (Cdim,), gamma, beta))


def _stats(x, B, T, Cdim, extra=None, n_extra=0):
    """biased mean / variance per (b, c) over T tokens (+ n_extra tokens of value extra[c])."""
    xb = x.reshape(B, T, Cdim).double()
    n = T + n_extra
    s1, s2 = xb.sum(1), (xb * xb).sum(1)
    if n_extra:
        s1, s2 = s1 + n_extra * extra.double(), s2 + n_extra * extra.double() ** 2
    mean = s1 / n
    return mean, s2 / n - mean * mean


def instnorm_stats(x, mean, rstd, B, T, Cdim, twice=False, gamma=None):
    m, var = _stats(x, B, T, Cdim)
    r = 1.0 / torch.sqrt(var + 1e-5)
    g = 1.0 if gamma is None else gamma.double().unsqueeze(0)
    if twice:  # IN(IN(x)) with the same (affine) module: the once-normalised map has variance g^2 var/(var+eps)
        r = r * g * g / torch.sqrt(g * g * var * r * r + 1e-5)
    else:
        r = r * g
    mean.copy_(m)
    rstd.copy_(r)


def instnorm_stats_padded(x, mean, rstd, B, T, Cdim, n_pad, pad_val, pad_norm=None, gamma=None, beta=None):
    m, var = _stats(x, B, T, Cdim, pad_val, n_pad)
    r = 1.0 / torch.sqrt(var + 1e-5)
    if gamma is not None:
        r = r * gamma.double().unsqueeze(0)
    mean.copy_(m)
    rstd.copy_(r)
    if pad_norm is not None:
        pad_norm.copy_((pad_val.double().unsqueeze(0) - m) * r + (0.0 if beta is None else beta.double().unsqueeze(0)))


def jointnorm_stats(x, mean, rstd, B, T, Cdim):
    xd = x.reshape(B, T * Cdim).double()
    m = xd.mean(1, keepdim=True)
    var = xd.var(1, unbiased=False, keepdim=True)
    mean.copy_(m.expand(B, Cdim))
    rstd.copy_((1.0 / torch.sqrt(var + 1e-5)).expand(B, Cdim))


def instnorm_apply(x, mean, rstd, B, T, Cdim, y16=None, y32=None, beta=None):
    y = (x.reshape(B, T, Cdim) - mean.unsqueeze(1)) * rstd.unsqueeze(1)
    if beta is not None:
        y = y + beta.reshape(1, 1, Cdim)
    for dst in (y16, y32):
        if dst is not None:
            dst.reshape(B, T, Cdim).copy_(y)


def instnorm(x, mean, rstd, y16, B, T, Cdim, twice=False, gamma=None, beta=None, n_pad=0, pad_val=None, pad_norm=None):
    """statistics + application in one call (ops.instnorm)"""
    if n_pad > 0 or pad_norm is not None:
        assert not twice
        instnorm_stats_padded(x, mean, rstd, B, T, Cdim, n_pad, pad_val, pad_norm=pad_norm, gamma=gamma, beta=beta)
    else:
        instnorm_stats(x, mean, rstd, B, T, Cdim, twice=twice, gamma=gamma)
    instnorm_apply(x, mean, rstd, B, T, Cdim, y16=y16, beta=beta)


def window_attention(q, k, v, out, bias_table, B, H, W, heads, ws, shift, ldq, ldk, ldv, ldo,
                     v2=None, out2=None, pad_q=None, pad_k=None, pad_v=None, pad_v2=None, pad_k_per_image=False):
    C = heads * 32
    T = B * H * W
    Hp, Wp = O.padded_dims(H, W, ws)

    def padded_map(t, ld, pad, per_image=False):
        m = torch.zeros(B, Hp, Wp, C)
        if pad is not None:
            m += pad.reshape(B, 1, 1, C) if per_image else pad.reshape(1, 1, 1, C)
        m[:, :H, :W] = _rows(t, T, C, ld).float().reshape(B, H, W, C)
        return m

    qw = O._to_windows(padded_map(q, ldq, pad_q), ws, shift)
    kw = O._to_windows(padded_map(k, ldk, pad_k, pad_k_per_image), ws, shift)
    p = O._softmax_probs(qw, kw, heads, O._bias_from_table(bias_table, ws), O.shift_mask(Hp, Wp, ws, shift), B)
    for val, pad, dst in ((v, pad_v, out), (v2, pad_v2, out2)):
        if val is None:
            continue
        o = O._apply_probs(p, O._to_windows(padded_map(val, ldv, pad), ws, shift), heads)
        _rows(dst, T, C, ldo).copy_(O._from_windows(o, B, Hp, Wp, ws, shift)[:, :H, :W].reshape(T, C))


class _AttnQkv:
    def __init__(self, ws_, bs_, C_, heads):
        self.w3, self.b3, self.C, self.heads = ws_, bs_, C_, heads


def pack_attn_qkv(wq, wk, wv, bq, bk, bv, heads):
    """bf16-rounded weights, fp32 biases (what mst_pack_attn_qkv stores)."""
    return _AttnQkv([t.detach().bfloat16().float() for t in (wq, wk, wv)], [t.detach().float() for t in (bq, bk, bv)], int(wq.shape[0]), heads)


def attn_block(x16, pk, bias_table, out, B, H, W, ws, shift, ldx=None, ldo=None, dbg_qkv=None):
    """Contract of mst_attn_block: q, k, v = bf16(x W^T + b) on the zero-padded token map (a padded token's projection is the
    bias), then the attention core of window_attention()."""
    C, T = pk.C, B * H * W
    Hp, Wp = O.padded_dims(H, W, ws)
    xm = torch.zeros(B, Hp, Wp, C)
    xm[:, :H, :W] = _rows(x16, T, C, ldx or C).float().reshape(B, H, W, C)
    xw = O._to_windows(xm, ws, shift)
    q, k, v = (F.linear(xw, w_, b_).bfloat16().float() for w_, b_ in zip(pk.w3, pk.b3))
    p = O._softmax_probs(q, k, pk.heads, O._bias_from_table(bias_table, ws), O.shift_mask(Hp, Wp, ws, shift), B)
    o = O._apply_probs(p, v, pk.heads)
    _rows(out, T, C, ldo or C).copy_(O._from_windows(o, B, Hp, Wp, ws, shift)[:, :H, :W].reshape(T, C))


def conv3x3_first(img, w, b, out, B, H, W, relu=True):
    y = F.conv2d(img[:B], w, b, padding=1)
    if relu:
        y = torch.relu(y)
    out.reshape(-1)[: B * H * W * 64].view(B, H, W, 64).copy_(y.permute(0, 2, 3, 1))


def maxpool2x2(x, y, B, H, W, Cdim):
    src = x.reshape(-1)[: B * H * W * Cdim].view(B, H, W, Cdim).float().permute(0, 3, 1, 2)
    y.reshape(-1)[: B * (H // 2) * (W // 2) * Cdim].view(B, H // 2, W // 2, Cdim).copy_(F.max_pool2d(src, 2).permute(0, 2, 3, 1))


def bn_relu(y, mean, var, gamma, beta, eps, M, Cdim, relu=True, x32=None, var_is_rstd=False):
    v = y.reshape(-1)[: M * Cdim].view(M, Cdim)
    src = v.float() if x32 is None else x32.reshape(-1)[: M * Cdim].view(M, Cdim)
    inv = var if var_is_rstd else 1.0 / torch.sqrt(var + eps)
    out = (src - mean) * (gamma * inv) + beta
    v.copy_(torch.relu(out) if relu else out)


def tap_stats(x, mean, var, B, T, Cdim, scratch=None):
    m, v = _stats(x.reshape(-1)[: B * T * Cdim].float(), B, T, Cdim)
    mean.copy_(m)
    var.copy_(v)  # biased


def content_term(fc, fo, mean_c, var_c, mean_o, var_o, B, T, Cdim, squared, partials):
    norm = lambda f, m, v: (f.reshape(B, T, Cdim).float() - m.unsqueeze(1)) / torch.sqrt(v.unsqueeze(1) + 1e-5)
    d = norm(fc, mean_c, var_c) - norm(fo, mean_o, var_o)
    partials.zero_()
    partials[0] = (d * d if squared else d.abs()).double().sum()


def loss_finalize(taps, lam, squared_style, out3):
    content = style = 0.0
    for t in taps:
        B, T, Cd = t["B"], t["T"], t["C"]
        content = content + t["partials"].double().sum() / (B * T * Cd)
        std = lambda v: torch.sqrt(v.double() * T / (T - 1))  # torch.std is unbiased
        dm, ds = t["mean_s"].double() - t["mean_o"].double(), std(t["var_s"]) - std(t["var_o"])
        style = style + ((dm * dm).mean() + (ds * ds).mean() if squared_style else dm.abs().mean() + ds.abs().mean())
    out3.copy_(torch.stack([content + lam * style, content, style]).float())


def install(monkeypatch):
    for name in ("pack_linear", "pack_mlp", "pack_conv3x3", "cast_bf16", "gemm", "mlp_fused", "layernorm", "instnorm_stats",
                 "instnorm_stats_padded", "jointnorm_stats", "pack_bf16_matrix", "softmax_rows", "instnorm_apply", "instnorm", "window_attention", "pack_attn_qkv", "attn_block", "upsample2x_nhwc", "patch_embed",
                 "patch_merge_layernorm", "conv3x3_first", "maxpool2x2", "bn_relu", "tap_stats", "content_term", "loss_finalize"):
        monkeypatch.setattr(ops, name, globals()[name])
