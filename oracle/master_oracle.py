"""CPU oracle for the Master stylization hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch fp32 restatement (torch CPU tensors, explicit gather maps) of
the reference algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package never does.

Pinned: oracle/make_golden.py imports the real reference from /root/reference in the build
container, runs both on the same seeded weights/inputs, asserts agreement (<= 2e-5 max-abs on
O(1) activations, bit-exact on integer maps) and writes tests/golden/*.npz, which
tests/test_oracle_golden.py re-checks wherever the repo travels.  The reference itself has
no tests or golden vectors, so this pinning against the reference's own
outputs is the only anchor there is.

Every function cites the reference file:line it follows (paths relative to the reference root;
"tv:" = torchvision 0.26 models/swin_transformer.py, the reference's un-vendored dependency).
All tensors are token-major BHWC unless noted.
"""
from __future__ import annotations

import functools
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# ----------------------------------------------------------------------------------------------
# integer maps (bit-exact contract)
# ----------------------------------------------------------------------------------------------


@functools.lru_cache(maxsize=None)
def relative_position_index(ws: int) -> Tensor:
    """codes/style_transformer.py:227-239: idx[i,j]=(yi-yj+ws-1)*(2ws-1)+(xi-xj+ws-1), flat [ws^4] int64."""
    n = ws * ws
    out = torch.empty(n * n, dtype=torch.int64)
    for i in range(n):
        yi, xi = divmod(i, ws)
        for j in range(n):
            yj, xj = divmod(j, ws)
            out[i * n + j] = (yi - yj + ws - 1) * (2 * ws - 1) + (xi - xj + ws - 1)
    return out


def padded_dims(H: int, W: int, ws: int) -> Tuple[int, int]:
    """codes/style_transformer.py:77-87: pad bottom/right up to a multiple of the window."""
    return H + (ws - H % ws) % ws, W + (ws - W % ws) % ws


def effective_shift(Hp: int, Wp: int, ws: int, shift: int) -> Tuple[int, int]:
    """codes/style_transformer.py:89-94: no shift along an axis the window already covers."""
    return (0 if ws >= Hp else shift), (0 if ws >= Wp else shift)


@functools.lru_cache(maxsize=None)
def window_gather_map(H: int, W: int, ws: int, shift: int) -> Tensor:
    """Source position of every window slot.

    Returns int64 [nW, ws*ws]: flat index y*Wp+x into the *padded, un-rolled* grid (values
    with y>=H or x>=W are zero padding), following pad -> roll(-s) -> partition
    (codes/style_transformer.py:83-111).  rolled[y,x] = padded[(y+sy)%Hp,(x+sx)%Wp].
    """
    Hp, Wp = padded_dims(H, W, ws)
    sy, sx = effective_shift(Hp, Wp, ws, shift)
    nwx = Wp // ws
    out = torch.empty((Hp // ws) * nwx, ws * ws, dtype=torch.int64)
    for wy in range(Hp // ws):
        for wx in range(nwx):
            for iy in range(ws):
                for ix in range(ws):
                    y = (wy * ws + iy + sy) % Hp
                    x = (wx * ws + ix + sx) % Wp
                    out[wy * nwx + wx, iy * ws + ix] = y * Wp + x
    return out


@functools.lru_cache(maxsize=None)
def region_labels(H: int, W: int, ws: int, shift: int) -> Optional[Tensor]:
    """9-region labels of each window slot on the rolled grid (codes/style_transformer.py:134-145).

    Returns int64 [nW, ws*ws] or None when no shift is active (then no mask is added, :134).
    """
    Hp, Wp = padded_dims(H, W, ws)
    sy, sx = effective_shift(Hp, Wp, ws, shift)
    if sy + sx == 0:
        return None

    def band(p: int, size: int, s: int) -> int:
        # slices (0,-ws), (-ws,-s), (-s,None); with s == 0 the last two are empty / whole-tail
        # exactly as python slicing makes them in the reference loop (later writes win).
        if s == 0:
            return 2  # slice(-ws, -0) is empty and slice(-0, None) is the whole axis: written last
        if p >= size - s:
            return 2
        return 1 if p >= size - ws else 0

    nwx = Wp // ws
    out = torch.empty((Hp // ws) * nwx, ws * ws, dtype=torch.int64)
    for wy in range(Hp // ws):
        for wx in range(nwx):
            for iy in range(ws):
                for ix in range(ws):
                    out[wy * nwx + wx, iy * ws + ix] = 3 * band(wy * ws + iy, Hp, sy) + band(wx * ws + ix, Wp, sx)
    return out


@functools.lru_cache(maxsize=None)
def shift_mask(H: int, W: int, ws: int, shift: int) -> Optional[Tensor]:
    """float32 [nW, N, N] in {0,-100}: label[j]-label[i] != 0 -> -100 (codes/style_transformer.py:146-147)."""
    lab = region_labels(H, W, ws, shift)
    if lab is None:
        return None
    diff = lab.unsqueeze(1) - lab.unsqueeze(2)
    return torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------


def _to_windows(x: Tensor, ws: int, shift: int) -> Tensor:
    """[B,H,W,C] -> [B*nW, N, C] via the gather map (zeros where the map points at padding)."""
    B, H, W, C = x.shape
    Hp, Wp = padded_dims(H, W, ws)
    gm = window_gather_map(H, W, ws, shift).to(x.device)  # (the maps are built on the CPU; bench.py's GPU-eager leg runs the same ops on cuda)
    xp = torch.zeros(B, Hp, Wp, C, dtype=x.dtype, device=x.device)
    xp[:, :H, :W] = x
    return xp.reshape(B, Hp * Wp, C)[:, gm.reshape(-1)].reshape(B * gm.shape[0], gm.shape[1], C)


def _from_windows(xw: Tensor, B: int, H: int, W: int, ws: int, shift: int) -> Tensor:
    """Inverse of _to_windows followed by the un-pad (codes/style_transformer.py:160-168)."""
    Hp, Wp = padded_dims(H, W, ws)
    gm = window_gather_map(H, W, ws, shift).reshape(-1).to(xw.device)
    C = xw.shape[-1]
    out = torch.empty(B, Hp * Wp, C, dtype=xw.dtype, device=xw.device)
    out[:, gm] = xw.reshape(B, -1, C)
    return out.reshape(B, Hp, Wp, C)[:, :H, :W].contiguous()


def _bias_from_table(table: Tensor, ws: int) -> Tensor:
    """codes/style_transformer.py:21-28: [heads, N, N]."""
    n = ws * ws
    return table[relative_position_index(ws).to(table.device)].reshape(n, n, -1).permute(2, 0, 1)


def _softmax_probs(q: Tensor, k: Tensor, heads: int, bias: Tensor, mask: Optional[Tensor], B: int) -> Tensor:
    """q,k [B*nW,N,C] -> P [B*nW, heads, N, N] (codes/style_transformer.py:120-152)."""
    bw, n, C = q.shape
    d = C // heads
    qh = q.reshape(bw, n, heads, d).permute(0, 2, 1, 3) * (d ** -0.5)
    kh = k.reshape(bw, n, heads, d).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-2, -1) + bias.unsqueeze(0)
    if mask is not None:
        mask = mask.to(s.device)
        nW = mask.shape[0]
        s = (s.reshape(B, nW, heads, n, n) + mask.reshape(1, nW, 1, n, n)).reshape(bw, heads, n, n)
    return torch.softmax(s, dim=-1)


def _apply_probs(p: Tensor, v: Tensor, heads: int) -> Tensor:
    bw, n, C = v.shape
    vh = v.reshape(bw, n, heads, C // heads).permute(0, 2, 1, 3)
    return (p @ vh).transpose(1, 2).reshape(bw, n, C)


def window_attention(xq: Tensor, xk: Tensor, xv: Tensor, wq, bq, wk, bk, wv, bv, wp, bp,
                     table: Tensor, ws: int, shift: int, heads: int) -> Tensor:
    """codes/style_transformer.py:37-169 (a3): three possibly different inputs, split Q/K/V weights."""
    B, H, W, C = xq.shape
    q = F.linear(_to_windows(xq, ws, shift), wq, bq)
    k = F.linear(_to_windows(xk, ws, shift), wk, bk)
    v = F.linear(_to_windows(xv, ws, shift), wv, bv)
    p = _softmax_probs(q, k, heads, _bias_from_table(table, ws), shift_mask(H, W, ws, shift), B)
    o = F.linear(_apply_probs(p, v, heads), wp, bp)
    return _from_windows(o, B, H, W, ws, shift)


def instance_norm_bhwc(x: Tensor, eps: float = 1e-5, affine: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
    """nn.InstanceNorm2d(C) on a BHWC tensor (codes/style_transformer.py:986,1056-1057); affine = (weight, bias) of the
    decoder_use_instance_norm_with_affine variant (:982-984)."""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = x.var(dim=(1, 2), unbiased=False, keepdim=True)
    y = (x - mean) / torch.sqrt(var + eps)
    return y if affine is None else y * affine[0] + affine[1]


def sigma_mu_attention(xq: Tensor, xk: Tensor, xvs: Tensor, xvh: Tensor, wk, bk, wvs, bvs, wvh, bvh, wp, bp,
                       table: Tensor, ws: int, shift: int, heads: int, key_in_after_linear: bool = True,
                       affine_q=None, affine_k=None) -> Tuple[Tensor, Tensor]:
    """codes/style_transformer.py:414-611 (a8), default flags: IN(q) again (:468), no Q projection
    (:511-514), IN over the whole (padded) map of Wk*K (:520-530), one softmax for both values,
    the same proj for sigma and mu (:575-607).
    key_in_after_linear=False (use_Key_instance_norm_after_linear_transformation=False, :470-472): the key input is
    instance-normalised once more on the UNPADDED map before Wk and Wk*K is used as it is (padded tokens = bk)."""
    B, H, W, C = xq.shape
    Hp, Wp = padded_dims(H, W, ws)
    q = _to_windows(instance_norm_bhwc(xq, affine=affine_q), ws, shift)  # affine_q / affine_k: the SAME modules as at :1052-1054
    if not key_in_after_linear:
        xk = instance_norm_bhwc(xk, affine=affine_k)
    k = F.linear(_to_windows(xk, ws, shift), wk, bk)
    vs = F.linear(_to_windows(xvs, ws, shift), wvs, bvs)
    vh = F.linear(_to_windows(xvh, ws, shift), wvh, bvh)
    if key_in_after_linear:
        # un-window k (still rolled, still padded), normalise per (b,c) over Hp*Wp, re-window
        gm = window_gather_map(H, W, ws, shift).reshape(-1).to(k.device)
        kmap = torch.empty(B, Hp * Wp, C, dtype=k.dtype, device=k.device)
        kmap[:, gm] = k.reshape(B, -1, C)
        kmap = instance_norm_bhwc(kmap.reshape(B, Hp, Wp, C), affine=affine_k).reshape(B, Hp * Wp, C)
        k = kmap[:, gm].reshape(k.shape)
    p = _softmax_probs(q, k, heads, _bias_from_table(table, ws), shift_mask(H, W, ws, shift), B)
    sig = F.linear(_apply_probs(p, vs, heads), wp, bp)
    mu = F.linear(_apply_probs(p, vh, heads), wp, bp)
    return _from_windows(sig, B, H, W, ws, shift), _from_windows(mu, B, H, W, ws, shift)


def mlp(x: Tensor, sd: SD, pre: str) -> Tensor:
    """torchvision ops/misc.py:264-306 MLP: Linear - GELU(erf) - Linear (dropout p=0)."""
    h = F.gelu(F.linear(x, sd[pre + "0.weight"], sd[pre + "0.bias"]))
    return F.linear(h, sd[pre + "3.weight"], sd[pre + "3.bias"])


def _attn_weights(sd: SD, pre: str):
    return [sd[pre + n] for n in ("Wq.weight", "Wq.bias", "Wk.weight", "Wk.bias", "Wv.weight", "Wv.bias",
                                  "proj.weight", "proj.bias", "relative_position_bias_table")]


# ----------------------------------------------------------------------------------------------
# style transformer (a5, a7, a9, a11)
# ----------------------------------------------------------------------------------------------


def _sd(x: Tensor, scales, i: int) -> Tensor:
    """torchvision StochasticDepth(p, "row") (ops/stochastic_depth.py) with the per-sample factors given explicitly:
    scales[i] is [B] = bernoulli(1-p)/(1-p); None = eval mode (identity)."""
    if scales is None:
        return x
    return x * scales[i].reshape(-1, *([1] * (x.dim() - 1)))


def style_encoder(sd: SD, key: Tensor, scale: Tensor, shift_t: Tensor, ws: int, sh: int, heads: int,
                  pre: str = "encoder.", sd_scales=None, processed_key: bool = True):
    """codes/style_transformer.py:855-882 default branch: one shared MHA (no norm, residual from
    input_q for the Key pass and from input_v for Scale/Shift, :383-386), three private MLPs.
    sd_scales: optional 9 per-sample stochastic-depth factor vectors in call order (:395,865,873,882).
    processed_key=False (encoder_if_use_processed_Key_in_Scale_and_Shift_calculation=False, :883-909): Scale and Shift
    attend with the layer's INPUT Key, the Key pass runs last (eval mode only: sd_scales must be None)."""
    aw = _attn_weights(sd, pre + "shared_MHA_without_MLP.attn.")
    if not processed_key:
        assert sd_scales is None
        scale = scale + window_attention(key, key, scale, *aw, ws, sh, heads)
        scale = scale + mlp(scale, sd, pre + "encoder_MLP_Scale.")
        shift_t = shift_t + window_attention(key, key, shift_t, *aw, ws, sh, heads)
        shift_t = shift_t + mlp(shift_t, sd, pre + "encoder_MLP_Shift.")
        key = key + window_attention(key, key, key, *aw, ws, sh, heads)
        key = key + mlp(key, sd, pre + "encoder_MLP_Key.")
        return key, scale, shift_t
    key = key + _sd(window_attention(key, key, key, *aw, ws, sh, heads), sd_scales, 0)
    key = key + _sd(mlp(key, sd, pre + "encoder_MLP_Key."), sd_scales, 1)
    scale = scale + _sd(window_attention(key, key, scale, *aw, ws, sh, heads), sd_scales, 2)
    scale = scale + _sd(mlp(scale, sd, pre + "encoder_MLP_Scale."), sd_scales, 3)
    shift_t = shift_t + _sd(window_attention(key, key, shift_t, *aw, ws, sh, heads), sd_scales, 4)
    shift_t = shift_t + _sd(mlp(shift_t, sd, pre + "encoder_MLP_Shift."), sd_scales, 5)
    return key, scale, shift_t


def _global_sigma_mu(sd: SD, query: Tensor, key: Tensor, scale: Tensor, shift_t: Tensor, pre: str, key_in_after_linear: bool):
    """decoder_use_regular_MHA_instead_of_Swin_at_the_end (codes/style_transformer.py:1063-1119): ONE head over all T = H*W
    tokens of an image, q = IN(Query) * C^-0.5 (no Q projection), k / v_scale / v_shift through linear_transformation_*, separate
    proj_sigma / proj_mu.  The reference feeds [B, C, T] tensors to nn.InstanceNorm2d, which reads a 3-D input as ONE unbatched
    (C', H', W') image: the statistics are taken over (C, T) JOINTLY per batch element, not per channel -- restated as it is."""
    B, H, W, C = query.shape
    lin = lambda x, n: F.linear(x, sd[pre + n + ".weight"], sd[pre + n + ".bias"])

    def joint_norm(x):  # x [B, T, C]
        mean = x.mean(dim=(1, 2), keepdim=True)
        var = x.var(dim=(1, 2), unbiased=False, keepdim=True)
        return (x - mean) / torch.sqrt(var + 1e-5)

    q, k = query.reshape(B, H * W, C), key.reshape(B, H * W, C)
    vs, vh = scale.reshape(B, H * W, C), shift_t.reshape(B, H * W, C)
    if key_in_after_linear:
        k = joint_norm(lin(k, "linear_transformation_Key"))
    else:
        k = lin(joint_norm(k), "linear_transformation_Key")
    q = joint_norm(q) * (C ** -0.5)
    p = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
    sigma = lin(p @ lin(vs, "linear_transformation_Scale"), "proj_sigma")
    mu = lin(p @ lin(vh, "linear_transformation_Shift"), "proj_mu")
    return sigma.reshape(B, H, W, C), mu.reshape(B, H, W, C)


def style_decoder(sd: SD, fcs: Tensor, key: Tensor, scale: Tensor, shift_t: Tensor, ws: int, sh: int, heads: int,
                  pre: str = "decoder.", sd_scales=None, key_in_after_linear: bool = True, exclude_mlp: bool = False,
                  affine_in: bool = False, regular_mha: bool = False):
    """codes/style_transformer.py:1045-1059,1123-1128 default branch (stochastic depth at :390,392,1125).
    exclude_mlp=True (decoder_exclude_MLP_after_Fcs_self_MHA, :339-343,365,389-392): the self-attention block has no
    norm2 / mlp; key_in_after_linear: see sigma_mu_attention."""
    b = pre + "MHA_self_attn."
    C = fcs.shape[-1]
    n1 = F.layer_norm(fcs, (C,), sd[b + "norm1.weight"], sd[b + "norm1.bias"])
    x = fcs + _sd(window_attention(n1, n1, n1, *_attn_weights(sd, b + "attn."), ws, sh, heads), sd_scales, 6)
    if not exclude_mlp:
        x = x + _sd(mlp(F.layer_norm(x, (C,), sd[b + "norm2.weight"], sd[b + "norm2.bias"]), sd, b + "mlp."), sd_scales, 7)
    if regular_mha:
        sigma, mu = _global_sigma_mu(sd, x, key, scale, shift_t, pre, key_in_after_linear)
        x = x * sigma + mu
        return x + _sd(mlp(x, sd, pre + "last_MLP."), sd_scales, 8)
    aq = (sd[pre + "instance_norm_Query.weight"], sd[pre + "instance_norm_Query.bias"]) if affine_in else None
    ak = (sd[pre + "instance_norm_Key.weight"], sd[pre + "instance_norm_Key.bias"]) if affine_in else None
    query_in = instance_norm_bhwc(x, affine=aq)
    key_in = instance_norm_bhwc(key, affine=ak)
    m = pre + "decoder_MHA_for_sigma_and_mu."
    sigma, mu = sigma_mu_attention(query_in, key_in, scale, shift_t,
                                   sd[m + "Wk.weight"], sd[m + "Wk.bias"],
                                   sd[m + "Wv_scale.weight"], sd[m + "Wv_scale.bias"],
                                   sd[m + "Wv_shift.weight"], sd[m + "Wv_shift.bias"],
                                   sd[m + "proj.weight"], sd[m + "proj.bias"],
                                   sd[m + "relative_position_bias_table"], ws, sh, heads,
                                   key_in_after_linear=key_in_after_linear, affine_q=aq, affine_k=ak)
    x = x * sigma + mu
    return x + _sd(mlp(x, sd, pre + "last_MLP."), sd_scales, 8)


def style_transformer(sd: SD, fc: Tensor, fs: Tensor, k: int = 1, ws: int = 8, sh: int = 4, heads: int = 8,
                      sd_scales=None, processed_key: bool = True, key_in_after_linear: bool = True,
                      exclude_mlp: bool = False, affine_in: bool = False, regular_mha: bool = False) -> Tensor:
    """codes/style_transformer.py:1229-1245: Scale=Shift=Fs, k times the same weights.
    sd_scales: optional [k, 9, B] train-mode stochastic-depth factors (None = eval).
    processed_key / key_in_after_linear / exclude_mlp: the reference's alternate configurations, see
    style_encoder / sigma_mu_attention / style_decoder."""
    scale, shift_t = fs, fs
    for l in range(k):
        sc = None if sd_scales is None else sd_scales[l]
        fs, scale, shift_t = style_encoder(sd, fs, scale, shift_t, ws, sh, heads, sd_scales=sc, processed_key=processed_key)
        fc = style_decoder(sd, fc, fs, scale, shift_t, ws, sh, heads, sd_scales=sc,
                           key_in_after_linear=key_in_after_linear, exclude_mlp=exclude_mlp, affine_in=affine_in,
                           regular_mha=regular_mha)
    return fc


# ----------------------------------------------------------------------------------------------
# Swin-B first two stages (a12) -- torchvision arithmetic
# ----------------------------------------------------------------------------------------------


def _tv_block(sd: SD, pre: str, x: Tensor, heads: int, shift: int, ws: int = 7) -> Tensor:
    """tv:401-456 block, tv:116-220 attention with fused qkv weight [3C,C]."""
    C = x.shape[-1]
    n1 = F.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    w, b = sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]
    a = window_attention(n1, n1, n1, w[:C], b[:C], w[C:2 * C], b[C:2 * C], w[2 * C:], b[2 * C:],
                         sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"],
                         sd[pre + "attn.relative_position_bias_table"], ws, shift, heads)
    x = x + a
    n2 = F.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    h = F.gelu(F.linear(n2, sd[pre + "mlp.0.weight"], sd[pre + "mlp.0.bias"]))
    return x + F.linear(h, sd[pre + "mlp.3.weight"], sd[pre + "mlp.3.bias"])


def swin_encoder(sd: SD, img: Tensor, pre: str = "") -> Tensor:
    """codes/utils.py:59-102 slice of tv swin_b: [B,3,S,S] NCHW -> [B,S/8,S/8,256] BHWC."""
    x = F.conv2d(img, sd[pre + "0.0.weight"], sd[pre + "0.0.bias"], stride=4).permute(0, 2, 3, 1)
    x = F.layer_norm(x, (128,), sd[pre + "0.2.weight"], sd[pre + "0.2.bias"])
    x = _tv_block(sd, pre + "1.0.", x, 4, 0)
    x = _tv_block(sd, pre + "1.1.", x, 4, 3)
    # patch merging tv:35-87 (H, W even here; odd sizes pad one row/col of zeros)
    B, H, W, C = x.shape
    x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    x = F.layer_norm(x, (4 * C,), sd[pre + "2.norm.weight"], sd[pre + "2.norm.bias"])
    x = F.linear(x, sd[pre + "2.reduction.weight"])
    x = _tv_block(sd, pre + "3.0.", x, 8, 0)
    x = _tv_block(sd, pre + "3.1.", x, 8, 3)
    return x


# ----------------------------------------------------------------------------------------------
# CNN decoder (a13), full model (a14)
# ----------------------------------------------------------------------------------------------

CNN_DECODER_LAYOUT = [  # (sequential index, upsample before this conv, relu after)
    (0, False, True), (3, True, True), (5, False, True), (7, False, True), (9, False, True),
    (12, True, True), (14, False, True), (17, True, True), (19, False, False)]


def cnn_decoder(sd: SD, x: Tensor, pre: str = "decoder.") -> Tensor:
    """codes/decoder.py:23-55: NCHW in, reflect-padded 3x3 convs, nearest x2 upsamples."""
    for idx, up, relu in CNN_DECODER_LAYOUT:
        if up:
            x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[f"{pre}{idx}.weight"], sd[f"{pre}{idx}.bias"])
        if relu:
            x = torch.relu(x)
    return x


def full_forward(sd: SD, content: Tensor, style: Tensor, k: int = 1, ws: int = 8, sh: int = 4, heads: int = 8) -> Tensor:
    """codes/full_model.py:214-226."""
    fc = swin_encoder(sd, content, "swin_encoder.")
    fs = swin_encoder(sd, style, "swin_encoder.")
    st = {n[len("style_transformer."):]: t for n, t in sd.items() if n.startswith("style_transformer.")}
    fcs = style_transformer(st, fc, fs, k, ws, sh, heads).permute(0, 3, 1, 2)
    return cnn_decoder(sd, fcs, "decoder.decoder.")


# ----------------------------------------------------------------------------------------------
# VGG-19 taps and loss (a15-a18)
# ----------------------------------------------------------------------------------------------

VGG_CONVS = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28]  # conv indices in features[:30]
VGG_POOL_BEFORE = {5, 10, 19, 28}                              # a 2x2 max-pool precedes these convs
VGG_TAPS = {5: 0, 10: 1, 19: 2, 28: 3}                         # relu after conv idx -> tap slot (relu2_1..5_1)


def vgg_taps(sd: SD, x: Tensor, pre: str = "") -> List[Tensor]:
    """codes/loss.py:23-37: relu2_1, relu3_1, relu4_1, relu5_1 of vgg19.features[:30]."""
    taps: List[Optional[Tensor]] = [None] * 4
    for idx in VGG_CONVS:
        if idx in VGG_POOL_BEFORE:
            x = F.max_pool2d(x, 2)
        x = torch.relu(F.conv2d(x, sd[f"{pre}{idx}.weight"], sd[f"{pre}{idx}.bias"], padding=1))
        if idx in VGG_TAPS:
            taps[VGG_TAPS[idx]] = x
    return taps  # type: ignore[return-value]


def vgg_bn_layout(last: int = 43):
    """(kind, index) of vgg19_bn.features[:43] (tv models/vgg.py cfg "E" with batch norm: conv-bn-relu triplets and max-pools;
    codes/utils.py:34-36 cuts at 43 = relu5_1)."""
    cfg = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]
    out, i = [], 0
    for v in cfg:
        if v == "M":
            out.append(("pool", i))
            i += 1
        else:
            out += [("conv", i), ("bn", i + 1), ("relu", i + 2)]
            i += 3
    return [e for e in out if e[1] < last]


VGG_BN_TAPS = {9: 0, 16: 1, 29: 2, 42: 3}  # last ReLU of features[:10], [10:17], [17:30], [30:43] (codes/loss.py:49-63)


def vgg_bn_taps(sd: SD, x: Tensor, pre: str = "", training: bool = True, eps: float = 1e-5) -> List[Tensor]:
    """codes/loss.py:41-63 VGG19_custom_with_batch_norm.  training=True is what the reference's scripts run: custom_loss is never
    put in eval mode (train.py:249-254), so every BatchNorm2d normalises with the statistics of the batch it is given -- biased
    variance over (N,H,W), content / style / output images in three separate calls (loss.py:223-225) -- and the running
    statistics only matter after an explicit .eval()."""
    taps: List[Optional[Tensor]] = [None] * 4
    for kind, i in vgg_bn_layout():
        if kind == "pool":
            x = F.max_pool2d(x, 2)
        elif kind == "conv":
            x = F.conv2d(x, sd[f"{pre}{i}.weight"], sd[f"{pre}{i}.bias"], padding=1)
        elif kind == "bn":
            if training:
                mean, var = x.mean(dim=(0, 2, 3)), x.var(dim=(0, 2, 3), unbiased=False)
            else:
                mean, var = sd[f"{pre}{i}.running_mean"], sd[f"{pre}{i}.running_var"]
            x = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps)
            x = x * sd[f"{pre}{i}.weight"].view(1, -1, 1, 1) + sd[f"{pre}{i}.bias"].view(1, -1, 1, 1)
        else:
            x = torch.relu(x)
            if i in VGG_BN_TAPS:
                taps[VGG_BN_TAPS[i]] = x
    return taps  # type: ignore[return-value]


def _in_nchw(x: Tensor, eps: float = 1e-5) -> Tensor:
    m = x.mean(dim=(2, 3), keepdim=True)
    v = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - m) / torch.sqrt(v + eps)


def content_loss(taps_c: List[Tensor], taps_o: List[Tensor], squared: bool = False) -> Tensor:
    """codes/loss.py:110-116,267-289: sum over taps of mean|IN(Fc)-IN(Fcs)| (or squared)."""
    tot = torch.zeros(())
    for a, b in zip(taps_c, taps_o):
        d = _in_nchw(a) - _in_nchw(b)
        tot = tot + (d.square().mean() if squared else d.abs().mean())
    return tot


def style_loss(taps_s: List[Tensor], taps_o: List[Tensor], squared: bool = False) -> Tensor:
    """codes/loss.py:122-130,293-315: mean|mu-mu'| + mean|std-std'| with torch's unbiased std."""
    tot = torch.zeros(())
    for a, b in zip(taps_s, taps_o):
        dm = a.mean(dim=(2, 3)) - b.mean(dim=(2, 3))
        ds = a.std(dim=(2, 3)) - b.std(dim=(2, 3))
        tot = tot + ((dm.square().mean() + ds.square().mean()) if squared else (dm.abs().mean() + ds.abs().mean()))
    return tot


def scaled_self_cosine_lower_triangle(a: Tensor, eps: float = 1e-6) -> Tensor:
    """codes/utils.py:105-133: a [B,C,H,W] -> [B,N,N]: cosine similarity of every pair of spatial positions
    (torch.cosine_similarity: each vector divided by max(its norm, 1e-8)), every COLUMN divided by its sum + eps, strict lower
    triangle kept.  (Restated with a matmul: the reference broadcasts a [B,N,N,C] tensor.)"""
    B, C = a.shape[:2]
    f = a.reshape(B, C, -1).permute(0, 2, 1)
    fh = f / f.norm(dim=2, keepdim=True).clamp_min(1e-8)
    d = fh @ fh.transpose(1, 2)
    return (d / (d.sum(dim=1) + eps).unsqueeze(1)).tril(diagonal=-1)


def similarity_loss(taps_a: List[Tensor], taps_b: List[Tensor], squared: bool = False) -> Tensor:
    """codes/loss.py:137-146,321-336: relu3_1 and relu4_1 terms, mean of |.| (or squares) over the WHOLE [B,N,N] maps.  The
    reference calls it with the content taps for BOTH arguments (:333-334) -> 0; the paper's form is (content, output)."""
    tot = torch.zeros(())
    for i in (1, 2):
        d = scaled_self_cosine_lower_triangle(taps_a[i]) - scaled_self_cosine_lower_triangle(taps_b[i])
        tot = tot + ((d * d).mean() if squared else d.abs().mean())
    return tot


def overall_loss(vgg_sd: SD, content: Tensor, style: Tensor, output: Tensor, lam: float = 10.0,
                 squared_content: bool = False, squared_style: bool = False, batchnorm: Optional[str] = None):
    """codes/loss.py:201-262: total = content + lambda*style; returns (total, content, style).
    batchnorm: None = vgg19 (default); "train" / "eval" = the use_vgg19_with_batchnorm variant in that module mode."""
    assert content.shape == style.shape == output.shape, "All images should be in the exact same shape"
    if batchnorm is None:
        taps = vgg_taps
    else:
        taps = lambda sd_, x: vgg_bn_taps(sd_, x, training=batchnorm == "train")  # noqa: E731
    tc, ts, to = taps(vgg_sd, content), taps(vgg_sd, style), taps(vgg_sd, output)
    lc = content_loss(tc, to, squared_content)
    ls = style_loss(ts, to, squared_style)
    return lc + lam * ls, lc, ls


# ----------------------------------------------------------------------------------------------
# uint8 image boundary either side of the path
# ----------------------------------------------------------------------------------------------

IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def images_u8_to_tensor(img_u8: Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> Tensor:
    """test_model.py:39-48 / get_dataloader.py:37-38 after the resize: transforms.ToTensor() (tv functional.to_tensor:
    HWC uint8 -> CHW, `.to(float32).div(255)`) then transforms.Normalize (tv functional.normalize: `sub_(mean).div_(std)` with
    fp32 mean / std tensors).  img_u8 [B,H,W,3] uint8 -> fp32 [B,3,H,W]; mean=None stops after the /255 (:111-125's flag off)."""
    x = img_u8.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
    if mean is None:
        return x
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    sd = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return x.sub(m).div(sd)


def tensor_to_images_u8(x: Tensor) -> Tensor:
    """test_model.py:207: np.clip(img.permute(1, 2, 0).numpy() * 255, 0, 255).astype(np.uint8), batched: [B,3,H,W] -> [B,H,W,3]."""
    import numpy as np
    a = x.detach().to(torch.float32).permute(0, 2, 3, 1).contiguous().numpy()
    return torch.from_numpy(np.clip(a * 255, 0, 255).astype(np.uint8))


# ----------------------------------------------------------------------------------------------
# training-image transform (codes/get_dataloader.py:30-36)
# ----------------------------------------------------------------------------------------------


def train_transform(img_u8_hwc, top: int, left: int, size=(512, 512), crop=(256, 256), mean=IMAGENET_MEAN, std=IMAGENET_STD) -> Tensor:
    """ToPILImage -> Resize(size) -> crop at (top, left) -> ToTensor -> Normalize, with the libraries the reference's transform
    itself is made of (torchvision.transforms.functional on a PIL image): the checker of the fused GPU transform.  RandomCrop's
    only random part is the choice of (top, left) (RandomCrop.get_params), passed in here."""
    import torchvision.transforms.functional as TF
    pil = TF.to_pil_image(img_u8_hwc if not isinstance(img_u8_hwc, Tensor) else img_u8_hwc.numpy())
    pil = TF.resize(pil, list(size))
    pil = TF.crop(pil, top, left, crop[0], crop[1])
    return TF.normalize(TF.to_tensor(pil), list(mean), list(std))

