"""Training-step engine: forward passes that keep what the adjoint needs, and
hand-sequenced backward passes over the C-ABI kernels -- no autograd inside; the nn.Module mirrors wrap these
in one torch.autograd.Function per reference module (style transformer, CNN decoder, loss), which is the
granularity at which the reference scripts call them (train.py:452-475).

Conventions: activations saved for the backward pass are bf16 token-major (exactly the
tensor-core operands the forward already produced) plus the few fp32 residual-stream tensors the LayerNorm /
InstanceNorm adjoints need; gradient residual streams are fp32 [T,C]; every GEMM operand gradient is bf16;
parameter gradients accumulate in fp32 (split-K tensor-core wgrad with fp32 RED).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .engine import CNN_LAYOUT, VGG_CONVS, VGG_POOL_BEFORE, VGG_TAP_AFTER, Workspace, _f32
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, GATE_GELU, GATE_RELU, PAD_REFLECT

BF16, F32 = torch.bfloat16, torch.float32


def _e16(dev, *shape):
    return torch.empty(*shape, dtype=BF16, device=dev)


def _e32(dev, *shape):
    return torch.empty(*shape, dtype=F32, device=dev)


class GradBook:
    """Flat fp32 gradient buffer with one view per parameter; selected runs of parameters are laid out adjacently so
    that a fused weight (e.g. [Wq;Wk;Wv]) has ONE contiguous gradient the wgrad kernel accumulates into."""

    def __init__(self, named_shapes: List[Tuple[str, torch.Size]], device):
        total = sum(int(torch.Size(s).numel()) for _, s in named_shapes)
        self.flat = torch.zeros(total, dtype=F32, device=device)
        self.views: Dict[str, torch.Tensor] = {}
        self.offsets: Dict[str, int] = {}
        off = 0
        for name, shape in named_shapes:
            n = int(torch.Size(shape).numel())
            self.views[name] = self.flat[off:off + n].view(shape)
            self.offsets[name] = off
            off += n

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.views[name]

    def span(self, first: str, rows: int, cols: Optional[int] = None) -> torch.Tensor:
        off = self.offsets[first]
        n = rows * (cols or 1)
        t = self.flat[off:off + n]
        return t.view(rows, cols) if cols else t


# ============================================================================================
# Style transformer (codes/style_transformer.py:777-1245, default flags)
# ============================================================================================


class _Lin:
    """A Linear in both directions: fwd = pack of W [N,K] (+bias), bwd = pack of W^T (data-gradient GEMM)."""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor]):
        self.N, self.K = int(w.shape[0]), int(w.shape[1])
        self.fwd = ops.pack_linear(w, b)
        self.bwd = ops.pack_linear(w.t().contiguous(), None)


class StyleTransformerTrainWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        g = lambda k: _f32(sd[prefix + k])
        lin = lambda p: _Lin(g(p + ".weight"), g(p + ".bias"))
        e = "encoder.shared_MHA_without_MLP.attn."
        wq, wk, wv = g(e + "Wq.weight"), g(e + "Wk.weight"), g(e + "Wv.weight")
        bq, bk, bv = g(e + "Wq.bias"), g(e + "Wk.bias"), g(e + "Wv.bias")
        self.C = int(wq.shape[0])
        self.enc_qkv = _Lin(torch.cat([wq, wk, wv], 0), torch.cat([bq, bk, bv], 0))
        self.enc_qk = _Lin(torch.cat([wq, wk], 0), torch.cat([bq, bk], 0))
        self.enc_v = _Lin(wv, bv)
        self.enc_proj = lin(e + "proj")
        self.enc_table = g(e + "relative_position_bias_table")
        self.mlp = {}
        d = "decoder.MHA_self_attn."
        # decoder_exclude_MLP_after_Fcs_self_MHA=True builds the block without norm2 / mlp (reference :339-343,365)
        self.has_dec_mlp = (prefix + d + "mlp.0.weight") in sd
        for tag, p in (("key", "encoder.encoder_MLP_Key."), ("scale", "encoder.encoder_MLP_Scale."), ("shift", "encoder.encoder_MLP_Shift."),
                       ("dec", d + "mlp."), ("last", "decoder.last_MLP.")):
            if tag == "dec" and not self.has_dec_mlp:
                continue
            self.mlp[tag] = (lin(p + "0"), lin(p + "3"), p)
        self.n1 = (g(d + "norm1.weight"), g(d + "norm1.bias"))
        self.n2 = (g(d + "norm2.weight"), g(d + "norm2.bias")) if self.has_dec_mlp else None
        a = d + "attn."
        self.dec_qkv = _Lin(torch.cat([g(a + "Wq.weight"), g(a + "Wk.weight"), g(a + "Wv.weight")], 0),
                            torch.cat([g(a + "Wq.bias"), g(a + "Wk.bias"), g(a + "Wv.bias")], 0))
        self.dec_proj = lin(a + "proj")
        self.dec_table = g(a + "relative_position_bias_table")
        m = "decoder.decoder_MHA_for_sigma_and_mu."
        self.sm_k, self.sm_vs, self.sm_vh, self.sm_proj = lin(m + "Wk"), lin(m + "Wv_scale"), lin(m + "Wv_shift"), lin(m + "proj")
        self.sm_table = g(m + "relative_position_bias_table")


E_ATT = "encoder.shared_MHA_without_MLP.attn."
D_BLK = "decoder.MHA_self_attn."
D_ATT = D_BLK + "attn."
SM = "decoder.decoder_MHA_for_sigma_and_mu."


def st_grad_book(param_shapes: Dict[str, torch.Size], device) -> GradBook:
    """Gradient layout for the style transformer: Wq/Wk/Wv (weights, then biases) adjacent for both attention blocks."""
    order = []
    for pre in (E_ATT, D_ATT):
        order += [pre + "Wq.weight", pre + "Wk.weight", pre + "Wv.weight", pre + "Wq.bias", pre + "Wk.bias", pre + "Wv.bias"]
    rest = [n for n in param_shapes if n not in order]
    return GradBook([(n, param_shapes[n]) for n in order + rest], device)


class _Geo:
    def __init__(self, B, H, W, C, heads, win, shift):
        self.B, self.H, self.W, self.C, self.heads, self.win, self.shift = B, H, W, C, heads, win, shift
        self.T = B * H * W
        self.HW = H * W
        # maps that are not a multiple of the window (7x7 windows on 32x32 maps, the reference CLI default, train.py:703-711) are
        # zero-padded at the bottom / right INSIDE every attention (style_transformer.py:77-87, :476-479): the training path
        # materialises the padded token map (pad -> projections + attention on [B,Hp,Wp] -> crop); T == Tp otherwise.
        self.Hp, self.Wp = -(-H // win) * win, -(-W // win) * win
        self.padded = (self.Hp, self.Wp) != (H, W)
        self.Tp, self.HWp = B * self.Hp * self.Wp, self.Hp * self.Wp


def _attn(g: _Geo, q, k, v, out, table, ldq, ldk, ldv, v2=None, out2=None):
    ops.window_attention(q, k, v, out, table, g.B, g.Hp, g.Wp, g.heads, g.win, g.shift, ldq, ldk, ldv, g.C, v2=v2, out2=out2)


def _attn_bwd(g: _Geo, q, k, v, dout, dq, dk, dv, table, dtable, ldq, ldk, ldv, lddq, lddk, lddv, v2=None, dout2=None, dv2=None):
    ops.window_attention_bwd(q, k, v, dout, dq, dk, dv, table, dtable, g.B, g.Hp, g.Wp, g.heads, g.win, g.shift, ldq, ldk, ldv, g.C,
                             lddq, lddk, lddv, v2=v2, dout2=dout2, dv2=dv2)


def _pad16(g: _Geo, x16, ws_: "Workspace" = None, name: str = ""):
    """bf16 [T,C] -> [Tp,C] with zero tokens at the bottom / right of every map; the tensor itself on window-multiple maps."""
    if not g.padded:
        return x16
    out = ws_.bf16(name, g.Tp, x16.shape[1]) if ws_ is not None else _e16(x16.device, g.Tp, x16.shape[1])
    ops.token_map_copy(x16, out, g.B, g.H, g.W, g.Hp, g.Wp)
    return out


def _crop16(g: _Geo, xp16, ws_: "Workspace" = None, name: str = ""):
    """bf16 [Tp,C] -> [T,C]: the reference's `x[:, :H, :W, :]` after the attention (and its adjoint's view of a padded gradient)."""
    if not g.padded:
        return xp16
    out = ws_.bf16(name, g.T, xp16.shape[1]) if ws_ is not None else _e16(xp16.device, g.T, xp16.shape[1])
    ops.token_map_copy(xp16, out, g.B, g.Hp, g.Wp, g.H, g.W)
    return out


def _mlp_fwd(x16, res32, fc1: _Lin, fc2: _Lin, T, dev, rs, rows, want16=True):
    """out32 = res32 + s * (fc2(gelu(fc1(x16)))) ; returns (out32, out16, hpre16, hact16)."""
    hid = fc1.N
    hpre, hact = _e16(dev, T, hid), _e16(dev, T, hid)
    ops.gemm(x16, fc1.fwd, T, act=ACT_GELU, out_bf16=hact, out_pre16=hpre)
    out32 = _e32(dev, T, fc2.N)
    out16 = _e16(dev, T, fc2.N) if want16 else None
    ops.gemm(hact, fc2.fwd, T, res=res32, out_f32=out32, out_bf16=out16, row_scale=rs, rows_per_scale=rows)
    return out32, out16, hpre, hact


_branch_streams: Dict[torch.device, list] = {}


def _parallel(main_fn, *side_fns):
    """main_fn on the current stream, every side_fn on a branch stream of its own (forked before, joined after) -- but only
    while the current stream is being captured into a CUDA graph, where the branches become parallel graph nodes: at batch 8
    the data-gradient GEMM, the weight-gradient GEMM and the bias column sum of one Linear each fill less than half of the
    148 SMs (64 row tiles), and all three only read dy.  Eager launches stay on one stream (the host is the limit there).
    The same side_fn slot always maps to the same branch stream, so accumulations into a shared gradient stay ordered."""
    if not side_fns or not torch.cuda.is_current_stream_capturing():
        main_fn()
        for f in side_fns:
            f()
        return
    cur = torch.cuda.current_stream()
    pool = _branch_streams.setdefault(cur.device, [])
    while len(pool) < len(side_fns):
        pool.append(torch.cuda.Stream(device=cur.device))
    for st, f in zip(pool, side_fns):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            f()
    main_fn()
    for st in pool[:len(side_fns)]:
        cur.wait_stream(st)


def _lin_bwd(dy16, x16, M, lin: _Lin, gW, gb, *, ld_dy=None, ld_x=None, res=None, out_f32=None, out_bf16=None, want_dx=True):
    """Adjoint of y = x W^T + b for bf16 dy [M,N]: dx (= dy W, optional residual add / fp32 / bf16 outputs), dW += dy^T x, db += colsum."""
    _parallel((lambda: ops.gemm(dy16, lin.bwd, M, lda=ld_dy, res=res, out_f32=out_f32, out_bf16=out_bf16)) if want_dx else (lambda: None),
              lambda: ops.wgrad(dy16, x16, gW, M, lin.N, lin.K, ld_dy=ld_dy, ld_x=ld_x),
              lambda: ops.colsum(dy16, M, lin.N, gb, ld=ld_dy))


def _lin_bwd_acc(g: _Geo, dy16p, x16p, lin: _Lin, gW, gb, gstream, ws_: "Workspace", *, ld_dy=None):
    """_lin_bwd on the (padded) attention map whose data gradient is added to the fp32 stream `gstream` on the unpadded map.
    The bias gradient sums over the padded tokens too (they hold the bias in the forward), the weight gradient sees their zero
    input, and their data gradient is dropped by the crop -- exactly the adjoint of F.pad -> Linear."""
    if not g.padded:
        _lin_bwd(dy16p, x16p, g.T, lin, gW, gb, ld_dy=ld_dy, res=gstream, out_f32=gstream)
        return
    tmp = ws_.f32("bw_dxp32", g.Tp, lin.K)
    _lin_bwd(dy16p, x16p, g.Tp, lin, gW, gb, ld_dy=ld_dy, out_f32=tmp)
    ops.token_map_copy(tmp, gstream, g.B, g.Hp, g.Wp, g.H, g.W, accumulate=True)


def _mlp_bwd(g16, x16, hpre, hact, fc1: _Lin, fc2: _Lin, T, book: GradBook, pre: str, ws_: Workspace, *, res=None, out_f32=None,
             out_bf16=None):
    """Adjoint of the MLP branch fc2(gelu(fc1(x))) for the (already stochastic-depth-scaled) bf16 branch gradient g16.
    The data gradient w.r.t. x goes to out_f32 (= res + dx when res is given) and/or out_bf16."""
    hid = fc1.N
    dh = ws_.bf16("bw_dh", T, hid)
    _parallel(lambda: ops.gemm(g16, fc2.bwd, T, out_bf16=dh, gate=hpre, gate_mode=GATE_GELU),  # (g W2) * gelu'(hpre)
              lambda: ops.wgrad(g16, hact, book[pre + "3.weight"], T, fc2.N, fc2.K),
              lambda: ops.colsum(g16, T, fc2.N, book[pre + "3.bias"]))
    _parallel(lambda: ops.gemm(dh, fc1.bwd, T, res=res, out_f32=out_f32, out_bf16=out_bf16),
              lambda: ops.wgrad(dh, x16, book[pre + "0.weight"], T, fc1.N, fc1.K),
              lambda: ops.colsum(dh, T, fc1.N, book[pre + "0.bias"]))


def _scaled16(g32, rs, rows, ws_: Workspace, name="bw_g16"):
    """bf16 copy of a gradient stream, times the per-sample stochastic-depth factor of the branch it feeds."""
    out = ws_.bf16(name, *g32.shape)
    if rs is None:
        ops.add_cast(g32, None, None, out)
    else:
        tmp = ws_.f32(name + "_f", *g32.shape)
        torch.mul(g32.view(rs.numel(), -1), rs.view(-1, 1), out=tmp.view(rs.numel(), -1))
        ops.add_cast(tmp, None, None, out)
    return out


def style_transformer_forward_train(w: StyleTransformerTrainWeights, fc32: torch.Tensor, fs32: torch.Tensor, k: int, g: _Geo,
                                    sd_scales: Optional[torch.Tensor], *, processed_key: bool = True, key_in_after_linear: bool = True,
                                    exclude_mlp: bool = False):
    """Forward of StyleTransformer.forward (:1229-1245) that records a tape.  sd_scales: None or fp32 [k, 9, B] per-sample
    stochastic-depth factors in the reference's draw order (enc MHA Key, MLP_Key, MHA Scale, MLP_Scale, MHA Shift, MLP_Shift,
    dec attn, dec mlp, last MLP).  Returns (out32 [T,C], tape).
    processed_key / key_in_after_linear / exclude_mlp: the reference's alternate orderings (engine.style_transformer_forward
    documents them): the Scale / Shift passes attend with the layer's INPUT Key; Key is instance-normalised twice BEFORE Wk and Wk.Key
    is used as it is; the decoder's self-attention block has no norm2 / MLP."""
    if exclude_mlp == w.has_dec_mlp:
        raise ValueError("decoder_exclude_MLP_after_Fcs_self_MHA does not match the packed state_dict (decoder.MHA_self_attn.mlp.*)")
    dev, T, C, Tp = fc32.device, g.T, g.C, g.Tp
    rows = g.HW
    x32 = fc32.reshape(T, C).contiguous()
    key32 = fs32.reshape(T, C).contiguous()
    scale32 = shift32 = key32
    key16 = _e16(dev, T, C)
    ops.cast_bf16(key32, key16)
    scale16 = shift16 = key16
    # attention inputs on the padded map (the tensors themselves when the map is a window multiple); carried across layers
    key16p = scale16p = shift16p = _pad16(g, key16)
    tape = []
    for l in range(k):
        s = (lambda i: sd_scales[l, i].contiguous()) if sd_scales is not None else (lambda i: None)
        t = {}
        # ---------------- StyleEncoder ----------------
        st_in = (key32, key16p, scale32, scale16p, shift32, shift16p)

        def enc_chain():
            key32, key16p, scale32, scale16p, shift32, shift16p = st_in
            t["key16_in"], t["scale16_in"], t["shift16_in"] = key16p, scale16p, shift16p
            qkv1, o1p = _e16(dev, Tp, 3 * C), _e16(dev, Tp, C)
            ops.gemm(key16p, w.enc_qkv.fwd, Tp, out_bf16=qkv1)
            _attn(g, qkv1, qkv1[:, C:], qkv1[:, 2 * C:], o1p, w.enc_table, 3 * C, 3 * C, 3 * C)
            o1 = _crop16(g, o1p)
            key32a, key16a = _e32(dev, T, C), _e16(dev, T, C)
            ops.gemm(o1, w.enc_proj.fwd, T, res=key32, out_f32=key32a, out_bf16=key16a, row_scale=s(0), rows_per_scale=rows)
            fc1, fc2, _ = w.mlp["key"]
            key32, key16, hp, ha = _mlp_fwd(key16a, key32a, fc1, fc2, T, dev, s(1), rows)
            key16p = _pad16(g, key16)
            t.update(qkv1=qkv1, o1=o1, key16a=key16a, hK=(hp, ha), key16b=key16p, key32b=key32)
            qk2, vs, vh = _e16(dev, Tp, 2 * C), _e16(dev, Tp, C), _e16(dev, Tp, C)
            # q, k of the Scale / Shift passes: the processed Key (default, :857-882) or this layer's input Key (:883-909); the
            # passes are independent of the Key pass in the second case, only their gradient lands on a different tensor
            t["qk2_in"] = key16p if processed_key else t["key16_in"]
            ops.gemm(t["qk2_in"], w.enc_qk.fwd, Tp, out_bf16=qk2)
            ops.gemm(scale16p, w.enc_v.fwd, Tp, out_bf16=vs)
            ops.gemm(shift16p, w.enc_v.fwd, Tp, out_bf16=vh)
            osp, ohp = _e16(dev, Tp, C), _e16(dev, Tp, C)
            _attn(g, qk2, qk2[:, C:], vs, osp, w.enc_table, 2 * C, 2 * C, C, v2=vh, out2=ohp)
            os_, oh = _crop16(g, osp), _crop16(g, ohp)
            scale32a, scale16a = _e32(dev, T, C), _e16(dev, T, C)
            ops.gemm(os_, w.enc_proj.fwd, T, res=scale32, out_f32=scale32a, out_bf16=scale16a, row_scale=s(2), rows_per_scale=rows)
            fc1, fc2, _ = w.mlp["scale"]
            scale32, scale16, hp, ha = _mlp_fwd(scale16a, scale32a, fc1, fc2, T, dev, s(3), rows)
            scale16p = _pad16(g, scale16)
            t.update(qk2=qk2, vs=vs, vh=vh, os=os_, oh=oh, scale16a=scale16a, hS=(hp, ha), scale16b=scale16p)
            shift32a, shift16a = _e32(dev, T, C), _e16(dev, T, C)
            ops.gemm(oh, w.enc_proj.fwd, T, res=shift32, out_f32=shift32a, out_bf16=shift16a, row_scale=s(4), rows_per_scale=rows)
            fc1, fc2, _ = w.mlp["shift"]
            shift32, shift16, hp, ha = _mlp_fwd(shift16a, shift32a, fc1, fc2, T, dev, s(5), rows)
            shift16p = _pad16(g, shift16)
            t.update(shift16a=shift16a, hH=(hp, ha), shift16b=shift16p)
            t["_enc_out"] = (key32, key16, key16p, scale32, scale16, scale16p, shift32, shift16, shift16p)

        # ---------------- StyleDecoder, self-attention half (independent of the encoder above: a parallel graph branch) ----------------
        x32_in = x32

        def dec_chain():
            ln1, qkv3, o3p = _e16(dev, T, C), _e16(dev, Tp, 3 * C), _e16(dev, Tp, C)
            ops.layernorm(x32_in, w.n1[0], w.n1[1], ln1, T, C)
            ln1 = _pad16(g, ln1)
            ops.gemm(ln1, w.dec_qkv.fwd, Tp, out_bf16=qkv3)
            _attn(g, qkv3, qkv3[:, C:], qkv3[:, 2 * C:], o3p, w.dec_table, 3 * C, 3 * C, 3 * C)
            o3 = _crop16(g, o3p)
            x32a = _e32(dev, T, C)
            ops.gemm(o3, w.dec_proj.fwd, T, res=x32_in, out_f32=x32a, row_scale=s(6), rows_per_scale=rows)
            if exclude_mlp:  # Query = Fcs + proj(attention) only (:389-392)
                query32 = x32a
                t.update(x32_in=x32_in, ln1=ln1, qkv3=qkv3, o3=o3, x32a=x32a, query32=query32)
            else:
                ln2 = _e16(dev, T, C)
                ops.layernorm(x32a, w.n2[0], w.n2[1], ln2, T, C)
                fc1, fc2, _ = w.mlp["dec"]
                query32, _, hp, ha = _mlp_fwd(ln2, x32a, fc1, fc2, T, dev, s(7), rows, want16=False)
                t.update(x32_in=x32_in, ln1=ln1, qkv3=qkv3, o3=o3, x32a=x32a, ln2=ln2, hD=(hp, ha), query32=query32)
            qmean, qrstd = _e32(dev, g.B, C), _e32(dev, g.B, C)
            qhat = _e16(dev, T, C)
            ops.instnorm_stats(query32, qmean, qrstd, g.B, g.HW, C, twice=True)
            ops.instnorm_apply(query32, qmean, qrstd, g.B, g.HW, C, y16=qhat)
            t["_qhat"] = qhat

        _parallel(enc_chain, dec_chain)
        key32, key16, key16p, scale32, scale16, scale16p, shift32, shift16, shift16p = t.pop("_enc_out")
        query32, qhat = t["query32"], t.pop("_qhat")
        mean, rstd = _e32(dev, g.B, C), _e32(dev, g.B, C)
        kin, khat = _e16(dev, T, C), _e16(dev, T, C)
        ops.instnorm_stats(key32, mean, rstd, g.B, g.HW, C, twice=not key_in_after_linear)
        ops.instnorm_apply(key32, mean, rstd, g.B, g.HW, C, y16=kin)
        # padded q tokens are zero (no Q projection, :511-514); Wk runs on the padded Key and its InstanceNorm over the PADDED map (:520-530)
        qhat, kin = _pad16(g, qhat), _pad16(g, kin)
        khat = khat if not g.padded else _e16(dev, Tp, C)
        if key_in_after_linear:
            kk32 = _e32(dev, Tp, C)
            ops.gemm(kin, w.sm_k.fwd, Tp, out_f32=kk32)
            ops.instnorm_stats(kk32, mean, rstd, g.B, g.HWp, C)
            ops.instnorm_apply(kk32, mean, rstd, g.B, g.HWp, C, y16=khat)
        else:  # IN(IN(Key)) on the unpadded map (:1057 then :470-472), k = Wk.Key + bk as it is (padded tokens: bk, :520)
            kk32 = None
            ops.gemm(kin, w.sm_k.fwd, Tp, out_bf16=khat)
        vs2, vh2, osgp, omup = _e16(dev, Tp, C), _e16(dev, Tp, C), _e16(dev, Tp, C), _e16(dev, Tp, C)
        ops.gemm(scale16p, w.sm_vs.fwd, Tp, out_bf16=vs2)
        ops.gemm(shift16p, w.sm_vh.fwd, Tp, out_bf16=vh2)
        _attn(g, qhat, khat, vs2, osgp, w.sm_table, C, C, C, v2=vh2, out2=omup)
        osg, omu = _crop16(g, osgp), _crop16(g, omup)
        sigma32, y32, y16 = _e32(dev, T, C), _e32(dev, T, C), _e16(dev, T, C)
        ops.gemm(osg, w.sm_proj.fwd, T, out_f32=sigma32)
        ops.gemm(omu, w.sm_proj.fwd, T, res=query32, mul=sigma32, out_f32=y32, out_bf16=y16)
        fc1, fc2, _ = w.mlp["last"]
        x32, _, hp, ha = _mlp_fwd(y16, y32, fc1, fc2, T, dev, s(8), rows, want16=False)
        t.update(qhat=qhat, kin=kin, kk32=kk32, khat=khat, vs2=vs2, vh2=vh2, osg=osg, omu=omu, sigma32=sigma32, y16=y16, hL=(hp, ha))
        tape.append(t)
    return x32, tape


def style_transformer_backward(w: StyleTransformerTrainWeights, tape, g_out: torch.Tensor, g: _Geo, sd_scales: Optional[torch.Tensor],
                               book: GradBook, ws_: Workspace, *, processed_key: bool = True, key_in_after_linear: bool = True,
                               exclude_mlp: bool = False):
    """Adjoint of style_transformer_forward_train: accumulates every parameter gradient into `book`; returns nothing for
    Fc / Fs (the Swin encoder is frozen in the reference's default training setup, train.py:216-218)."""
    dev, T, C, Tp = g_out.device, g.T, g.C, g.Tp
    rows = g.HW
    gx = g_out.reshape(T, C).clone()  # grad w.r.t. the layer output (fp32 stream, updated in place)
    gkey = torch.zeros(T, C, dtype=F32, device=dev)
    gscale, gshift = torch.zeros_like(gkey), torch.zeros_like(gkey)
    coef = ws_.f32("bw_coef", g.B, C, 4)
    enc_qkv_w, enc_qkv_b = book.span(E_ATT + "Wq.weight", 3 * C, C), book.span(E_ATT + "Wq.bias", 3 * C)
    dec_qkv_w, dec_qkv_b = book.span(D_ATT + "Wq.weight", 3 * C, C), book.span(D_ATT + "Wq.bias", 3 * C)
    for l in reversed(range(len(tape))):
        t = tape[l]
        s = (lambda i: sd_scales[l, i].contiguous()) if sd_scales is not None else (lambda i: None)
        # ---- out = y + s8 * MLP_L(y16)
        fc1, fc2, pre = w.mlp["last"]
        g16 = _scaled16(gx, s(8), rows, ws_)
        _mlp_bwd(g16, t["y16"], t["hL"][0], t["hL"][1], fc1, fc2, T, book, pre, ws_, res=gx, out_f32=gx)  # gx = gy
        # ---- y = query * sigma + mu
        gquery = _e32(dev, T, C)
        gsig16, gmu16 = ws_.bf16("bw_gsig", T, C), ws_.bf16("bw_gmu", T, C)
        ops.blend_bwd(gx, t["sigma32"], t["query32"], gquery, gsig16, gmu16)
        # ---- sigma = proj(o_sigma), mu = proj(o_mu) (one shared proj)
        dosg, domu = ws_.bf16("bw_dosg", T, C), ws_.bf16("bw_domu", T, C)
        _lin_bwd(gsig16, t["osg"], T, w.sm_proj, book[SM + "proj.weight"], book[SM + "proj.bias"], out_bf16=dosg)
        _lin_bwd(gmu16, t["omu"], T, w.sm_proj, book[SM + "proj.weight"], book[SM + "proj.bias"], out_bf16=domu)
        # ---- shared-softmax sigma/mu attention
        dqhat, dkhat, dvs2, dvh2 = (ws_.bf16(n, Tp, C) for n in ("bw_dqhat", "bw_dkhat", "bw_dvs2", "bw_dvh2"))
        dosg, domu = _pad16(g, dosg, ws_, "bw_dosg_p"), _pad16(g, domu, ws_, "bw_domu_p")  # the crop's adjoint: zero gradient at padded tokens
        _attn_bwd(g, t["qhat"], t["khat"], t["vs2"], dosg, dqhat, dkhat, dvs2, w.sm_table, book[SM + "relative_position_bias_table"],
                  C, C, C, C, C, C, v2=t["vh2"], dout2=domu, dv2=dvh2)
        _lin_bwd_acc(g, dvs2, t["scale16b"], w.sm_vs, book[SM + "Wv_scale.weight"], book[SM + "Wv_scale.bias"], gscale, ws_)
        _lin_bwd_acc(g, dvh2, t["shift16b"], w.sm_vh, book[SM + "Wv_shift.weight"], book[SM + "Wv_shift.bias"], gshift, ws_)
        # ---- khat = IN_padded(Wk pad(IN(Key))) ; qhat = pad(IN(IN(Query)))
        dkk16, dkin16 = ws_.bf16("bw_dkk", Tp, C), ws_.bf16("bw_dkin", Tp, C)
        if key_in_after_linear:
            ops.instnorm_bwd(t["kk32"], dkhat, coef, g.B, g.HWp, C, dx16=dkk16)
        else:  # khat = Wk pad(IN(IN(Key))) + bk: no normalisation after the projection
            dkk16 = dkhat
        _lin_bwd(dkk16, t["kin"], Tp, w.sm_k, book[SM + "Wk.weight"], book[SM + "Wk.bias"], out_bf16=dkin16)
        dkin16, dqhat = _crop16(g, dkin16, ws_, "bw_dkin_c"), _crop16(g, dqhat, ws_, "bw_dqhat_c")
        ops.instnorm_bwd(t["key32b"], dkin16, coef, g.B, g.HW, C, twice=not key_in_after_linear, dx_accum=gkey)
        ops.instnorm_bwd(t["query32"], dqhat, coef, g.B, g.HW, C, twice=True, dx_accum=gquery)
        # ---- query = x_a + s7 * MLP_D(LN2(x_a))   (query = x_a with decoder_exclude_MLP_after_Fcs_self_MHA)
        dln = ws_.bf16("bw_dln", T, C)
        if not exclude_mlp:
            fc1, fc2, pre = w.mlp["dec"]
            g16 = _scaled16(gquery, s(7), rows, ws_)
            _mlp_bwd(g16, t["ln2"], t["hD"][0], t["hD"][1], fc1, fc2, T, book, pre, ws_, out_bf16=dln)
            ops.layernorm_bwd(t["x32a"], w.n2[0], dln, gquery, book[D_BLK + "norm2.weight"], book[D_BLK + "norm2.bias"], T, C)  # gquery = g(x_a)
        # ---- x_a = x + s6 * proj(attn(qkv(LN1(x))))
        g16 = _scaled16(gquery, s(6), rows, ws_)
        do3 = ws_.bf16("bw_do", T, C)
        _lin_bwd(g16, t["o3"], T, w.dec_proj, book[D_ATT + "proj.weight"], book[D_ATT + "proj.bias"], out_bf16=do3)
        dqkv = ws_.bf16("bw_dqkv", Tp, 3 * C)
        q3 = t["qkv3"]
        do3 = _pad16(g, do3, ws_, "bw_do_p")
        _attn_bwd(g, q3, q3[:, C:], q3[:, 2 * C:], do3, dqkv, dqkv[:, C:], dqkv[:, 2 * C:], w.dec_table,
                  book[D_ATT + "relative_position_bias_table"], 3 * C, 3 * C, 3 * C, 3 * C, 3 * C, 3 * C)
        dlnp = ws_.bf16("bw_dln_p", Tp, C) if g.padded else dln
        _lin_bwd(dqkv, t["ln1"], Tp, w.dec_qkv, dec_qkv_w, dec_qkv_b, out_bf16=dlnp)
        if g.padded:
            ops.token_map_copy(dlnp, dln, g.B, g.Hp, g.Wp, g.H, g.W)
        ops.layernorm_bwd(t["x32_in"], w.n1[0], dln, gquery, book[D_BLK + "norm1.weight"], book[D_BLK + "norm1.bias"], T, C)
        gx = gquery  # grad w.r.t. this layer's Fcs input = previous layer's output

        # ---------------- StyleEncoder (reverse) ----------------
        pw, pb = book[E_ATT + "proj.weight"], book[E_ATT + "proj.bias"]
        dos, doh = ws_.bf16("bw_dos", T, C), ws_.bf16("bw_doh", T, C)
        for gstream, mtag, a16, hkey, o_saved, d_o, i_mlp, i_att in ((gshift, "shift", "shift16a", "hH", "oh", doh, 5, 4),
                                                                    (gscale, "scale", "scale16a", "hS", "os", dos, 3, 2)):
            fc1, fc2, pre = w.mlp[mtag]
            g16 = _scaled16(gstream, s(i_mlp), rows, ws_)
            _mlp_bwd(g16, t[a16], t[hkey][0], t[hkey][1], fc1, fc2, T, book, pre, ws_, res=gstream, out_f32=gstream)
            g16 = _scaled16(gstream, s(i_att), rows, ws_)
            _lin_bwd(g16, t[o_saved], T, w.enc_proj, pw, pb, out_bf16=d_o)
        dqk2, dvs, dvh = ws_.bf16("bw_dqk2", Tp, 2 * C), ws_.bf16("bw_dvs", Tp, C), ws_.bf16("bw_dvh", Tp, C)
        qk2 = t["qk2"]
        etab = book[E_ATT + "relative_position_bias_table"]
        dosp, dohp = _pad16(g, dos, ws_, "bw_dos_p"), _pad16(g, doh, ws_, "bw_doh_p")
        _attn_bwd(g, qk2, qk2[:, C:], t["vs"], dosp, dqk2, dqk2[:, C:], dvs, w.enc_table, etab, 2 * C, 2 * C, C, 2 * C, 2 * C, C,
                  v2=t["vh"], dout2=dohp, dv2=dvh)
        wv_w, wv_b = enc_qkv_w[2 * C:], enc_qkv_b[2 * C:]
        _lin_bwd_acc(g, dvs, t["scale16_in"], w.enc_v, wv_w, wv_b, gscale, ws_)
        _lin_bwd_acc(g, dvh, t["shift16_in"], w.enc_v, wv_w, wv_b, gshift, ws_)
        if processed_key:  # q, k of the Scale / Shift passes came from the PROCESSED Key: their gradient joins gkey before the Key pass' adjoint
            _lin_bwd_acc(g, dqk2, t["qk2_in"], w.enc_qk, enc_qkv_w[:2 * C], enc_qkv_b[:2 * C], gkey, ws_)
        fc1, fc2, pre = w.mlp["key"]
        g16 = _scaled16(gkey, s(1), rows, ws_)
        _mlp_bwd(g16, t["key16a"], t["hK"][0], t["hK"][1], fc1, fc2, T, book, pre, ws_, res=gkey, out_f32=gkey)
        g16 = _scaled16(gkey, s(0), rows, ws_)
        _lin_bwd(g16, t["o1"], T, w.enc_proj, pw, pb, out_bf16=dos)
        q1 = t["qkv1"]
        dosp = _pad16(g, dos, ws_, "bw_dos_p")
        _attn_bwd(g, q1, q1[:, C:], q1[:, 2 * C:], dosp, dqkv, dqkv[:, C:], dqkv[:, 2 * C:], w.enc_table, etab,
                  3 * C, 3 * C, 3 * C, 3 * C, 3 * C, 3 * C)
        _lin_bwd_acc(g, dqkv, t["key16_in"], w.enc_qkv, enc_qkv_w, enc_qkv_b, gkey, ws_)
        if not processed_key:  # ... from the layer's INPUT Key: it joins gkey after the Key pass' adjoint has mapped gkey back to the input
            _lin_bwd_acc(g, dqk2, t["qk2_in"], w.enc_qk, enc_qkv_w[:2 * C], enc_qkv_b[:2 * C], gkey, ws_)
        if l == 0:
            break
        # layer l-1's Key/Scale/Shift outputs feed this layer: the streams carry over as they are
    return None


# ============================================================================================
# CNN decoder (codes/decoder.py:23-55)
# ============================================================================================


class CnnDecoderTrainWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = "decoder."):
        self.layers = []
        for idx, up, relu in CNN_LAYOUT:
            wt, bias = _f32(sd[f"{prefix}{idx}.weight"]), _f32(sd[f"{prefix}{idx}.bias"])
            cout, cin = int(wt.shape[0]), int(wt.shape[1])
            fwd = ops.pack_conv3x3(wt, bias)
            # data gradient = full correlation with the transposed, spatially flipped kernel; a 3-channel dY travels padded to 8
            wT = wt.permute(1, 0, 2, 3).flip(2, 3).contiguous()
            cpad = (cout + 7) // 8 * 8
            if cpad != cout:
                wT = torch.cat([wT, wT.new_zeros(cin, cpad - cout, 3, 3)], 1).contiguous()
            bwd = ops.pack_conv3x3(wT, None)
            self.layers.append(dict(idx=idx, up=up, relu=relu, cin=cin, cout=cout, cpad=cpad, fwd=fwd, bwd=bwd,
                                    wname=f"{prefix}{idx}.weight", bname=f"{prefix}{idx}.bias"))


def cnn_decoder_forward_train(w: CnnDecoderTrainWeights, x16: torch.Tensor, B: int, H: int, W: int, out: torch.Tensor):
    """x16 bf16 [B*H*W,256] -> out fp32 [B,3,8H,8W]; returns the list of conv inputs (bf16 NHWC) for the backward pass."""
    dev = x16.device
    acts = [x16]
    cur, h, wd = x16, H, W
    last = len(w.layers) - 1
    for i, L in enumerate(w.layers):
        if L["up"]:
            h, wd = 2 * h, 2 * wd
        M = B * h * wd
        conv = dict(H=h, W=wd, Cin=L["cin"], pad_mode=PAD_REFLECT, upsample=L["up"])
        if i == last:
            conv.update(out_nchw=True, n_real=L["cout"])
            ops.gemm(cur, L["fwd"], M, act=ACT_NONE, out_f32=out, conv=conv)
        else:
            nxt = _e16(dev, M, L["fwd"].n_pad)
            ops.gemm(cur, L["fwd"], M, act=ACT_RELU if L["relu"] else ACT_NONE, out_bf16=nxt, conv=conv)
            acts.append(nxt)
            cur = nxt
    return acts


def cnn_decoder_backward(w: CnnDecoderTrainWeights, acts, g_out: torch.Tensor, B: int, H: int, W: int, book: GradBook, ws_: Workspace,
                         dbg: Optional[dict] = None):
    """g_out fp32 [B,3,8H,8W] -> returns bf16 [B*H*W,256] gradient w.r.t. the decoder input; parameter grads into `book`.
    dbg (tests only): receives a copy of the gradient w.r.t. every conv's output."""
    S_h, S_w = 8 * H, 8 * W
    g = ws_.bf16("cb_g8", B * S_h * S_w, 8)
    ops.nchw3_to_nhwc8(g_out.contiguous(), g, B, S_h, S_w)
    h, wd = S_h, S_w
    for i in reversed(range(len(w.layers))):
        L = w.layers[i]
        M = B * h * wd
        cpad = L["cpad"]
        if dbg is not None:
            dbg[i] = g.clone()
        ops.wgrad(g, acts[i], book[L["wname"]], M, cpad, 9 * L["cin"], ld_dy=cpad,
                  conv=dict(H=h, W=wd, Cin=L["cin"], pad_mode=PAD_REFLECT, upsample=L["up"]), n_real=L["cout"] if cpad != L["cout"] else 0)
        if cpad == L["cout"]:
            ops.colsum(g, M, cpad, book[L["bname"]], ld=cpad)
        else:
            tmp = ws_.f32("cb_bias8", cpad)
            tmp.zero_()
            ops.colsum(g, M, cpad, tmp, ld=cpad)
            book[L["bname"]].add_(tmp[:L["cout"]])
        # data gradient on the reflect-padded grid, then fold (+ 2x2 sum for the upsample, + previous ReLU mask)
        Hp, Wp = h + 2, wd + 2
        dxp = ws_.bf16("cb_dxp", B * Hp * Wp, L["bwd"].n_pad)
        ops.gemm(g, L["bwd"], B * Hp * Wp, out_bf16=dxp, conv=dict(H=Hp, W=Wp, Cin=cpad, pad_mode=0, full=True))
        ho, wo = (h // 2, wd // 2) if L["up"] else (h, wd)
        gprev = ws_.bf16(f"cb_g{i % 2}", B * ho * wo, L["bwd"].n_pad)
        ops.reflect_fold(dxp, acts[i] if i > 0 else None, gprev, B, h, wd, L["bwd"].n_pad, upsample=L["up"])
        g, h, wd = gprev, ho, wo
    return g


# ============================================================================================
# VGG-19 perceptual loss (codes/loss.py:15-37,71-336)
# ============================================================================================


class VggTrainWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = "features."):
        self.first_w, self.first_b = _f32(sd[prefix + "0.weight"]), _f32(sd[prefix + "0.bias"])
        self.first_bwd = ops.pack_conv3x3(self.first_w.permute(1, 0, 2, 3).flip(2, 3).contiguous(), None)  # [3(->16), 9*64]
        self.convs = {}
        for idx in VGG_CONVS[1:]:
            wt = _f32(sd[f"{prefix}{idx}.weight"])
            self.convs[idx] = dict(fwd=ops.pack_conv3x3(wt, _f32(sd[f"{prefix}{idx}.bias"])), cin=int(wt.shape[1]), cout=int(wt.shape[0]),
                                   bwd=ops.pack_conv3x3(wt.permute(1, 0, 2, 3).flip(2, 3).contiguous(), None))


def vgg_forward_saving(w: VggTrainWeights, imgs: torch.Tensor):
    """VGG-19 features[:30] on imgs fp32 [N,3,H,W] keeping every ReLU output (bf16 NHWC) for the backward pass.
    Returns (acts: {conv idx: (tensor, h, w, c)}, pooled: {conv idx: tensor feeding that conv}, taps[4])."""
    dev = imgs.device
    N, _, H, W = imgs.shape
    acts, pooled = {}, {}
    cur = _e16(dev, N * H * W, 64)
    ops.conv3x3_first(imgs, w.first_w, w.first_b, cur, N, H, W, relu=True)
    acts[0] = (cur, H, W, 64)
    h, wd, c = H, W, 64
    taps = [None] * 4
    for idx in VGG_CONVS[1:]:
        L = w.convs[idx]
        if idx in VGG_POOL_BEFORE:
            p = _e16(dev, N * (h // 2) * (wd // 2), c)
            ops.maxpool2x2(cur, p, N, h, wd, c)
            cur, h, wd = p, h // 2, wd // 2
            pooled[idx] = p
        out = _e16(dev, N * h * wd, L["fwd"].n_pad)
        ops.gemm(cur, L["fwd"], N * h * wd, act=ACT_RELU, out_bf16=out, conv=dict(H=h, W=wd, Cin=L["cin"], pad_mode=0, upsample=False))
        cur, c = out, L["fwd"].n_pad
        acts[idx] = (out, h, wd, c)
        if idx in VGG_TAP_AFTER:
            taps[VGG_TAP_AFTER[idx]] = (out, h, wd, c)
    return acts, pooled, taps


def content_style_taps(infer_w, content, style, ws_: Workspace):
    """VGG taps and per-channel statistics of the content and style images (nothing saved for a backward: they carry no
    gradient).  Independent of the stylised image, so a captured training step runs it as a parallel graph branch next to the
    encoder / style transformer / decoder forward (training.InnerLoopTrainer.step)."""
    from .engine import vgg_taps_forward
    B = int(content.shape[0])
    dev = content.device
    cs = ws_.f32("lt_imgs", 2 * B, 3, content.shape[2], content.shape[3])
    cs[:B].copy_(content)
    cs[B:].copy_(style)
    taps_cs = vgg_taps_forward(infer_w, cs, ws_, "lt_cs_")
    stats = []
    for tcs, h, wd, c in taps_cs:
        mean_cs, var_cs = _e32(dev, 2 * B, c), _e32(dev, 2 * B, c)
        ops.tap_stats(tcs, mean_cs, var_cs, 2 * B, h * wd, c)
        stats.append((mean_cs, var_cs))
    return taps_cs, stats


def perceptual_loss_forward_train(w: VggTrainWeights, infer_w, content, style, output, lam: float, squared_content: bool,
                                  squared_style: bool, ws_: Workspace, pre=None):
    """Loss forward that keeps the stylised image's VGG activations.  content/style taps come from the inference path
    (ping-pong buffers, nothing saved); pre = an earlier content_style_taps(...) result for the same images.
    Returns (out3 device tensor, ctx dict for perceptual_loss_backward)."""
    B = int(content.shape[0])
    dev = output.device
    taps_cs, stats_cs = pre if pre is not None else content_style_taps(infer_w, content, style, ws_)
    acts, pooled, taps_o = vgg_forward_saving(w, output)
    descs, per_tap = [], []
    for i in range(4):
        tcs, h, wd, c = taps_cs[i]
        to = taps_o[i][0]
        T = h * wd
        mean_cs, var_cs = stats_cs[i]
        mean_o, var_o = _e32(dev, B, c), _e32(dev, B, c)
        ops.tap_stats(to, mean_o, var_o, B, T, c)
        partials = ws_.f32(f"lt_part{i}", 592)
        fc = tcs.view(2 * B, T * c)[:B]
        ops.content_term(fc, to, mean_cs[:B], var_cs[:B], mean_o, var_o, B, T, c, squared_content, partials)
        descs.append(dict(partials=partials, mean_s=mean_cs[B:], var_s=var_cs[B:], mean_o=mean_o, var_o=var_o, B=B, T=T, C=c))
        # the content tap must outlive the workspace's ping-pong reuse -> private copy (bf16, B*T*c)
        per_tap.append(dict(fc=fc.clone(), fo=to, mean_c=mean_cs[:B], var_c=var_cs[:B], mean_s=mean_cs[B:], var_s=var_cs[B:],
                            mean_o=mean_o, var_o=var_o, T=T, C=c, h=h, w=wd))
    out3 = torch.empty(3, dtype=F32, device=dev)
    ops.loss_finalize(descs, lam, squared_style, out3)
    ctx = dict(acts=acts, pooled=pooled, taps=per_tap, B=B, H=int(output.shape[2]), W=int(output.shape[3]),
               sq_c=squared_content, sq_s=squared_style)
    return out3, ctx


def perceptual_loss_backward(w: VggTrainWeights, ctx, coef2: torch.Tensor, ws_: Workspace) -> torch.Tensor:
    """coef2 = device fp32 [2] {d total/d content, d total/d style} -> gradient w.r.t. the stylised image, fp32 [B,3,H,W]."""
    B, H, W = ctx["B"], ctx["H"], ctx["W"]
    dev = coef2.device
    acts, pooled = ctx["acts"], ctx["pooled"]
    tapgrad = {}
    for i, tp in enumerate(ctx["taps"]):
        dfo = ws_.bf16(f"lb_tap{i}", B, tp["T"], tp["C"])
        s = ws_.f32(f"lb_s{i}", B, tp["C"], 2)
        ops.loss_bwd(tp["fc"], tp["fo"], tp["mean_c"], tp["var_c"], tp["mean_o"], tp["var_o"], tp["mean_s"], tp["var_s"], s, coef2,
                     B, tp["T"], tp["C"], ctx["sq_c"], ctx["sq_s"], dfo)
        tapgrad[i] = dfo
    conv_ids = VGG_CONVS
    tap_of = VGG_TAP_AFTER
    g = tapgrad[3].view(-1, ctx["taps"][3]["C"])  # grad w.r.t. conv 28's pre-activation (already ReLU-masked)
    dbg = ctx.get("debug")
    for pos in range(len(conv_ids) - 1, 0, -1):
        idx, prev = conv_ids[pos], conv_ids[pos - 1]
        if dbg is not None:
            dbg[idx] = g.clone()
        L = w.convs[idx]
        out_t, h, wd, _ = acts[idx]
        prev_t, ph, pw_, pc = acts[prev]
        M = B * h * wd
        conv = dict(H=h, W=wd, Cin=L["cout"], pad_mode=0, upsample=False, impl="gather")
        if idx in VGG_POOL_BEFORE:
            gp = ws_.bf16("lb_gpool", M, L["bwd"].n_pad)
            ops.gemm(g, L["bwd"], M, out_bf16=gp, conv=conv)
            gn = ws_.bf16(f"lb_g{pos % 2}", B * ph * pw_, pc)
            ops.maxpool2x2_bwd(prev_t, gp, gn, B, ph, pw_, pc)
        else:
            gn = ws_.bf16(f"lb_g{pos % 2}", M, L["bwd"].n_pad)
            add = tapgrad[tap_of[prev]].view(M, -1) if prev in tap_of else None
            ops.gemm(g, L["bwd"], M, out_bf16=gn, conv=conv, gate=prev_t, gate_mode=GATE_RELU, add16=add)
        g = gn
    if dbg is not None:
        dbg[0] = g.clone()
    # features[0]: Conv2d(3,64) on the fp32 NCHW image
    dimg = torch.empty(B, 3, H, W, dtype=F32, device=dev)
    ops.gemm(g, w.first_bwd, B * H * W, out_f32=dimg,
             conv=dict(H=H, W=W, Cin=64, pad_mode=0, upsample=False, out_nchw=True, n_real=3, impl="gather"))
    return dimg
