"""torch.autograd.Function wrappers that put the training engine (train_engine.py) behind the reference's module
boundaries: StyleTransformer.forward, Decoder.forward and custom_loss.forward each record ONE autograd node whose
backward runs the hand-written adjoint kernels.  This is what lets the reference's own training step

    total_loss.backward(); inner_loop_optimizer.step()          (train.py:515-517)

run unmodified on top of the sm_100a path: autograd only sequences three nodes and accumulates .grad.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import engine, ops, train_engine as te
from .style_transformer import packed_weights, workspace_of


def _param_lists(module: torch.nn.Module):
    named = list(module.named_parameters())
    return [n for n, _ in named], [p for _, p in named]


def _grads_out(names, params, book: te.GradBook):
    return tuple(book[n] if p.requires_grad else None for n, p in zip(names, params))


# --------------------------------------------------------------------------------------------
# style transformer
# --------------------------------------------------------------------------------------------


def _draw_stochastic_depth(module, B: int, k: int, device) -> Optional[torch.Tensor]:
    """Per-sample StochasticDepth('row') factors, drawn exactly as torchvision does (ops/stochastic_depth.py: one
    bernoulli_ of shape [B,1,1,1] per call, divided by the survival rate) and in the reference's call order, so a run
    seeded like the reference drops the same residual branches.  Returns fp32 [k, 9, B] or None in eval mode / p = 0."""
    p_enc = float(module.encoder.encoder_stochastic_depth_prob)
    p_dec = float(module.decoder.stochastic_depth.p)
    if not module.training or (p_enc == 0.0 and p_dec == 0.0):
        return None
    out = torch.empty(k, 9, B, dtype=torch.float32, device=device)
    for l in range(k):
        for i in range(9):
            p = p_enc if i < 6 else p_dec
            if p == 0.0:
                out[l, i].fill_(1.0)
                continue
            survival = 1.0 - p
            noise = torch.empty([B, 1, 1, 1], dtype=torch.float32, device=device).bernoulli_(survival)
            if survival > 0.0:
                noise.div_(survival)
            out[l, i].copy_(noise.view(B))
    return out


class _StyleTransformerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, Fc, Fs, k, *params):
        B, H, W, C = Fc.shape
        cfg = module._cfg
        w = packed_weights(module, te.StyleTransformerTrainWeights)
        geo = te._Geo(B, H, W, C, cfg["heads"], cfg["window"][0], cfg["shift"][0])
        sd = _draw_stochastic_depth(module, B, k, Fc.device)
        flags = module.engine_flags()  # the reference's alternate orderings: same kernels, re-sequenced forward and adjoint
        out, tape = te.style_transformer_forward_train(w, Fc.detach().float().contiguous(), Fs.detach().float().contiguous(), k, geo, sd, **flags)
        ctx.module, ctx.w, ctx.geo, ctx.tape, ctx.sd, ctx.flags = module, w, geo, tape, sd, flags
        return out.view(B, H, W, C)

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        names, params = _param_lists(module)
        book = te.st_grad_book({n: p.shape for n, p in zip(names, params)}, g.device)
        if ctx.tape:
            te.style_transformer_backward(ctx.w, ctx.tape, g.float().contiguous(), ctx.geo, ctx.sd, book, workspace_of(module, g.device),
                                          **ctx.flags)
        ctx.tape = None  # free the saved activations
        return (None, None, None, None) + _grads_out(names, params, book)


def style_transformer_apply(module, Fc, Fs, k: int):
    if Fc.requires_grad or Fs.requires_grad:
        raise NotImplementedError("gradients w.r.t. the encoder features are not produced: the Swin encoder is frozen in the "
                                  "reference's training setup (train.py:216-218)")
    _, params = _param_lists(module)
    if k == 0:
        return Fc.float()
    return _StyleTransformerFn.apply(module, Fc, Fs, k, *params)


# --------------------------------------------------------------------------------------------
# CNN decoder
# --------------------------------------------------------------------------------------------


class _CnnDecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        B, C, H, W = x.shape
        w = packed_weights(module, te.CnnDecoderTrainWeights)
        tok = x.detach().permute(0, 2, 3, 1).reshape(B * H * W, C).float().contiguous()  # free when x is a permuted BHWC tensor
        x16 = torch.empty(B * H * W, C, dtype=torch.bfloat16, device=x.device)
        ops.cast_bf16(tok, x16)
        out = torch.empty(B, 3, 8 * H, 8 * W, dtype=torch.float32, device=x.device)
        acts = te.cnn_decoder_forward_train(w, x16, B, H, W, out)
        ctx.module, ctx.w, ctx.acts, ctx.shape, ctx.need_dx = module, w, acts, (B, C, H, W), x.requires_grad
        return out

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        B, C, H, W = ctx.shape
        names, params = _param_lists(module)
        book = te.GradBook([(n, p.shape) for n, p in zip(names, params)], g.device)
        gx16 = te.cnn_decoder_backward(ctx.w, ctx.acts, g.float().contiguous(), B, H, W, book, workspace_of(module, g.device))
        ctx.acts = None
        gx = gx16.float().view(B, H, W, C).permute(0, 3, 1, 2) if ctx.need_dx else None
        return (None, gx) + _grads_out(names, params, book)


def cnn_decoder_apply(module, x):
    _, params = _param_lists(module)
    return _CnnDecoderFn.apply(module, x, *params)


# --------------------------------------------------------------------------------------------
# perceptual loss
# --------------------------------------------------------------------------------------------


class _PerceptualLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, content, style, output, lam):
        fe = module.feature_extractor_model
        w = packed_weights(fe, te.VggTrainWeights)
        wi = packed_weights(fe, engine.VggWeights)
        ws = workspace_of(module, output.device)
        pre = module.__dict__.pop("_prefetched_taps", None)  # content/style taps computed ahead (prefetch_content_style_taps)
        if pre is not None and pre[0] != (content.data_ptr(), style.data_ptr(), tuple(content.shape)):
            pre = None
        out3, saved = te.perceptual_loss_forward_train(
            w, wi, content.detach().float().contiguous(), style.detach().float().contiguous(), output.detach().float().contiguous(),
            lam, module.distance_content == "euclidian_squared", module.distance_style == "euclidian_squared", ws,
            pre=None if pre is None else pre[1])
        ctx.module, ctx.w, ctx.saved, ctx.lam = module, w, saved, lam
        return out3

    @staticmethod
    def backward(ctx, g3):
        # out3 = (content + lam*style, content, style)
        g3 = g3.float()
        coef2 = torch.stack([g3[0] + g3[1], ctx.lam * g3[0] + g3[2]]).contiguous()
        dimg = te.perceptual_loss_backward(ctx.w, ctx.saved, coef2, workspace_of(ctx.module, g3.device))
        ctx.saved = None
        return None, None, None, dimg, None


def prefetch_content_style_taps(module, content, style) -> None:
    """Compute the content / style VGG taps + statistics of the NEXT loss call on the current stream (they do not depend on
    the stylised image) and park them on the loss module; _PerceptualLossFn.forward picks them up if it gets the same tensors."""
    if not (content.is_cuda and style.is_cuda) or content.dtype != torch.float32 or style.dtype != torch.float32:
        return
    if not (content.is_contiguous() and style.is_contiguous()) or content.shape != style.shape:
        return
    fe = module.feature_extractor_model
    with torch.no_grad():
        data = te.content_style_taps(packed_weights(fe, engine.VggWeights), content, style, workspace_of(module, content.device))
    module.__dict__["_prefetched_taps"] = ((content.data_ptr(), style.data_ptr(), tuple(content.shape)), data)


def perceptual_loss_apply(module, content, style, output, lam: float):
    if not (content.is_cuda and style.is_cuda):
        raise RuntimeError("mastermetastyletransfer_b200 runs on sm_100a only: inputs must be CUDA tensors (no CPU fallback)")
    return _PerceptualLossFn.apply(module, content, style, output, lam)
