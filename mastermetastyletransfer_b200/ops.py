"""Thin torch-tensor wrappers over the C ABI (raw device pointers + the current CUDA stream).

PyTorch is used here only for device memory and streams.  Every wrapper checks dtype, device
and contiguity and raises instead of falling back.
"""
from __future__ import annotations

import os
import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import MstAttnBlock, MstGemm, MstLossTap, MstLossTaps, MstMlp, MstWgrad, MstWindowAttn, MstWindowAttnBwd, check

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
GATE_NONE, GATE_RELU, GATE_GELU = 0, 1, 2
A_PLAIN, A_CONV3X3 = 0, 1
PAD_ZERO, PAD_REFLECT = 0, 1

# number of kernels this process has enqueued through the C ABI (bench.py reports it as gpu_launches)
launch_count = 0
_records = None  # list of (kernel family, algorithmic flops, algorithmic bytes, start event, end event) while timing()


class timing:
    """Context manager: CUDA-event bracket around every kernel launched through this module (bench.py roofline pass)."""

    def __enter__(self):
        global _records
        _records = []
        return _records

    def __exit__(self, *exc):
        global _records
        _records = None
        return False


def _launch(name: str, fn, flops: float = 0.0, nbytes: float = 0.0, desc: str = "") -> None:
    """Run one C-ABI call (= one kernel launch), count it, and time it when a timing() context is active."""
    global launch_count
    if _records is None:
        check(fn(), name)
    else:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(fn(), name)
        b.record()
        _records.append((KERNEL_OF.get(name, name), flops, nbytes, a, b, desc))
    launch_count += 1


KERNEL_OF = {"mst_adam_step": "adam_kernel", "mst_reptile_delta": "reptile_kernel", "mst_reptile_apply": "reptile_kernel",
             "mst_gemm": "gemm_tc_kernel", "mst_mlp_fused": "mlp_fused_kernel", "mst_pack_mlp_weights": "pack_kernel", "mst_conv3x3_band": "conv_band_kernel", "mst_conv3x3_rows": "conv_rows_kernel", "mst_conv3x3_cm": "conv_cm_kernel", "mst_window_attention": "window_attn_kernel", "mst_window_attention:core": "attn_core_kernel", "mst_attn_block": "attn_fused_kernel", "mst_pack_attn_qkv": "pack_kernel", "mst_layernorm": "layernorm_kernel",
             "mst_patch_merge_layernorm": "layernorm_kernel", "mst_instnorm_stats": "instnorm_stats_kernel",
             "mst_instnorm_apply": "instnorm_apply_kernel", "mst_instnorm": "instnorm_fused_kernel", "mst_jointnorm_stats": "jointnorm_stats_kernel", "mst_softmax_rows": "softmax_rows_kernel", "mst_pack_bf16_matrix": "pack_kernel", "mst_patch_embed": "patch_embed_kernel",
             "mst_cast_bf16": "cast_bf16_kernel", "mst_images_u8_to_nchw": "images_u8_to_nchw_kernel", "mst_images_nchw_to_u8": "images_nchw_to_u8_kernel", "mst_resize_crop_normalize": "resize_crop_normalize_kernel", "mst_upsample2x_nhwc": "upsample2x_kernel", "mst_pack_linear_weight": "pack_kernel", "mst_pack_conv3x3_weight": "pack_kernel",
             "mst_window_maps": "window_maps_kernel", "mst_conv3x3_first": "conv3x3_first_kernel",
             "mst_maxpool2x2": "maxpool2x2_kernel", "mst_bn_relu": "bn_relu_kernel", "mst_tap_stats": "tap_stats_kernel", "mst_content_term": "content_term_kernel",
             "mst_loss_finalize": "loss_finalize_kernel", "mst_sim_prepare": "sim_prepare_kernels", "mst_sim_tiles": "sim_tile_kernel", "mst_sim_finalize": "sim_finalize_kernel", "mst_wgrad": "wgrad_tc_kernel", "mst_colsum": "colsum_kernel",
             "mst_window_attention_bwd": "window_attn_bwd_kernel", "mst_layernorm_bwd": "layernorm_bwd_kernel",
             "mst_instnorm_bwd_stats": "instnorm_bwd_stats_kernel", "mst_instnorm_bwd_apply": "instnorm_bwd_apply_kernel",
             "mst_blend_bwd": "blend_bwd_kernel", "mst_add_cast": "add_cast_kernel", "mst_token_map_copy": "token_map_kernel", "mst_reflect_fold": "reflect_fold_kernel",
             "mst_maxpool2x2_bwd": "maxpool2x2_bwd_kernel", "mst_nchw3_to_nhwc8": "nchw3_to_nhwc8_kernel",
             "mst_loss_bwd_stats": "loss_bwd_stats_kernel", "mst_loss_bwd_apply": "loss_bwd_apply_kernel"}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor], dtype=None, name="tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def n_pad_of(n: int) -> int:
    return round_up(n, 16)


class PackedMatrix:
    """bf16 [n_pad, k_pad] tensor-core operand plus the fp32 bias padded to n_pad."""

    __slots__ = ("w", "bias", "N", "K", "n_pad", "k_pad")

    def __init__(self, w, bias, N, K, n_pad, k_pad):
        self.w, self.bias, self.N, self.K, self.n_pad, self.k_pad = w, bias, N, K, n_pad, k_pad


def _pad_bias(bias: Optional[torch.Tensor], n_pad: int, device) -> Optional[torch.Tensor]:
    if bias is None:
        return None
    b = bias.detach().reshape(-1)
    if b.numel() == n_pad and b.dtype == torch.float32 and b.is_contiguous() and b.device == torch.device(device):
        return b  # no padding needed: the kernels read the live parameter (re-packing a training step costs no fill + copy per bias)
    out = torch.zeros(n_pad, dtype=torch.float32, device=device)
    out[: bias.numel()].copy_(bias.detach().reshape(-1))
    return out


def pack_linear(weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> PackedMatrix:
    """nn.Linear weight [N,K] fp32 -> PackedMatrix (several weights may be concatenated along N first)."""
    w = weight.detach().contiguous()
    N, K = w.shape
    n_pad, k_pad = n_pad_of(N), round_up(K, 64)
    dst = torch.empty(n_pad, k_pad, dtype=torch.bfloat16, device=w.device)
    _launch("mst_pack_linear_weight", lambda: _lib.lib().mst_pack_linear_weight(_ptr(w, torch.float32, "weight"), N, K, dst.data_ptr(), n_pad, k_pad, _stream()))
    return PackedMatrix(dst, _pad_bias(bias, n_pad, w.device), N, K, n_pad, k_pad)


def pack_bf16_matrix(src: torch.Tensor, N: int, K: int, ld: int, trans: bool = False, dst: Optional[torch.Tensor] = None) -> PackedMatrix:
    """A bf16 device matrix (an activation) as the packed B operand of gemm(): W[n][k] = src[n*ld+k], or src[k*ld+n] if trans."""
    n_pad, k_pad = n_pad_of(N), round_up(K, 64)
    if dst is None:
        dst = torch.empty(n_pad, k_pad, dtype=torch.bfloat16, device=src.device)
    _launch("mst_pack_bf16_matrix", lambda: _lib.lib().mst_pack_bf16_matrix(_ptr(src, torch.bfloat16, "src"), N, K, ld, int(trans),
                                                                           _ptr(dst, torch.bfloat16, "dst"), n_pad, k_pad, _stream()))
    return PackedMatrix(dst, None, N, K, n_pad, k_pad)


def softmax_rows(S: torch.Tensor, P: torch.Tensor, rows: int, n: int, scale: float) -> None:
    _launch("mst_softmax_rows", lambda: _lib.lib().mst_softmax_rows(_ptr(S, torch.float32, "S"), _ptr(P, torch.bfloat16, "P"), rows, n,
                                                                   float(scale), _stream()), nbytes=10.0 * rows * n)


def pack_conv3x3(weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> PackedMatrix:
    """nn.Conv2d weight [N,Cin,3,3] fp32 -> PackedMatrix with k = (ky*3+kx)*Cin + ci."""
    w = weight.detach().contiguous()
    N, Cin, kh, kw = w.shape
    if (kh, kw) != (3, 3):
        raise ValueError("pack_conv3x3: expected a 3x3 kernel")
    n_pad, k_pad = n_pad_of(N), round_up(9 * Cin, 64)
    dst = torch.empty(n_pad, k_pad, dtype=torch.bfloat16, device=w.device)
    _launch("mst_pack_conv3x3_weight", lambda: _lib.lib().mst_pack_conv3x3_weight(_ptr(w, torch.float32, "weight"), N, Cin, dst.data_ptr(), n_pad, k_pad, _stream()))
    return PackedMatrix(dst, _pad_bias(bias, n_pad, w.device), N, 9 * Cin, n_pad, k_pad)


def gemm(A: torch.Tensor, pm: PackedMatrix, M: int, *, lda: Optional[int] = None, act: int = ACT_NONE,
         res: Optional[torch.Tensor] = None, mul: Optional[torch.Tensor] = None,
         out_f32: Optional[torch.Tensor] = None, out_bf16: Optional[torch.Tensor] = None,
         ld_out32: Optional[int] = None, ld_out16: Optional[int] = None, ld_res: Optional[int] = None,
         conv: Optional[dict] = None, gate: Optional[torch.Tensor] = None, gate_mode: int = GATE_NONE,
         add16: Optional[torch.Tensor] = None, ld_gate: Optional[int] = None, out_pre16: Optional[torch.Tensor] = None,
         row_scale: Optional[torch.Tensor] = None, rows_per_scale: int = 0, out_u8: Optional[torch.Tensor] = None) -> None:
    """acc = A . Wt^T ; x = act(acc + bias) ; x = res*mul + x | res + x ; store (see include/mst_b200.h).
    out_u8 (3x3 convolutions the row-streaming kernel takes -- rows_supported): a uint8 [B,H,W,n_real] image receiving
    (uint8) clip(x * 255, 0, 255) instead of an fp32 / bf16 result (MST_OUT_IMAGE_U8)."""
    if out_u8 is not None:
        if conv is None or out_f32 is not None or out_bf16 is not None or res is not None or not _use_rows(dict(conv, impl="rows"), pm.n_pad):
            raise ValueError("gemm: out_u8 is the only output of a 3x3 convolution on the row-streaming kernel")
        n_real = conv.get("n_real", pm.N)
        if out_u8.dtype != torch.uint8 or out_u8.numel() != M * n_real or not out_u8.is_contiguous():
            raise ValueError("gemm: out_u8 must be a contiguous uint8 [B,H,W,n_real] image")
        conv = dict(conv, out_nchw=2, impl="rows")
    g = MstGemm()
    g.A = _ptr(A, torch.bfloat16, "A")
    g.Wt = pm.w.data_ptr()
    g.bias = _ptr(pm.bias, torch.float32, "bias")
    g.res = _ptr(res, torch.float32, "res")
    g.mul = _ptr(mul, torch.float32, "mul")
    g.out_f32 = _ptr(out_f32, torch.float32, "out_f32") if out_u8 is None else _ptr(out_u8, torch.uint8, "out_u8")
    g.out_bf16 = _ptr(out_bf16, torch.bfloat16, "out_bf16")
    N = pm.n_pad
    g.M, g.N, g.K, g.k_pad = M, N, pm.K, pm.k_pad
    g.lda = pm.K if lda is None else lda
    g.ld_res = N if ld_res is None else ld_res
    g.ld_out32 = N if ld_out32 is None else ld_out32
    g.ld_out16 = N if ld_out16 is None else ld_out16
    g.act = act
    g.gate, g.add16 = _ptr(gate, torch.bfloat16, "gate"), _ptr(add16, torch.bfloat16, "add16")
    g.out_pre16 = _ptr(out_pre16, torch.bfloat16, "out_pre16")
    g.row_scale, g.rows_per_scale = _ptr(row_scale, torch.float32, "row_scale"), rows_per_scale
    g.gate_mode, g.ld_gate = gate_mode, (N if ld_gate is None else ld_gate)
    training_ext = gate is not None or add16 is not None or out_pre16 is not None or row_scale is not None
    if conv is None:
        g.a_mode = A_PLAIN
    else:
        g.a_mode = A_CONV3X3
        g.H, g.W, g.Cin = conv["H"], conv["W"], conv["Cin"]
        g.pad_mode, g.upsample = conv.get("pad_mode", PAD_ZERO), int(conv.get("upsample", False))
        g.out_nchw, g.n_real = int(conv.get("out_nchw", False)), conv.get("n_real", pm.N)
        g.conv_full = int(conv.get("full", False))
        training_ext = training_ext or bool(g.conv_full)
    desc = f"M={M} N={N} K={pm.K} conv={conv is not None} act={act} res={res is not None} o32={out_f32 is not None} o16={out_bf16 is not None}"
    flops = 2.0 * M * min(N, pm.N) * pm.K
    plain16 = out_f32 is None and res is None and mul is None and out_bf16 is not None
    if conv is not None and not training_ext and plain16 and _use_cm(conv, N, act):
        _launch("mst_conv3x3_cm", lambda: _lib.lib().mst_conv3x3_cm(C.byref(g), _stream()), flops=flops, desc=desc + " cm")
    elif conv is not None and not training_ext and _use_rows(conv, N):
        _launch("mst_conv3x3_rows", lambda: _lib.lib().mst_conv3x3_rows(C.byref(g), _stream()), flops=flops, desc=desc + " rows")
    elif conv is not None and not training_ext and _use_band(conv, N):
        _launch("mst_conv3x3_band", lambda: _lib.lib().mst_conv3x3_band(C.byref(g), _stream()), flops=flops, desc=desc + " band")
    else:
        _launch("mst_gemm", lambda: _lib.lib().mst_gemm(C.byref(g), _stream()), flops=flops, desc=desc)


BAND_MAX_CIN = 64  # measured: from Cin = 128 up the gathered implicit GEMM is faster (tools/gemm_bench.py band)


def band_supported(n: int, cin: int, H: int, W: int) -> bool:
    return cin % 16 == 0 and bool(_lib.lib().mst_conv3x3_band_supported(n_pad_of(n), cin, H, W))


def rows_supported(n: int, cin: int, H: int, W: int) -> bool:
    return bool(_lib.lib().mst_conv3x3_rows_supported(n_pad_of(n), cin, H, W))


def cm_supported(n: int, cin: int, H: int, W: int) -> bool:
    return bool(_lib.lib().mst_conv3x3_cm_supported(n_pad_of(n), cin, H, W))


def _use_cm(conv: dict, n_pad: int, act: int) -> bool:
    """The channel-major kernel for the wide layers (conv_cm.cu): 'auto' takes it whenever the shape fits."""
    impl = conv.get("impl", "auto")
    if impl not in ("auto", "cm"):
        return False
    ok = (not conv.get("out_nchw", False) and act in (ACT_NONE, ACT_RELU)
          and bool(_lib.lib().mst_conv3x3_cm_supported(n_pad, conv["Cin"], conv["H"], conv["W"])))
    if impl == "cm" and not ok:
        raise ValueError("conv3x3 channel-major kernel does not support this shape")
    return ok


def _use_rows(conv: dict, n_pad: int) -> bool:
    impl = conv.get("impl", "auto")
    if impl not in ("auto", "rows"):
        return False
    ok = bool(_lib.lib().mst_conv3x3_rows_supported(n_pad, conv["Cin"], conv["H"], conv["W"]))
    if impl == "rows" and not ok:
        raise ValueError("conv3x3 row-streaming kernel does not support this shape")
    return ok


def _use_band(conv: dict, n_pad: int) -> bool:
    impl = conv.get("impl", "auto")
    if impl == "gather":
        return False
    ok = conv["Cin"] % 16 == 0 and bool(_lib.lib().mst_conv3x3_band_supported(n_pad, conv["Cin"], conv["H"], conv["W"]))
    if impl == "band":
        if not ok:
            raise ValueError("conv3x3 band kernel does not support this shape")
        return True
    return ok and conv["Cin"] <= BAND_MAX_CIN


def window_attention(q, k, v, out, bias_table, B, H, W, heads, ws, shift, ldq, ldk, ldv, ldo,
                     v2=None, out2=None, pad_q=None, pad_k=None, pad_v=None, pad_v2=None, pad_k_per_image: bool = False) -> None:
    a = MstWindowAttn()
    a.q, a.k, a.v = _ptr(q, torch.bfloat16, "q"), _ptr(k, torch.bfloat16, "k"), _ptr(v, torch.bfloat16, "v")
    a.v2, a.out, a.out2 = _ptr(v2, torch.bfloat16, "v2"), _ptr(out, torch.bfloat16, "out"), _ptr(out2, torch.bfloat16, "out2")
    a.bias_table = _ptr(bias_table, torch.float32, "bias_table")
    a.pad_q, a.pad_k = _ptr(pad_q, torch.float32, "pad_q"), _ptr(pad_k, torch.float32, "pad_k")
    a.pad_v, a.pad_v2 = _ptr(pad_v, torch.float32, "pad_v"), _ptr(pad_v2, torch.float32, "pad_v2")
    a.B, a.H, a.W, a.heads, a.ws, a.shift = B, H, W, heads, ws, shift
    a.ldq, a.ldk, a.ldv, a.ldo = ldq, ldk, ldv, ldo
    a.pad_k_stride = heads * 32 if pad_k_per_image else 0
    n_tok = ws * ws
    n_win = B * (-(-H // ws)) * (-(-W // ws))
    # which kernel the library picks (csrc/attn_core.cu: attn_core_try): the dual passes on maps the windows tile run on tcgen05
    core = (v2 is not None and ws in (7, 8) and H % ws == 0 and W % ws == 0 and heads % 2 == 0 and (ldq | ldk | ldv | ldo) % 16 == 0
            and all(t.data_ptr() % 32 == 0 for t in (q, k, v, v2, out, out2)) and os.environ.get("MST_ATTN_CORE", "1") != "0")
    _launch("mst_window_attention:core" if core else "mst_window_attention", lambda: _lib.lib().mst_window_attention(C.byref(a), _stream()),
            flops=2.0 * n_win * heads * n_tok * n_tok * 32 * (3 if v2 is not None else 2),
            desc=f"B={B} H={H} heads={heads} ws={ws} shift={shift} dual={v2 is not None}")


class PackedAttnQkv:
    """Wq | Wk | Wv of a self-attention, packed per head pair for the fused attention-block kernel (csrc/attn_fused.cu)."""

    __slots__ = ("w", "b", "C", "heads")

    def __init__(self, w, b, C_, heads):
        self.w, self.b, self.C, self.heads = w, b, C_, heads


def pack_attn_qkv(wq, wk, wv, bq, bk, bv, heads: int) -> PackedAttnQkv:
    """Three nn.Linear weights [C,C] (slices of a fused [3C,C] qkv weight work) + biases [C] -> PackedAttnQkv."""
    wq, wk, wv = (t.detach().contiguous() for t in (wq, wk, wv))
    bq, bk, bv = (None if t is None else t.detach().contiguous() for t in (bq, bk, bv))
    Cdim = int(wq.shape[0])
    nbytes = _lib.lib().mst_attn_qkv_packed_bytes(Cdim, heads)
    if nbytes == 0 or wq.shape != (Cdim, Cdim) or wk.shape != wq.shape or wv.shape != wq.shape:
        raise ValueError("pack_attn_qkv: needs square [C,C] weights with C in {128, 256} and head_dim 32")
    dst_w = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=wq.device)
    dst_b = torch.empty(3 * Cdim, dtype=torch.float32, device=wq.device)
    _launch("mst_pack_attn_qkv", lambda: _lib.lib().mst_pack_attn_qkv(
        _ptr(wq, torch.float32, "wq"), _ptr(wk, torch.float32, "wk"), _ptr(wv, torch.float32, "wv"), _ptr(bq, torch.float32, "bq"),
        _ptr(bk, torch.float32, "bk"), _ptr(bv, torch.float32, "bv"), dst_w.data_ptr(), dst_b.data_ptr(), Cdim, heads, _stream()))
    return PackedAttnQkv(dst_w, dst_b, Cdim, heads)


def attn_block(x16, pk: PackedAttnQkv, bias_table, out, B, H, W, ws, shift, ldx=None, ldo=None, dbg_qkv=None) -> None:
    """out = window_attention(x Wq^T + bq, x Wk^T + bk, x Wv^T + bv) before the output projection: the three projections and
    the shifted-window attention core in ONE kernel (q, k, v stay on chip)."""
    a = MstAttnBlock()
    a.x, a.wqkv, a.bqkv = _ptr(x16, torch.bfloat16, "x"), _ptr(pk.w, torch.bfloat16, "wqkv"), _ptr(pk.b, torch.float32, "bqkv")
    a.bias_table, a.out = _ptr(bias_table, torch.float32, "bias_table"), _ptr(out, torch.bfloat16, "out")
    a.dbg_qkv = _ptr(dbg_qkv, torch.bfloat16, "dbg_qkv")
    a.B, a.H, a.W, a.C, a.heads, a.ws, a.shift = B, H, W, pk.C, pk.heads, ws, shift
    a.ldx, a.ldo = ldx or pk.C, ldo or pk.C
    T = B * H * W
    # reference-algorithm FLOPs: three linears on the real tokens + QK^T and PV over ws*ws keys per token
    _launch("mst_attn_block", lambda: _lib.lib().mst_attn_block(C.byref(a), _stream()),
            flops=6.0 * T * pk.C * pk.C + 4.0 * T * ws * ws * pk.C, nbytes=4.0 * T * pk.C,
            desc=f"B={B} H={H} C={pk.C} ws={ws} shift={shift}")


def window_maps(H: int, W: int, ws: int, shift: int, device="cuda"):
    """(gather [nW,N] int32, labels [nW,N] int32, relidx [N*N] int32) as the attention kernel computes them."""
    Hp, Wp = H + (ws - H % ws) % ws, W + (ws - W % ws) % ws
    nW, N = (Hp // ws) * (Wp // ws), ws * ws
    gather = torch.empty(nW, N, dtype=torch.int32, device=device)
    labels = torch.empty(nW, N, dtype=torch.int32, device=device)
    relidx = torch.empty(N * N, dtype=torch.int32, device=device)
    _launch("mst_window_maps", lambda: _lib.lib().mst_window_maps(H, W, ws, shift, gather.data_ptr(), labels.data_ptr(), relidx.data_ptr(), _stream()))
    return gather, labels, relidx


def layernorm(x, gamma, beta, y, rows, Cdim) -> None:
    _launch("mst_layernorm", lambda: _lib.lib().mst_layernorm(_ptr(x, torch.float32, "x"), _ptr(gamma, torch.float32, "gamma"),
                                   _ptr(beta, torch.float32, "beta"), _ptr(y, torch.bfloat16, "y"), rows, Cdim, _stream()),
            nbytes=6.0 * rows * Cdim)


def patch_merge_layernorm(x, gamma, beta, y, B, H, W, Cdim) -> None:
    _launch("mst_patch_merge_layernorm", lambda: _lib.lib().mst_patch_merge_layernorm(_ptr(x, torch.float32, "x"), _ptr(gamma, torch.float32, "gamma"),
                                               _ptr(beta, torch.float32, "beta"), _ptr(y, torch.bfloat16, "y"), B, H, W, Cdim,
                                               _stream()), nbytes=6.0 * B * H * W * Cdim)


def instnorm_stats(x, mean, rstd, B, T, Cdim, twice=False, gamma=None) -> None:
    """gamma [C]: the affine InstanceNorm's weight, folded into rstd (see mst_instnorm_stats_affine); pass the bias to
    instnorm_apply(beta=...)."""
    if gamma is None:
        fn = lambda: _lib.lib().mst_instnorm_stats(_ptr(x, torch.float32, "x"), _ptr(mean, torch.float32, "mean"),
                                                   _ptr(rstd, torch.float32, "rstd"), B, T, Cdim, int(twice), _stream())
    else:
        fn = lambda: _lib.lib().mst_instnorm_stats_affine(_ptr(x, torch.float32, "x"), _ptr(mean, torch.float32, "mean"),
                                                          _ptr(rstd, torch.float32, "rstd"), B, T, Cdim, int(twice), 0, None, None,
                                                          _ptr(gamma, torch.float32, "gamma"), None, _stream())
    _launch("mst_instnorm_stats", fn, nbytes=4.0 * B * T * Cdim)


def instnorm_stats_padded(x, mean, rstd, B, T, Cdim, n_pad, pad_val, pad_norm=None, gamma=None, beta=None) -> None:
    """InstanceNorm statistics over a map with n_pad extra tokens of value pad_val[c] (window-padded map after a Linear);
    pad_norm [B,C] receives the normalised padding value (affine: gamma folded into rstd, beta added to pad_norm)."""
    if gamma is None and beta is None:
        fn = lambda: _lib.lib().mst_instnorm_stats_padded(
            _ptr(x, torch.float32, "x"), _ptr(mean, torch.float32, "mean"), _ptr(rstd, torch.float32, "rstd"), B, T, Cdim, n_pad,
            _ptr(pad_val, torch.float32, "pad_val"), _ptr(pad_norm, torch.float32, "pad_norm"), _stream())
    else:
        fn = lambda: _lib.lib().mst_instnorm_stats_affine(
            _ptr(x, torch.float32, "x"), _ptr(mean, torch.float32, "mean"), _ptr(rstd, torch.float32, "rstd"), B, T, Cdim, 0, n_pad,
            _ptr(pad_val, torch.float32, "pad_val"), _ptr(pad_norm, torch.float32, "pad_norm"), _ptr(gamma, torch.float32, "gamma"),
            _ptr(beta, torch.float32, "beta"), _stream())
    _launch("mst_instnorm_stats", fn, nbytes=4.0 * B * T * Cdim)


def jointnorm_stats(x, mean, rstd, B, T, Cdim) -> None:
    """mean / rstd of every image over (T, C) jointly, replicated into [B, C] (the regular-MHA variant's InstanceNorm quirk)."""
    _launch("mst_jointnorm_stats", lambda: _lib.lib().mst_jointnorm_stats(_ptr(x, torch.float32, "x"), _ptr(mean, torch.float32, "mean"),
                                                                         _ptr(rstd, torch.float32, "rstd"), B, T, Cdim, _stream()),
            nbytes=4.0 * B * T * Cdim)


def instnorm_apply(x, mean, rstd, B, T, Cdim, y16=None, y32=None, beta=None) -> None:
    if beta is None:
        fn = lambda: _lib.lib().mst_instnorm_apply(_ptr(x, torch.float32, "x"), _ptr(mean, torch.float32, "mean"),
                                                   _ptr(rstd, torch.float32, "rstd"), _ptr(y16, torch.bfloat16, "y16"),
                                                   _ptr(y32, torch.float32, "y32"), B, T, Cdim, _stream())
    else:
        fn = lambda: _lib.lib().mst_instnorm_apply_affine(_ptr(x, torch.float32, "x"), _ptr(mean, torch.float32, "mean"),
                                                          _ptr(rstd, torch.float32, "rstd"), _ptr(beta, torch.float32, "beta"),
                                                          _ptr(y16, torch.bfloat16, "y16"), _ptr(y32, torch.float32, "y32"), B, T, Cdim, _stream())
    _launch("mst_instnorm_apply", fn,
            nbytes=B * T * Cdim * (4.0 + (2.0 if y16 is not None else 0.0) + (4.0 if y32 is not None else 0.0)))


def instnorm(x, mean, rstd, y16, B, T, Cdim, twice=False, gamma=None, beta=None, n_pad=0, pad_val=None, pad_norm=None) -> None:
    """InstanceNorm statistics (instnorm_stats / instnorm_stats_padded) + application to bf16 (instnorm_apply) of one tensor: one kernel
    that reads x once when an (image, 32-channel) slice fits in shared memory (csrc/instnorm_fused.cu; bit-identical), else the two."""
    if not _lib.lib().mst_instnorm_fused_supported(T, Cdim) or os.environ.get("MST_INSTNORM_FUSED", "1") == "0":
        if n_pad > 0 or pad_norm is not None:
            instnorm_stats_padded(x, mean, rstd, B, T, Cdim, n_pad, pad_val, pad_norm=pad_norm, gamma=gamma, beta=beta)
        else:
            instnorm_stats(x, mean, rstd, B, T, Cdim, twice=twice, gamma=gamma)
        instnorm_apply(x, mean, rstd, B, T, Cdim, y16=y16, beta=beta)
        return
    _launch("mst_instnorm", lambda: _lib.lib().mst_instnorm(
        _ptr(x, torch.float32, "x"), _ptr(mean, torch.float32, "mean"), _ptr(rstd, torch.float32, "rstd"), _ptr(y16, torch.bfloat16, "y16"),
        B, T, Cdim, int(twice), int(n_pad), _ptr(pad_val, torch.float32, "pad_val"), _ptr(pad_norm, torch.float32, "pad_norm"),
        _ptr(gamma, torch.float32, "gamma"), _ptr(beta, torch.float32, "beta"), _stream()), nbytes=6.0 * B * T * Cdim)


def patch_embed_u8_supported(S: int) -> bool:
    return bool(_lib.lib().mst_patch_embed_ln_u8_supported(int(S)))


def patch_embed(img, w, b, gamma, beta, x, B, S, gamma1=None, beta1=None, y16=None, exact: bool = False, u8_mean=None, u8_std=None) -> None:
    """tv swin features[0] (conv 4x4/s4 + LayerNorm).  With y16 the first block's norm1 (gamma1/beta1) is fused and its bf16
    output written to y16.  exact=True forces the fp32 SIMT kernel (the default tensor-core path rounds the weights to bf16).
    A uint8 [B,S,S,3] img is read directly, ToTensor + Normalize(u8_mean, u8_std) applied in the image loads (None: / 255 only)
    -- the same values images_u8_to_nchw would have written, the same kernel after that."""
    if img.dtype == torch.uint8:
        if tuple(img.shape) != (B, S, S, 3) or exact:
            raise ValueError("patch_embed: uint8 images are [B,S,S,3] and run on the tensor-core kernel only")
        m = C.cast((C.c_float * 3)(*u8_mean), C.c_void_p) if u8_mean is not None else None
        sd = C.cast((C.c_float * 3)(*u8_std), C.c_void_p) if u8_mean is not None else None
        _launch("mst_patch_embed", lambda: _lib.lib().mst_patch_embed_ln_u8(
            _ptr(img, torch.uint8, "img"), m, sd, _ptr(w, torch.float32, "w"), _ptr(b, torch.float32, "b"),
            _ptr(gamma, torch.float32, "gamma"), _ptr(beta, torch.float32, "beta"), _ptr(x, torch.float32, "x"),
            _ptr(gamma1, torch.float32, "gamma1"), _ptr(beta1, torch.float32, "beta1"), _ptr(y16, torch.bfloat16, "y16"),
            B, S, _stream()),
                nbytes=1.0 * B * S * S * 3 + (4.0 + (2.0 if y16 is not None else 0.0)) * B * (S // 4) ** 2 * 128)
        return
    _launch("mst_patch_embed", lambda: _lib.lib().mst_patch_embed_ln(
        _ptr(img, torch.float32, "img"), _ptr(w, torch.float32, "w"), _ptr(b, torch.float32, "b"),
        _ptr(gamma, torch.float32, "gamma"), _ptr(beta, torch.float32, "beta"), _ptr(x, torch.float32, "x"),
        _ptr(gamma1, torch.float32, "gamma1"), _ptr(beta1, torch.float32, "beta1"), _ptr(y16, torch.bfloat16, "y16"),
        B, S, int(exact), _stream()),
            nbytes=4.0 * B * S * S * 3 + (4.0 + (2.0 if y16 is not None else 0.0)) * B * (S // 4) ** 2 * 128)


def upsample2x_nhwc(x, y, B, H, W, C_) -> None:
    """nearest x2 upsample of a bf16 NHWC tensor [B,H,W,C] -> [B,2H,2W,C]."""
    _launch("mst_upsample2x_nhwc", lambda: _lib.lib().mst_upsample2x_nhwc(_ptr(x, torch.bfloat16, "x"), _ptr(y, torch.bfloat16, "y"), B, H, W, C_, _stream()),
            nbytes=2.0 * 5 * B * H * W * C_)


def cast_bf16(x, y) -> None:
    _launch("mst_cast_bf16", lambda: _lib.lib().mst_cast_bf16(_ptr(x, torch.float32, "x"), _ptr(y, torch.bfloat16, "y"), x.numel(), _stream()),
            nbytes=6.0 * x.numel())


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def images_u8_to_nchw(src_u8, dst32, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> None:
    """uint8 [B,H,W,3] -> fp32 [B,3,H,W]: transforms.ToTensor() + transforms.Normalize(mean, std) (test_model.py:39-48); mean=None: /255 only."""
    B, H, W, ch = src_u8.shape
    if ch != 3 or tuple(dst32.shape) != (B, 3, H, W):
        raise ValueError("images_u8_to_nchw: src [B,H,W,3] uint8, dst [B,3,H,W] fp32")
    m = C.cast((C.c_float * 3)(*mean), C.c_void_p) if mean is not None else None
    sd = C.cast((C.c_float * 3)(*std), C.c_void_p) if mean is not None else None
    _launch("mst_images_u8_to_nchw", lambda: _lib.lib().mst_images_u8_to_nchw(_ptr(src_u8, torch.uint8, "src"), _ptr(dst32, torch.float32, "dst"),
                                                                              B, H, W, m, sd, _stream()), nbytes=15.0 * B * H * W)


def resize_crop_normalize(img_u8, coeffs_x, coeffs_y, top: int, left: int, out32, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> None:
    """Decoded uint8 [H,W,3] image -> fp32 [3,ch,cw]: the (top, left) crop of Pillow's bilinear resize, ToTensor, Normalize
    (codes/get_dataloader.py:30-36).  coeffs_* = (xmin int32 [out], count int32 [out], k int32 [out, ksize]) device tensors from
    data.pil_resize_coeffs for the image's width / height."""
    H, W, ch3 = img_u8.shape
    if ch3 != 3 or out32.dim() != 3 or out32.shape[0] != 3 or not out32.is_contiguous() or not img_u8.is_contiguous():
        raise ValueError("resize_crop_normalize: img [H,W,3] uint8 contiguous, out [3,ch,cw] fp32 contiguous")
    (xm, xc, xk), (ym, yc, yk) = coeffs_x, coeffs_y
    chh, cw = int(out32.shape[1]), int(out32.shape[2])
    if top + chh > ym.numel() or left + cw > xm.numel():
        raise ValueError("resize_crop_normalize: the crop window leaves the resized image")
    m = C.cast((C.c_float * 3)(*mean), C.c_void_p) if mean is not None else None
    sd = C.cast((C.c_float * 3)(*std), C.c_void_p) if mean is not None else None
    _launch("mst_resize_crop_normalize", lambda: _lib.lib().mst_resize_crop_normalize(
        _ptr(img_u8, torch.uint8, "img"), H, W, _ptr(xm, torch.int32, "xmin"), _ptr(xc, torch.int32, "xcnt"), _ptr(xk, torch.int32, "xk"),
        int(xk.shape[1]), _ptr(ym, torch.int32, "ymin"), _ptr(yc, torch.int32, "ycnt"), _ptr(yk, torch.int32, "yk"), int(yk.shape[1]),
        int(top), int(left), chh, cw, m, sd, _ptr(out32, torch.float32, "out"), _stream()), nbytes=3.0 * H * W + 12.0 * chh * cw)


def images_nchw_to_u8(src32, dst_u8) -> None:
    """fp32 [B,3,H,W] -> uint8 [B,H,W,3]: np.clip(x * 255, 0, 255).astype(np.uint8) (test_model.py:207)."""
    B, ch, H, W = src32.shape
    if ch != 3 or tuple(dst_u8.shape) != (B, H, W, 3):
        raise ValueError("images_nchw_to_u8: src [B,3,H,W] fp32, dst [B,H,W,3] uint8")
    _launch("mst_images_nchw_to_u8", lambda: _lib.lib().mst_images_nchw_to_u8(_ptr(src32, torch.float32, "src"), _ptr(dst_u8, torch.uint8, "dst"),
                                                                              B, H, W, _stream()), nbytes=15.0 * B * H * W)


def conv3x3_first(img, w, b, out, B, H, W, relu=True) -> None:
    _launch("mst_conv3x3_first", lambda: _lib.lib().mst_conv3x3_first(_ptr(img, torch.float32, "img"), _ptr(w, torch.float32, "w"),
                                                                       _ptr(b, torch.float32, "b"), _ptr(out, torch.bfloat16, "out"),
                                                                       B, H, W, int(relu), _stream()),
            flops=2.0 * B * H * W * 27 * 64, nbytes=B * H * W * (12.0 + 128.0))


def maxpool2x2(x, y, B, H, W, Cdim) -> None:
    _launch("mst_maxpool2x2", lambda: _lib.lib().mst_maxpool2x2(_ptr(x, torch.bfloat16, "x"), _ptr(y, torch.bfloat16, "y"), B, H, W, Cdim, _stream()),
            nbytes=2.5 * B * H * W * Cdim)


_tap_scratch = {}


def bn_relu(y, mean, var, gamma, beta, eps: float, M: int, Cdim: int, relu: bool = True, x32=None, var_is_rstd: bool = False) -> None:
    """BatchNorm2d (+ ReLU) -> bf16 y [M, C]; input = x32 (fp32 conv output) or y itself in place; mean / var [C] are the batch
    (train) or running (eval) statistics (var_is_rstd: `var` holds 1/sqrt(var + eps) already)."""
    _launch("mst_bn_relu", lambda: _lib.lib().mst_bn_relu(_ptr(x32, torch.float32, "x32"), _ptr(y, torch.bfloat16, "y"), _ptr(mean, torch.float32, "mean"),
                                                         _ptr(var, torch.float32, "var"), _ptr(gamma, torch.float32, "gamma"),
                                                         _ptr(beta, torch.float32, "beta"), float(eps), M, Cdim, int(relu), int(var_is_rstd), _stream()),
            nbytes=(6.0 if x32 is not None else 4.0) * M * Cdim)


def tap_stats(x, mean, var, B, T, Cdim, scratch=None) -> None:
    """scratch: optional fp32 workspace of mst_tap_stats_scratch_floats(B,T,C) floats; a per-device one is kept otherwise."""
    need = _lib.lib().mst_tap_stats_scratch_floats(B, T, Cdim)
    if scratch is None:
        scratch = _tap_scratch.get(x.device)
        if scratch is None or scratch.numel() < need:
            scratch = torch.empty(max(need, 1 << 20), dtype=torch.float32, device=x.device)
            _tap_scratch[x.device] = scratch
    _launch("mst_tap_stats", lambda: _lib.lib().mst_tap_stats(_ptr(x, torch.bfloat16, "x"), _ptr(mean, torch.float32, "mean"),
                                                               _ptr(var, torch.float32, "var"), B, T, Cdim,
                                                               _ptr(scratch, torch.float32, "scratch"), scratch.numel(), _stream()),
            nbytes=2.0 * B * T * Cdim)


def content_term(fc, fo, mean_c, var_c, mean_o, var_o, B, T, Cdim, squared, partials) -> None:
    _launch("mst_content_term", lambda: _lib.lib().mst_content_term(
        _ptr(fc, torch.bfloat16, "fc"), _ptr(fo, torch.bfloat16, "fo"), _ptr(mean_c, torch.float32, "mean_c"),
        _ptr(var_c, torch.float32, "var_c"), _ptr(mean_o, torch.float32, "mean_o"), _ptr(var_o, torch.float32, "var_o"),
        B, T, Cdim, int(squared), _ptr(partials, torch.float32, "partials"), partials.numel(), _stream()),
        nbytes=4.0 * B * T * Cdim)


def sim_prepare(feat, B, N, Cdim, ahat, svec, inv_cs) -> None:
    """Row-normalised features (bf16), per-image sum vector and the reciprocal column sums of the cosine self-similarity map."""
    _launch("mst_sim_prepare", lambda: _lib.lib().mst_sim_prepare(_ptr(feat, torch.bfloat16, "feat"), B, N, Cdim, _ptr(ahat, torch.bfloat16, "ahat"),
                                                                 _ptr(svec, torch.float32, "svec"), _ptr(inv_cs, torch.float32, "inv_cs"), _stream()),
            nbytes=6.0 * B * N * Cdim)


def sim_num_tiles(B: int, N: int) -> int:
    n = _lib.lib().mst_sim_num_tiles(B, N)
    if n <= 0:
        raise ValueError("similarity loss: the number of tokens per image must be a multiple of 128")
    return n


def sim_tiles(ahat_c, inv_c, ahat_o, inv_o, B, N, Cdim, squared, partials) -> None:
    """partials[tile] = sum over the tile's strict-lower-triangle entries of |S_c - S_o| (or squares), both maps on the tensor cores."""
    _launch("mst_sim_tiles", lambda: _lib.lib().mst_sim_tiles(
        _ptr(ahat_c, torch.bfloat16, "ahat_c"), _ptr(inv_c, torch.float32, "inv_c"), _ptr(ahat_o, torch.bfloat16, "ahat_o"),
        _ptr(inv_o, torch.float32, "inv_o"), B, N, Cdim, int(squared), _ptr(partials, torch.float32, "partials"), partials.numel(), _stream()),
        flops=2.0 * 2.0 * B * N * N * Cdim)  # reference-algorithm FLOPs: two full N x N x C cosine maps


def sim_finalize(p0, count0, p1, count1, out) -> None:
    _launch("mst_sim_finalize", lambda: _lib.lib().mst_sim_finalize(_ptr(p0, torch.float32, "p0"), p0.numel(), float(count0),
                                                                   _ptr(p1, torch.float32, "p1"), 0 if p1 is None else p1.numel(),
                                                                   float(count1 or 1.0), _ptr(out, torch.float32, "out"), _stream()))


def loss_finalize(taps, lam: float, squared_style: bool, out3) -> None:
    """taps: list of dicts(partials, mean_s, var_s, mean_o, var_o, B, T, C) for relu2_1..relu5_1."""
    pack = MstLossTaps()
    pack.n_taps = len(taps)
    for i, t in enumerate(taps):
        e = pack.tap[i]
        e.partials = _ptr(t["partials"], torch.float32, "partials")
        e.mean_s, e.var_s = _ptr(t["mean_s"], torch.float32, "mean_s"), _ptr(t["var_s"], torch.float32, "var_s")
        e.mean_o, e.var_o = _ptr(t["mean_o"], torch.float32, "mean_o"), _ptr(t["var_o"], torch.float32, "var_o")
        e.n_partials, e.B, e.T, e.C = t["partials"].numel(), t["B"], t["T"], t["C"]
    _launch("mst_loss_finalize", lambda: _lib.lib().mst_loss_finalize(C.byref(pack), float(lam), int(squared_style),
                                                                       _ptr(out3, torch.float32, "out3"), _stream()))


class PackedMlp:
    """fc1/fc2 of one MLP packed as the fused kernel's weight stream, plus the fp32 biases.  With bpre set the stream also
    carries the [C, C] attention-output projection in front (MstMlp::pre)."""

    __slots__ = ("stream", "b1", "b2", "C", "bpre")

    def __init__(self, stream, b1, b2, Cdim, bpre=None):
        self.stream, self.b1, self.b2, self.C, self.bpre = stream, b1, b2, Cdim, bpre


def pack_mlp(w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
             wpre: Optional[torch.Tensor] = None, bpre: Optional[torch.Tensor] = None) -> PackedMlp:
    """nn.Linear weights fc1 [4C,C], fc2 [C,4C] (fp32) -> PackedMlp.  C must be 128 or 256.
    wpre [C,C] / bpre [C]: the attention projection fused in front of the MLP (mlp_fused(..., pre=True))."""
    w1, w2 = w1.detach().contiguous(), w2.detach().contiguous()
    hidden, Cdim = w1.shape
    if hidden != 4 * Cdim or tuple(w2.shape) != (Cdim, hidden) or Cdim not in (128, 256):
        raise ValueError("pack_mlp: expected fc1 [4C,C], fc2 [C,4C] with C in {128, 256}")
    if wpre is None:
        nbytes = _lib.lib().mst_mlp_stream_bytes(Cdim)
        stream = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w1.device)
        _launch("mst_pack_mlp_weights", lambda: _lib.lib().mst_pack_mlp_weights(_ptr(w1, torch.float32, "w1"), _ptr(w2, torch.float32, "w2"),
                                                                                stream.data_ptr(), Cdim, _stream()))
        return PackedMlp(stream, b1.detach().float().contiguous(), b2.detach().float().contiguous(), Cdim)
    wpre = wpre.detach().contiguous()
    if tuple(wpre.shape) != (Cdim, Cdim) or bpre is None:
        raise ValueError("pack_mlp: wpre must be [C,C] with a bias")
    nbytes = _lib.lib().mst_mlp_stream_bytes_pre(Cdim)
    stream = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w1.device)
    _launch("mst_pack_mlp_weights", lambda: _lib.lib().mst_pack_mlp_weights_pre(
        _ptr(wpre, torch.float32, "wpre"), _ptr(w1, torch.float32, "w1"), _ptr(w2, torch.float32, "w2"), stream.data_ptr(), Cdim, _stream()))
    return PackedMlp(stream, b1.detach().float().contiguous(), b2.detach().float().contiguous(), Cdim,
                     bpre.detach().float().contiguous())


def mlp_next_ln_supported(C_: int) -> bool:
    return C_ in (128, 256)


def mlp_fused(A, pm: PackedMlp, M: int, *, lda=None, res=None, out_f32=None, out_bf16=None,
              pre: bool = False, mul=None, ln_g=None, ln_b=None, next_ln=None) -> None:
    """out = res + fc2(gelu(fc1(A) + b1)) + b2, hidden activation kept on chip.
    pre=True (pm packed with wpre): x1 = res (*mul) + A.Wpre^T + bpre; out = x1 + mlp([LayerNorm](x1)) -- the whole
    attention-output half of a transformer block in one kernel (see include/mst_b200.h).
    next_ln = (gamma, beta[, rows]) (pre=True): out_bf16 = LayerNorm(out), the next block's norm1, on the first `rows` rows (default
    all); the other rows keep the plain bf16 copy."""
    g = MstMlp()
    g.A, g.Wstream = _ptr(A, torch.bfloat16, "A"), pm.stream.data_ptr()
    g.b1, g.b2 = _ptr(pm.b1, torch.float32, "b1"), _ptr(pm.b2, torch.float32, "b2")
    g.res, g.out_f32, g.out_bf16 = _ptr(res, torch.float32, "res"), _ptr(out_f32, torch.float32, "out_f32"), _ptr(out_bf16, torch.bfloat16, "out_bf16")
    g.M, g.C = M, pm.C
    g.lda = pm.C if lda is None else lda
    g.ld_res = g.ld_out32 = g.ld_out16 = pm.C
    flops = 16.0 * M * pm.C * pm.C
    if pre:
        if pm.bpre is None:
            raise ValueError("mlp_fused(pre=True) needs a PackedMlp built with wpre/bpre")
        g.pre = 1
        g.bpre, g.mul = _ptr(pm.bpre, torch.float32, "bpre"), _ptr(mul, torch.float32, "mul")
        g.ln_g, g.ln_b = _ptr(ln_g, torch.float32, "ln_g"), _ptr(ln_b, torch.float32, "ln_b")
        flops += 2.0 * M * pm.C * pm.C
    elif pm.bpre is not None or mul is not None or ln_g is not None:
        raise ValueError("mlp_fused: mul / ln / a pre-packed stream need pre=True")
    if next_ln is not None:
        if not (pre and mlp_next_ln_supported(pm.C)) or out_bf16 is None:
            raise ValueError("mlp_fused: next_ln needs pre=True and out_bf16")
        g.lnn_g, g.lnn_b = _ptr(next_ln[0], torch.float32, "next_ln gamma"), _ptr(next_ln[1], torch.float32, "next_ln beta")
        g.lnn_rows = int(next_ln[2]) if len(next_ln) > 2 else 0
    _launch("mst_mlp_fused", lambda: _lib.lib().mst_mlp_fused(C.byref(g), _stream()), flops=flops,
            desc=f"M={M} C={pm.C} res={res is not None} o32={out_f32 is not None} o16={out_bf16 is not None} pre={int(pre)} ln={ln_g is not None} mul={mul is not None} next_ln={next_ln is not None}")


# --------------------------------------------------------------------------------------------
# training step: backward kernels (include/mst_b200.h, "Backward kernels of the training step")
# --------------------------------------------------------------------------------------------


def wgrad(dY, X, dW, M, N, K, *, ld_dy=None, ld_x=None, conv: Optional[dict] = None, n_real: int = 0) -> None:
    """dW[n,k'] += sum_m dY[m,n] * X(m,k'); conv = dict(H, W, Cin, pad_mode, upsample) for a 3x3 Conv2d weight."""
    g = MstWgrad()
    g.dY, g.X, g.dW = _ptr(dY, torch.bfloat16, "dY"), _ptr(X, torch.bfloat16, "X"), _ptr(dW, torch.float32, "dW")
    g.M, g.N, g.K = M, N, K
    g.ld_dy = N if ld_dy is None else ld_dy
    g.n_real = n_real
    if conv is None:
        g.x_mode, g.ld_x = A_PLAIN, (K if ld_x is None else ld_x)
    else:
        g.x_mode, g.ld_x = A_CONV3X3, 0
        g.H, g.W, g.Cin = conv["H"], conv["W"], conv["Cin"]
        g.pad_mode, g.upsample = conv.get("pad_mode", PAD_ZERO), int(conv.get("upsample", False))
    _launch("mst_wgrad", lambda: _lib.lib().mst_wgrad(C.byref(g), _stream()), flops=2.0 * M * (n_real or N) * K,
            desc=f"wgrad M={M} N={N} K={K} conv={conv is not None}")


def colsum(dY, M, N, out, ld=None) -> None:
    _launch("mst_colsum", lambda: _lib.lib().mst_colsum(_ptr(dY, torch.bfloat16, "dY"), M, N, N if ld is None else ld,
                                                        _ptr(out, torch.float32, "out"), _stream()), nbytes=2.0 * M * N)


def window_attention_bwd(q, k, v, dout, dq, dk, dv, bias_table, dbias_table, B, H, W, heads, ws, shift, ldq, ldk, ldv, ldo,
                         lddq, lddk, lddv, v2=None, dout2=None, dv2=None) -> None:
    a = MstWindowAttnBwd()
    a.q, a.k, a.v, a.v2 = _ptr(q, torch.bfloat16, "q"), _ptr(k, torch.bfloat16, "k"), _ptr(v, torch.bfloat16, "v"), _ptr(v2, torch.bfloat16, "v2")
    a.dout, a.dout2 = _ptr(dout, torch.bfloat16, "dout"), _ptr(dout2, torch.bfloat16, "dout2")
    a.dq, a.dk, a.dv, a.dv2 = _ptr(dq, torch.bfloat16, "dq"), _ptr(dk, torch.bfloat16, "dk"), _ptr(dv, torch.bfloat16, "dv"), _ptr(dv2, torch.bfloat16, "dv2")
    a.bias_table, a.dbias_table = _ptr(bias_table, torch.float32, "bias_table"), _ptr(dbias_table, torch.float32, "dbias_table")
    a.B, a.H, a.W, a.heads, a.ws, a.shift = B, H, W, heads, ws, shift
    a.ldq, a.ldk, a.ldv, a.ldo, a.lddq, a.lddk, a.lddv = ldq, ldk, ldv, ldo, lddq, lddk, lddv
    n_win = B * (H // ws) * (W // ws)
    _launch("mst_window_attention_bwd", lambda: _lib.lib().mst_window_attention_bwd(C.byref(a), _stream()),
            flops=2.0 * n_win * heads * 64 * 64 * 32 * (9 if v2 is not None else 7), desc=f"B={B} H={H} dual={v2 is not None}")


def layernorm_bwd(x, gamma, dy, dx_accum, dgamma, dbeta, rows, Cdim) -> None:
    _launch("mst_layernorm_bwd", lambda: _lib.lib().mst_layernorm_bwd(
        _ptr(x, torch.float32, "x"), _ptr(gamma, torch.float32, "gamma"), _ptr(dy, torch.bfloat16, "dy"), _ptr(dx_accum, torch.float32, "dx_accum"),
        _ptr(dgamma, torch.float32, "dgamma"), _ptr(dbeta, torch.float32, "dbeta"), rows, Cdim, _stream()), nbytes=14.0 * rows * Cdim)


def instnorm_bwd(x, dy, coef, B, T, Cdim, *, twice=False, dx_accum=None, dx16=None) -> None:
    """InstanceNorm2d(affine=False) adjoint (two launches: per-(b,c) coefficients, then the elementwise apply)."""
    f32 = dy.dtype == torch.float32
    if not f32 and dy.dtype != torch.bfloat16:
        raise TypeError("instnorm_bwd: dy must be fp32 or bf16")
    _launch("mst_instnorm_bwd_stats", lambda: _lib.lib().mst_instnorm_bwd_stats(
        _ptr(x, torch.float32, "x"), _ptr(dy), int(f32), _ptr(coef, torch.float32, "coef"), B, T, Cdim, int(twice), _stream()),
        nbytes=(8.0 if f32 else 6.0) * B * T * Cdim + 4.0 * B * T * Cdim)
    _launch("mst_instnorm_bwd_apply", lambda: _lib.lib().mst_instnorm_bwd_apply(
        _ptr(x, torch.float32, "x"), _ptr(dy), int(f32), _ptr(coef, torch.float32, "coef"), _ptr(dx_accum, torch.float32, "dx_accum"),
        _ptr(dx16, torch.bfloat16, "dx16"), B, T, Cdim, _stream()), nbytes=12.0 * B * T * Cdim)


def blend_bwd(gy, sigma, query, gquery, gsigma16, gmu16) -> None:
    n = gy.numel()
    _launch("mst_blend_bwd", lambda: _lib.lib().mst_blend_bwd(
        _ptr(gy, torch.float32, "gy"), _ptr(sigma, torch.float32, "sigma"), _ptr(query, torch.float32, "query"), _ptr(gquery, torch.float32, "gquery"),
        _ptr(gsigma16, torch.bfloat16, "gsigma16"), _ptr(gmu16, torch.bfloat16, "gmu16"), n, _stream()), nbytes=20.0 * n)


def add_cast(a, b=None, out32=None, out16=None) -> None:
    n = a.numel()
    _launch("mst_add_cast", lambda: _lib.lib().mst_add_cast(_ptr(a, torch.float32, "a"), _ptr(b, torch.float32, "b"), _ptr(out32, torch.float32, "out32"),
                                                            _ptr(out16, torch.bfloat16, "out16"), n, _stream()), nbytes=8.0 * n)


def token_map_copy(src, dst, B, Hs, Ws, Hd, Wd, accumulate=False) -> None:
    """Pad (zero fill) / crop a [B,Hs,Ws,C] token map into [B,Hd,Wd,C]; accumulate=True: fp32 dst += cropped fp32 src."""
    if src.dtype != dst.dtype or src.dtype not in (torch.float32, torch.bfloat16) or (accumulate and src.dtype != torch.float32):
        raise TypeError("token_map_copy: bf16 -> bf16 or fp32 -> fp32 (accumulate: fp32 only)")
    Cdim = src.numel() // (B * Hs * Ws)
    if src.numel() != B * Hs * Ws * Cdim or dst.numel() != B * Hd * Wd * Cdim:
        raise ValueError("token_map_copy: shapes do not match the map sizes")
    tb = Cdim * src.element_size()
    _launch("mst_token_map_copy", lambda: _lib.lib().mst_token_map_copy(_ptr(src, src.dtype, "src"), _ptr(dst, dst.dtype, "dst"), B, Hs, Ws, Hd, Wd,
                                                                        tb, int(accumulate), _stream()),
            nbytes=float(tb) * B * (min(Hs, Hd) * min(Ws, Wd) * (3 if accumulate else 1) + Hd * Wd))


def reflect_fold(dxp, gate, dx, B, H, W, Cdim, upsample=False) -> None:
    _launch("mst_reflect_fold", lambda: _lib.lib().mst_reflect_fold(_ptr(dxp, torch.bfloat16, "dxp"), _ptr(gate, torch.bfloat16, "gate"),
                                                                    _ptr(dx, torch.bfloat16, "dx"), B, H, W, Cdim, int(upsample), _stream()),
            nbytes=4.0 * B * H * W * Cdim)


def maxpool2x2_bwd(x, dy, dx, B, H, W, Cdim) -> None:
    _launch("mst_maxpool2x2_bwd", lambda: _lib.lib().mst_maxpool2x2_bwd(_ptr(x, torch.bfloat16, "x"), _ptr(dy, torch.bfloat16, "dy"),
                                                                        _ptr(dx, torch.bfloat16, "dx"), B, H, W, Cdim, _stream()),
            nbytes=4.5 * B * H * W * Cdim)


def nchw3_to_nhwc8(g, out, B, H, W) -> None:
    _launch("mst_nchw3_to_nhwc8", lambda: _lib.lib().mst_nchw3_to_nhwc8(_ptr(g, torch.float32, "g"), _ptr(out, torch.bfloat16, "out"), B, H, W, _stream()),
            nbytes=28.0 * B * H * W)


def loss_bwd(fc, fo, mean_c, var_c, mean_o, var_o, mean_s, var_s, s, w, B, T, Cdim, squared_content, squared_style, dfo) -> None:
    """Gradient of content + style terms w.r.t. one tap of the stylised image (two launches); s [B,C,2] scratch is zeroed here."""
    s.zero_()
    _launch("mst_loss_bwd_stats", lambda: _lib.lib().mst_loss_bwd_stats(
        _ptr(fc, torch.bfloat16, "fc"), _ptr(fo, torch.bfloat16, "fo"), _ptr(mean_c, torch.float32), _ptr(var_c, torch.float32),
        _ptr(mean_o, torch.float32), _ptr(var_o, torch.float32), B, T, Cdim, int(squared_content), _ptr(s, torch.float32, "s"), _stream()),
        nbytes=4.0 * B * T * Cdim)
    _launch("mst_loss_bwd_apply", lambda: _lib.lib().mst_loss_bwd_apply(
        _ptr(fc, torch.bfloat16, "fc"), _ptr(fo, torch.bfloat16, "fo"), _ptr(mean_c, torch.float32), _ptr(var_c, torch.float32),
        _ptr(mean_o, torch.float32), _ptr(var_o, torch.float32), _ptr(mean_s, torch.float32), _ptr(var_s, torch.float32),
        _ptr(s, torch.float32, "s"), _ptr(w, torch.float32, "w"), B, T, Cdim, int(squared_content), int(squared_style),
        _ptr(dfo, torch.bfloat16, "dfo"), _stream()), nbytes=6.0 * B * T * Cdim)
