"""ctypes binding of libmst_b200.so (the C ABI declared in include/mst_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, this
raises.  The library is built in-tree by ``__graft_entry__.build()`` / ``csrc/build.py``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MST_LIB_PATH") or os.path.join(_HERE, "libmst_b200.so")  # MST_LIB_PATH: an instrumented build (tools/build_prof.sh)

c_bf16_p = C.c_void_p
c_f32_p = C.c_void_p


class MstGemm(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("Wt", C.c_void_p), ("bias", C.c_void_p), ("res", C.c_void_p), ("mul", C.c_void_p),
        ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("k_pad", C.c_int),
        ("lda", C.c_int), ("ld_res", C.c_int), ("ld_out32", C.c_int), ("ld_out16", C.c_int),
        ("a_mode", C.c_int), ("act", C.c_int),
        ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("pad_mode", C.c_int), ("upsample", C.c_int),
        ("out_nchw", C.c_int), ("n_real", C.c_int),
        ("gate", C.c_void_p), ("add16", C.c_void_p), ("out_pre16", C.c_void_p), ("row_scale", C.c_void_p),
        ("gate_mode", C.c_int), ("ld_gate", C.c_int), ("rows_per_scale", C.c_int), ("conv_full", C.c_int),
    ]


class MstWgrad(C.Structure):
    _fields_ = [
        ("dY", C.c_void_p), ("X", C.c_void_p), ("dW", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("ld_dy", C.c_int), ("ld_x", C.c_int), ("x_mode", C.c_int),
        ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("pad_mode", C.c_int), ("upsample", C.c_int),
        ("n_real", C.c_int),
    ]


class MstWindowAttnBwd(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("v2", C.c_void_p),
        ("dout", C.c_void_p), ("dout2", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p), ("dv2", C.c_void_p),
        ("bias_table", C.c_void_p), ("dbias_table", C.c_void_p),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("heads", C.c_int), ("ws", C.c_int), ("shift", C.c_int),
        ("ldq", C.c_int), ("ldk", C.c_int), ("ldv", C.c_int), ("ldo", C.c_int),
        ("lddq", C.c_int), ("lddk", C.c_int), ("lddv", C.c_int),
    ]


class MstWindowAttn(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("v2", C.c_void_p),
        ("out", C.c_void_p), ("out2", C.c_void_p), ("bias_table", C.c_void_p),
        ("pad_q", C.c_void_p), ("pad_k", C.c_void_p), ("pad_v", C.c_void_p), ("pad_v2", C.c_void_p),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("heads", C.c_int), ("ws", C.c_int), ("shift", C.c_int),
        ("ldq", C.c_int), ("ldk", C.c_int), ("ldv", C.c_int), ("ldo", C.c_int), ("pad_k_stride", C.c_int),
    ]


class MstAttnBlock(C.Structure):
    _fields_ = [("x", C.c_void_p), ("wqkv", C.c_void_p), ("bqkv", C.c_void_p), ("bias_table", C.c_void_p), ("out", C.c_void_p),
                ("dbg_qkv", C.c_void_p),
                ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("heads", C.c_int), ("ws", C.c_int), ("shift", C.c_int),
                ("ldx", C.c_int), ("ldo", C.c_int)]


class MstMlp(C.Structure):
    _fields_ = [("A", C.c_void_p), ("Wstream", C.c_void_p), ("b1", C.c_void_p), ("b2", C.c_void_p), ("res", C.c_void_p),
                ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p),
                ("M", C.c_int), ("C", C.c_int), ("lda", C.c_int), ("ld_res", C.c_int), ("ld_out32", C.c_int), ("ld_out16", C.c_int),
                ("bpre", C.c_void_p), ("mul", C.c_void_p), ("ln_g", C.c_void_p), ("ln_b", C.c_void_p), ("pre", C.c_int),
                ("lnn_g", C.c_void_p), ("lnn_b", C.c_void_p), ("lnn_rows", C.c_int)]


class MstTensorTable(C.Structure):
    _fields_ = [("chunk_start", C.c_void_p), ("numel", C.c_void_p), ("flat_offset", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p),
                ("c", C.c_void_p), ("d", C.c_void_p), ("n_tensors", C.c_int), ("n_chunks", C.c_int)]


class MstLossTap(C.Structure):
    _fields_ = [("partials", C.c_void_p), ("mean_s", C.c_void_p), ("var_s", C.c_void_p), ("mean_o", C.c_void_p),
                ("var_o", C.c_void_p), ("n_partials", C.c_int), ("B", C.c_int), ("T", C.c_int), ("C", C.c_int)]


class MstLossTaps(C.Structure):
    _fields_ = [("tap", MstLossTap * 4), ("n_taps", C.c_int)]


# every symbol include/mst_b200.h declares: name -> (restype, argtypes)
_I, _P, _Z = C.c_int, C.c_void_p, C.c_size_t
SYMBOLS = {
    "mst_version": (_I, []),
    "mst_sm_arch": (_I, []),
    "mst_error_string": (C.c_char_p, [_I]),
    "mst_pack_linear_weight": (_I, [_P, _I, _I, _P, _I, _I, _P]),
    "mst_pack_conv3x3_weight": (_I, [_P, _I, _I, _P, _I, _I, _P]),
    "mst_gemm_tile_n": (_I, [_I]),
    "mst_gemm": (_I, [C.POINTER(MstGemm), _P]),
    "mst_conv3x3_band": (_I, [C.POINTER(MstGemm), _P]),
    "mst_conv3x3_band_supported": (_I, [_I, _I, _I, _I]),
    "mst_conv3x3_rows": (_I, [C.POINTER(MstGemm), _P]),
    "mst_conv3x3_rows_supported": (_I, [_I, _I, _I, _I]),
    "mst_conv3x3_cm": (_I, [C.POINTER(MstGemm), _P]),
    "mst_conv3x3_cm_supported": (_I, [_I, _I, _I, _I]),
    "mst_mlp_stream_bytes": (_Z, [_I]),
    "mst_pack_mlp_weights": (_I, [_P, _P, _P, _I, _P]),
    "mst_mlp_stream_bytes_pre": (_Z, [_I]),
    "mst_pack_mlp_weights_pre": (_I, [_P, _P, _P, _P, _I, _P]),
    "mst_mlp_fused": (_I, [C.POINTER(MstMlp), _P]),
    "mst_window_attention": (_I, [C.POINTER(MstWindowAttn), _P]),
    "mst_window_maps": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "mst_attn_qkv_packed_bytes": (_Z, [_I, _I]),
    "mst_pack_attn_qkv": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "mst_attn_block": (_I, [C.POINTER(MstAttnBlock), _P]),
    "mst_sim_num_tiles": (_I, [_I, _I]),
    "mst_sim_prepare": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "mst_sim_tiles": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "mst_sim_finalize": (_I, [_P, _I, C.c_double, _P, _I, C.c_double, _P, _P]),
    "mst_layernorm": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "mst_patch_merge_layernorm": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "mst_instnorm_stats": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "mst_instnorm_stats_padded": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "mst_instnorm_apply": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "mst_instnorm_stats_affine": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "mst_instnorm_apply_affine": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "mst_instnorm": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "mst_instnorm_fused_supported": (_I, [_I, _I]),
    "mst_jointnorm_stats": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "mst_softmax_rows": (_I, [_P, _P, _I, _I, C.c_float, _P]),
    "mst_pack_bf16_matrix": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _P]),
    "mst_patch_embed": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "mst_patch_embed_ln": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "mst_patch_embed_ln_u8": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "mst_patch_embed_ln_u8_supported": (_I, [_I]),
    "mst_cast_bf16": (_I, [_P, _P, _Z, _P]),
    "mst_images_u8_to_nchw": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "mst_images_nchw_to_u8": (_I, [_P, _P, _I, _I, _I, _P]),
    "mst_resize_crop_normalize": (_I, [_P, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "mst_upsample2x_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "mst_opt_chunk_elems": (_I, []),
    "mst_adam_step": (_I, [C.POINTER(MstTensorTable), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _I, _P]),
    "mst_adam_step_dev": (_I, [C.POINTER(MstTensorTable), C.c_float, C.c_float, C.c_float, C.c_float, _P, _I, _P]),
    "mst_reptile_delta": (_I, [C.POINTER(MstTensorTable), _P, _P]),
    "mst_reptile_apply": (_I, [C.POINTER(MstTensorTable), _P, C.c_float, _P]),
    "mst_conv3x3_first": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "mst_maxpool2x2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "mst_bn_relu": (_I, [_P, _P, _P, _P, _P, _P, C.c_float, _Z, _I, _I, _I, _P]),
    "mst_tap_stats_scratch_floats": (_Z, [_I, _I, _I]),
    "mst_tap_stats": (_I, [_P, _P, _P, _I, _I, _I, _P, _Z, _P]),
    "mst_content_term": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "mst_loss_finalize": (_I, [C.POINTER(MstLossTaps), C.c_float, _I, _P, _P]),
    "mst_wgrad": (_I, [C.POINTER(MstWgrad), _P]),
    "mst_colsum": (_I, [_P, _I, _I, _I, _P, _P]),
    "mst_window_attention_bwd": (_I, [C.POINTER(MstWindowAttnBwd), _P]),
    "mst_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "mst_instnorm_bwd_stats": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _P]),
    "mst_instnorm_bwd_apply": (_I, [_P, _P, _I, _P, _P, _P, _I, _I, _I, _P]),
    "mst_blend_bwd": (_I, [_P, _P, _P, _P, _P, _P, _Z, _P]),
    "mst_add_cast": (_I, [_P, _P, _P, _P, _Z, _P]),
    "mst_token_map_copy": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "mst_reflect_fold": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "mst_maxpool2x2_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "mst_nchw3_to_nhwc8": (_I, [_P, _P, _I, _I, _I, _P]),
    "mst_loss_bwd_stats": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "mst_loss_bwd_apply": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the hot path).")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code == 0:
        return
    msg = lib().mst_error_string(code).decode()
    if code < 0:
        raise ValueError(f"{what}: {msg} ({code})")
    raise RuntimeError(f"{what}: CUDA error {code}: {msg}")
