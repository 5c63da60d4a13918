"""Host-side mirror of the reference's codes/full_model.py (MasterStyleTransferModel).

Same constructor keywords, forward signature and 135-entry state_dict layout
(swin_encoder.* with torchvision names, style_transformer.*, decoder.decoder.*), and the same
attribute paths the reference scripts reach into (.swin_encoder, .style_transformer, .decoder).
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional

import torch
from torch import Tensor, nn

from . import engine
from .decoder import Decoder
from .style_transformer import StyleTransformer, packed_weights, require_cuda, require_inference, wants_grad, workspace_of


class SwinEncoderB200(nn.Sequential):
    """The pickled torchvision swin_b.features[:4] slice (codes/utils.py:59-102) with its children and
    state_dict keys unchanged; forward ([B,3,S,S] NCHW -> [B,S/8,S/8,256] BHWC) runs the sm_100a kernels."""

    def forward(self, img: Tensor) -> Tensor:  # type: ignore[override]
        require_inference(self, img)
        if img.dim() != 4 or img.shape[1] != 3 or img.shape[2] != img.shape[3] or img.shape[2] % 8:
            raise ValueError("swin encoder expects [B,3,S,S] with S a multiple of 8")
        B, _, S, _ = img.shape
        with torch.no_grad():
            w = packed_weights(self, engine.SwinEncoderWeights)
            ws = workspace_of(self, img.device)
            out = torch.empty(B, S // 8, S // 8, 256, dtype=torch.float32, device=img.device)
            engine.swin_encode(w, [img.float().contiguous()], ws, S, out, None)
        return out


class MasterStyleTransferModel(nn.Module):
    def __init__(self,
                 project_absolute_path: str = os.path.abspath(os.path.join(os.path.dirname(__file__), "..")),
                 swin_model_relative_path: str = None,
                 swin_variant: str = "swin_B",
                 style_encoder_dim: int = 256, style_decoder_dim: int = 256,
                 style_encoder_num_heads: int = 8, style_decoder_num_heads: int = 8,
                 style_encoder_window_size: List[int] = [8, 8], style_decoder_window_size: List[int] = [8, 8],
                 style_encoder_shift_size: List[int] = [4, 4], style_decoder_shift_size: List[int] = [4, 4],
                 style_encoder_mlp_ratio: float = 4.0, style_decoder_mlp_ratio: float = 4.0,
                 style_encoder_dropout: float = 0.0, style_decoder_dropout: float = 0.0,
                 style_encoder_attention_dropout: float = 0.0, style_decoder_attention_dropout: float = 0.0,
                 style_encoder_qkv_bias: bool = True, style_decoder_qkv_bias: bool = True,
                 style_encoder_proj_bias: bool = True, style_decoder_proj_bias: bool = True,
                 style_encoder_stochastic_depth_prob: float = 0.1, style_decoder_stochastic_depth_prob: float = 0.1,
                 style_encoder_norm_layer: Callable[..., nn.Module] = None,
                 style_decoder_norm_layer: Callable[..., nn.Module] = nn.LayerNorm,
                 style_encoder_MLP_activation_layer: Optional[nn.Module] = nn.GELU,
                 style_decoder_MLP_activation_layer: Optional[nn.Module] = nn.GELU,
                 style_encoder_if_use_processed_Key_in_Scale_and_Shift_calculation: bool = True,
                 style_decoder_use_instance_norm_with_affine: bool = False,
                 style_decoder_use_regular_MHA_instead_of_Swin_at_the_end: bool = False,
                 style_decoder_use_Key_instance_norm_after_linear_transformation: bool = True,
                 style_decoder_exclude_MLP_after_Fcs_self_MHA: bool = False,
                 style_transformer_load_pretrained_weights: bool = False,
                 style_transformer_pretrained_weights_path: str = None,
                 decoder_initializer: str = "kaiming_normal_",
                 direct_pretrained_style_transformer_path: str = '',
                 direct_pretrained_decoder_path: str = ''):
        super().__init__()
        kw = dict(locals())
        for name, value in kw.items():
            if name.startswith(("style_", "decoder_initializer", "direct_pretrained")):
                setattr(self, name, value)  # the reference keeps every option as an attribute (:107-137)
        if swin_variant != "swin_B":
            raise NotImplementedError("only swin_B has sm_100a kernels")

        # The reference downloads + pickles the torchvision slice on first use (utils.py:59-102) and torch.load()s
        # it (:69).  Offline, a missing file is built with random init of the same architecture instead.
        path = os.path.join(project_absolute_path, swin_model_relative_path) if swin_model_relative_path else None
        if path and os.path.exists(path):
            seq = torch.load(path, weights_only=False)
        else:
            from .synthetic import build_swin_b_first_two_stages
            seq = build_swin_b_first_two_stages()
        self.swin_encoder = SwinEncoderB200(*list(seq.children()))

        self.style_transformer = StyleTransformer(
            encoder_dim=style_encoder_dim, decoder_dim=style_decoder_dim, encoder_num_heads=style_encoder_num_heads,
            decoder_num_heads=style_decoder_num_heads, encoder_window_size=style_encoder_window_size,
            decoder_window_size=style_decoder_window_size, encoder_shift_size=style_encoder_shift_size,
            decoder_shift_size=style_decoder_shift_size, encoder_mlp_ratio=style_encoder_mlp_ratio,
            decoder_mlp_ratio=style_decoder_mlp_ratio, encoder_dropout=style_encoder_dropout,
            decoder_dropout=style_decoder_dropout, encoder_attention_dropout=style_encoder_attention_dropout,
            decoder_attention_dropout=style_decoder_attention_dropout, encoder_qkv_bias=style_encoder_qkv_bias,
            decoder_qkv_bias=style_decoder_qkv_bias, encoder_proj_bias=style_encoder_proj_bias,
            decoder_proj_bias=style_decoder_proj_bias, encoder_stochastic_depth_prob=style_encoder_stochastic_depth_prob,
            decoder_stochastic_depth_prob=style_decoder_stochastic_depth_prob, encoder_norm_layer=style_encoder_norm_layer,
            decoder_norm_layer=style_decoder_norm_layer, encoder_MLP_activation_layer=style_encoder_MLP_activation_layer,
            decoder_MLP_activation_layer=style_decoder_MLP_activation_layer,
            encoder_if_use_processed_Key_in_Scale_and_Shift_calculation=style_encoder_if_use_processed_Key_in_Scale_and_Shift_calculation,
            decoder_use_instance_norm_with_affine=style_decoder_use_instance_norm_with_affine,
            decoder_use_regular_MHA_instead_of_Swin_at_the_end=style_decoder_use_regular_MHA_instead_of_Swin_at_the_end,
            decoder_use_Key_instance_norm_after_linear_transformation=style_decoder_use_Key_instance_norm_after_linear_transformation,
            decoder_exclude_MLP_after_Fcs_self_MHA=style_decoder_exclude_MLP_after_Fcs_self_MHA)
        self.decoder = Decoder(channel_dim=style_decoder_dim, initializer=decoder_initializer)

        if style_transformer_load_pretrained_weights and not direct_pretrained_style_transformer_path:
            self.load_pretained_weights_to_style_transformer(pretrained_weights_path=style_transformer_pretrained_weights_path)
        if direct_pretrained_style_transformer_path != '':
            self.style_transformer.load_state_dict(torch.load(direct_pretrained_style_transformer_path))
        if direct_pretrained_decoder_path != '':
            self.decoder.load_state_dict(torch.load(direct_pretrained_decoder_path))

    def load_pretained_weights_to_style_transformer(self, pretrained_weights_path: str):
        """Initialise every attention / MLP of the (7x7-window) style transformer from one pretrained Swin block
        (reference full_model.py:160-212, same method name -- typo included -- and error behaviour)."""
        from .pretrained_weights import load_block_into_style_transformer
        if pretrained_weights_path is None:
            raise ValueError("Please provide the path of the pretrained weights")
        if self.style_decoder_use_regular_MHA_instead_of_Swin_at_the_end:
            raise ValueError("The pretrained weights are not compatible with the current model configuration. Please set the "
                             "style_decoder_use_regular_MHA_instead_of_Swin_at_the_end to False")
        unchanged = load_block_into_style_transformer(
            self.style_transformer, pretrained_weights_path,
            encoder_dim=self.style_encoder_dim, decoder_dim=self.style_decoder_dim,
            encoder_mlp_ratio=self.style_decoder_mlp_ratio,  # (sic: the reference passes the decoder's ratio twice, :180-181)
            decoder_mlp_ratio=self.style_decoder_mlp_ratio,
            encoder_window_size=self.style_encoder_window_size, decoder_window_size=self.style_decoder_window_size,
            encoder_qkv_bias=self.style_encoder_qkv_bias, decoder_qkv_bias=self.style_decoder_qkv_bias,
            encoder_proj_bias=self.style_encoder_proj_bias, decoder_proj_bias=self.style_decoder_proj_bias,
            encoder_norm_layer=self.style_encoder_norm_layer, decoder_norm_layer=self.style_decoder_norm_layer,
            decoder_exclude_MLP_after_Fcs_self_MHA=self.style_decoder_exclude_MLP_after_Fcs_self_MHA)
        for key in unchanged:
            print(f"PRETRAINED WEIGHTS ARE NOT LOADED CORRECTLY FOR KEY: {key}")
        return unchanged

    def forward(self, content_image: Tensor, style_image: Tensor, transformer_layer_count: int = 1) -> Tensor:
        """[B,3,S,S] x2 -> stylised [B,3,S,S] (reference :214-226), one fused pass over shared buffers."""
        require_cuda(content_image, style_image)
        if content_image.shape != style_image.shape or content_image.dim() != 4 or content_image.shape[1] != 3:
            raise ValueError("content and style must both be [B,3,S,S]")
        st_sd = self.style_transformer.training and (self.style_transformer.encoder.encoder_stochastic_depth_prob > 0
                                                     or self.style_transformer.decoder.stochastic_depth.p > 0)
        if wants_grad(self, content_image, style_image) or st_sd:
            # training: module by module as the reference composes them (full_model.py:219-226); each module records its own tape
            fc = self.swin_encoder(content_image)
            fs = self.swin_encoder(style_image)
            fcs = self.style_transformer(fc, fs, transformer_layer_count)
            return self.decoder(fcs.permute(0, 3, 1, 2))
        B, _, S, S2 = content_image.shape
        if S != S2 or S % 8:
            raise ValueError("square inputs with S a multiple of 8 only")
        return self._stylize([content_image.float().contiguous(), style_image.float().contiguous()], B, S, transformer_layer_count, None)

    def forward_u8(self, content_u8: Tensor, style_u8: Tensor, transformer_layer_count: int = 1, normalize=None,
                   out_u8: Optional[Tensor] = None) -> Tensor:
        """forward() on decoded uint8 [B,S,S,3] images (test_model.py:39-48's boundary): transforms.ToTensor() and, with
        normalize = (mean, std), transforms.Normalize are applied inside the patch-embedding kernel's image loads -- the result is
        bit-identical to forward(images_u8_to_nchw(content), images_u8_to_nchw(style)).  Inference only; S % 16 == 0.
        out_u8 (uint8 [B,S,S,3]): receives the image test_model.py:207 saves, np.clip(out * 255, 0, 255).astype(np.uint8), and
        is returned instead of the fp32 tensor (written by the last convolution's epilogue when its kernel can, else converted)."""
        require_cuda(content_u8, style_u8)
        if (content_u8.dtype != torch.uint8 or style_u8.dtype != torch.uint8 or content_u8.shape != style_u8.shape or content_u8.dim() != 4
                or content_u8.shape[3] != 3 or content_u8.shape[1] != content_u8.shape[2]):
            raise ValueError("content and style must both be uint8 [B,S,S,3]")
        if self.training:
            raise RuntimeError("forward_u8 is an inference entry point: call model.eval() first")
        B, S = int(content_u8.shape[0]), int(content_u8.shape[1])
        from . import ops
        if not ops.patch_embed_u8_supported(S):
            raise ValueError("forward_u8: S must be a multiple of 16")
        if out_u8 is not None and (out_u8.dtype != torch.uint8 or tuple(out_u8.shape) != (B, S, S, 3) or not out_u8.is_contiguous()
                                   or out_u8.device != content_u8.device):
            raise ValueError("out_u8 must be a contiguous uint8 [B,S,S,3] tensor on the inputs' device")
        return self._stylize([content_u8.contiguous(), style_u8.contiguous()], B, S, transformer_layer_count, normalize, out_u8)

    def _stylize(self, images, B: int, S: int, transformer_layer_count: int, u8_norm, out_u8: Optional[Tensor] = None) -> Tensor:
        st = self.style_transformer
        st._check_config()
        dev = images[0].device
        with torch.no_grad():
            ew = packed_weights(self.swin_encoder, engine.SwinEncoderWeights)
            sw = packed_weights(st, engine.StyleTransformerWeights)
            dw = packed_weights(self.decoder, engine.CnnDecoderWeights)
            ws = workspace_of(self, dev)
            Hf = S // 8
            feats = ws.f32("feats", 2 * B, Hf, Hf, 256)
            feats16 = ws.bf16("feats16", 2 * B, Hf, Hf, 256)
            engine.swin_encode(ew, images, ws, S, feats, feats16, u8_norm=u8_norm)
            fcs32 = ws.f32("fcs32", B, Hf, Hf, 256)
            fcs16 = ws.bf16("fcs16", B, Hf, Hf, 256)
            engine.style_transformer_forward(sw, feats[:B], feats[B:], int(transformer_layer_count), ws, B, Hf, Hf,
                                             st._cfg["window"][0], st._cfg["shift"][0], st._cfg["heads"], fcs32, fcs16,
                                             fs16_in=feats16[B:], **st.engine_flags())
            if int(transformer_layer_count) == 0:
                from . import ops
                ops.cast_bf16(fcs32.view(-1, 256), fcs16.view(-1, 256))
            if out_u8 is not None and engine.decoder_u8_supported(dw, Hf, Hf):
                engine.cnn_decoder_forward(dw, fcs16.view(B * Hf * Hf, 256), ws, B, Hf, Hf, out_u8)
                return out_u8
            out = torch.empty(B, 3, S, S, dtype=torch.float32, device=dev)
            engine.cnn_decoder_forward(dw, fcs16.view(B * Hf * Hf, 256), ws, B, Hf, Hf, out)
            if out_u8 is not None:
                from . import ops
                ops.images_nchw_to_u8(out, out_u8)
                return out_u8
        return out
