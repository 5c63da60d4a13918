"""Host-side mirror of the reference's codes/decoder.py (AdaIN-style CNN decoder).

Same constructor, forward signature (NCHW in, NCHW out) and state_dict keys
(decoder.{0,3,5,7,9,12,14,17,19}.{weight,bias}); the nine reflect-padded 3x3 convolutions run
as implicit-GEMM tcgen05 tiles with ReLU and the nearest x2 upsamples folded in (engine.py).
"""
from __future__ import annotations

import torch
from torch import nn

from . import engine
from .style_transformer import packed_weights, require_cuda, wants_grad, workspace_of

_INITS = ("default", "kaiming_normal_", "kaiming_uniform_", "xavier_normal_", "xavier_uniform_", "orthogonal_")


class Decoder(nn.Module):
    def __init__(self, channel_dim: int = 256, initializer: str = "kaiming_normal_"):
        super().__init__()
        assert initializer in _INITS, "Invalid initializer. Please choose one of the following: " + ", ".join(_INITS)
        c = channel_dim
        plan = [(c, c // 2, True), "up", (c // 2, c // 2, True), (c // 2, c // 2, True), (c // 2, c // 2, True),
                (c // 2, c // 4, True), "up", (c // 4, c // 4, True), (c // 4, c // 8, True), "up",
                (c // 8, c // 8, True), (c // 8, 3, False)]
        layers = []
        for item in plan:
            if item == "up":
                layers.append(nn.Upsample(scale_factor=2, mode="nearest"))
                continue
            cin, cout, relu = item
            layers.append(nn.Conv2d(cin, cout, (3, 3), padding=(1, 1), padding_mode="reflect"))
            if relu:
                layers.append(nn.ReLU())
        self.decoder = nn.Sequential(*layers)
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                if initializer == "kaiming_normal_":
                    nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                elif initializer == "kaiming_uniform_":
                    nn.init.kaiming_uniform_(m.weight, mode="fan_out", nonlinearity="relu")
                elif initializer == "xavier_normal_":
                    nn.init.xavier_normal_(m.weight)
                elif initializer == "xavier_uniform_":
                    nn.init.xavier_uniform_(m.weight)
                elif initializer == "orthogonal_":
                    nn.init.orthogonal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B,C,H,W] (the reference passes the BHWC transformer output permuted, full_model.py:222)."""
        require_cuda(x)
        if x.dim() != 4 or x.shape[1] != self.decoder[0].in_channels:
            raise ValueError("Decoder expects [B, channel_dim, H, W]")
        B, C, H, W = x.shape
        if wants_grad(self, x):
            from .autograd_fns import cnn_decoder_apply
            return cnn_decoder_apply(self, x)
        with torch.no_grad():
            w = packed_weights(self, engine.CnnDecoderWeights)
            ws = workspace_of(self, x.device)
            tok = x.permute(0, 2, 3, 1)  # a free view when x came from a BHWC tensor
            x16 = ws.bf16("in16", B * H * W, C)
            x16.copy_(tok.reshape(B * H * W, C))
            out = torch.empty(B, 3, 8 * H, 8 * W, dtype=torch.float32, device=x.device)
            engine.cnn_decoder_forward(w, x16, ws, B, H, W, out)
        return out
