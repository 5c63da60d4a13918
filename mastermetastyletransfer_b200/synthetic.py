"""Deterministic synthetic weights and inputs.

The reference's trained checkpoints and the torchvision pretrained weights are not
obtainable offline, so parity tests and benchmarks use seeded weights.
The fill is keyed by state_dict *name*, not by construction order, so applying it to
the reference's modules and to this package's modules gives bit-identical tensors
(CPU torch RNG, same torch version on both boxes).
"""
from __future__ import annotations

import math
import zlib

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
LINEAR_GAIN = 0.7  # keeps Query*sigma+mu from compounding over k layers (random init has no trained scale)


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2**31 - 1))
    return g


@torch.no_grad()
def fill_state_dict_(module: torch.nn.Module, seed: int = 0, prefix: str = "") -> torch.nn.Module:
    """Overwrite every floating-point entry of ``module.state_dict()`` in place.

    Scales keep activations O(1) through the whole path (fan-in scaled weights, LayerNorm
    gains near 1, relative-position tables with visible magnitude) so that every term of
    the arithmetic matters in the parity checks.
    """
    sd = module.state_dict()
    for key in sorted(sd.keys()):
        t = sd[key]
        if not t.is_floating_point():
            continue  # relative_position_index buffers are integer maps: left as built
        g = _gen(seed, prefix + key)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "relative_position_bias_table":
            v = torch.randn(t.shape, generator=g) * 0.5
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            gain = math.sqrt(2.0) if t.dim() == 4 else LINEAR_GAIN
            v = torch.randn(t.shape, generator=g) * (gain / math.sqrt(fan_in))
        elif leaf == "weight":  # 1-D weight: LayerNorm gain
            v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        else:  # biases
            v = 0.05 * torch.randn(t.shape, generator=g)
        t.copy_(v.to(t.dtype))
    return module


def synthetic_images(batch: int, size: int, seed: int = 0, normalize: bool = True):
    """content, style ~ U[0,1) float32 [B,3,S,S], ImageNet-normalised as test_model.py:48,111,125."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1234 + seed)
    content = torch.rand(batch, 3, size, size, generator=g)
    style = torch.rand(batch, 3, size, size, generator=g)
    if normalize:
        mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
        content = (content - mean) / std
        style = (style - mean) / std
    return content, style


def build_swin_b_first_two_stages() -> torch.nn.Sequential:
    """Random-init equivalent of what codes/utils.py:59-102 pickles (swin_b features[:4])."""
    from torchvision.models import swin_transformer

    base = swin_transformer.swin_b(weights=None)
    return torch.nn.Sequential(*list(base.features)[:4])


def build_vgg19_to_relu5_1() -> torch.nn.Sequential:
    """Random-init equivalent of what codes/utils.py:10-56 pickles (vgg19 features[:30])."""
    from torchvision.models import vgg19

    return torch.nn.Sequential(*list(vgg19(weights=None).features)[0:30])
