"""Initialise a 7x7-window StyleTransformer from ONE pretrained Swin block.

Host-side mirror of the reference's codes/load_pretrained_weights_to_style_transformer.py:16-733: the block is the
`swin_base_patch4_window7_224` stage-2 shifted-window block the reference cuts out of the Microsoft / timm model (:17-20),
saved as a state_dict with the keys

    0.{weight,bias}                          norm1
    1.relative_position_bias_table [169,8]   1.relative_position_index [49,49]
    1.qkv.{weight [768,256], bias [768]}     1.proj.{weight,bias}
    3.{weight,bias}                          norm2
    4.fc1.{weight,bias}   4.fc2.{weight,bias}

and its tensors are copied into EVERY attention / MLP of the style transformer: the fused qkv is split into Wq | Wk | Wv
(:56-64), the sigma/mu attention takes Wk <- k and both value projections <- v (:377-406), all five MLPs take fc1 / fc2, the
decoder's self-attention block takes norm1 / norm2.  This file is a table of (destination, source) pairs plus one loop; the
reference spells the same thing out as one 300-line function (:409-668).  Pure state_dict bookkeeping: no kernels involved.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

# source tensors of the block, by the short names used in the tables below
_BLOCK_KEYS = {
    "norm1.weight": "0.weight", "norm1.bias": "0.bias",
    "table": "1.relative_position_bias_table", "index": "1.relative_position_index",
    "proj.weight": "1.proj.weight", "proj.bias": "1.proj.bias",
    "norm2.weight": "3.weight", "norm2.bias": "3.bias",
    "fc1.weight": "4.fc1.weight", "fc1.bias": "4.fc1.bias", "fc2.weight": "4.fc2.weight", "fc2.bias": "4.fc2.bias",
}


def split_block(block: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The block's tensors on the CPU under short names, the fused qkv cut into thirds (:23-64)."""
    src = {name: block[key].cpu() for name, key in _BLOCK_KEYS.items()}
    src["index"] = src["index"].flatten()
    for part in ("weight", "bias"):
        fused = block[f"1.qkv.{part}"].cpu()
        n = fused.shape[0] // 3
        src[f"q.{part}"], src[f"k.{part}"], src[f"v.{part}"] = fused[:n], fused[n:2 * n], fused[2 * n:]
    return src


def _mlp(prefix: str) -> List[Tuple[str, str]]:
    return [(prefix + "0.weight", "fc1.weight"), (prefix + "0.bias", "fc1.bias"),
            (prefix + "3.weight", "fc2.weight"), (prefix + "3.bias", "fc2.bias")]


def _attention(prefix: str, qkv_bias: bool, proj_bias: bool) -> List[Tuple[str, str]]:
    pairs = [(prefix + "relative_position_bias_table", "table"), (prefix + "relative_position_index", "index"),
             (prefix + "Wq.weight", "q.weight"), (prefix + "Wk.weight", "k.weight"), (prefix + "Wv.weight", "v.weight"),
             (prefix + "proj.weight", "proj.weight")]
    if qkv_bias:
        pairs += [(prefix + "Wq.bias", "q.bias"), (prefix + "Wk.bias", "k.bias"), (prefix + "Wv.bias", "v.bias")]
    if proj_bias:
        pairs.append((prefix + "proj.bias", "proj.bias"))
    return pairs


def mapping(encoder_qkv_bias: bool = True, decoder_qkv_bias: bool = True, encoder_proj_bias: bool = True,
            decoder_proj_bias: bool = True, encoder_norm_layer=None, decoder_norm_layer=torch.nn.LayerNorm,
            decoder_exclude_MLP_after_Fcs_self_MHA: bool = False) -> List[Tuple[str, str]]:
    """(destination key of StyleTransformer.state_dict(), source name of split_block) in the reference's order
    (encoder :111-137, decoder :139-167)."""
    e, d, m = "encoder.shared_MHA_without_MLP.", "decoder.MHA_self_attn.", "decoder.decoder_MHA_for_sigma_and_mu."
    pairs: List[Tuple[str, str]] = []
    if encoder_norm_layer:  # (:185-194; None by default: the shared block has no norms)
        pairs += [(e + "norm1.weight", "norm1.weight"), (e + "norm1.bias", "norm1.bias")]
    pairs += _attention(e + "attn.", encoder_qkv_bias, encoder_proj_bias)
    if encoder_norm_layer:
        pairs += [(e + "norm2.weight", "norm2.weight"), (e + "norm2.bias", "norm2.bias")]
    for name in ("Key", "Scale", "Shift"):
        pairs += _mlp(f"encoder.encoder_MLP_{name}.")
    if decoder_norm_layer:
        pairs += [(d + "norm1.weight", "norm1.weight"), (d + "norm1.bias", "norm1.bias")]
    pairs += _attention(d + "attn.", decoder_qkv_bias, decoder_proj_bias)
    if decoder_norm_layer:
        pairs += [(d + "norm2.weight", "norm2.weight"), (d + "norm2.bias", "norm2.bias")]
    if not decoder_exclude_MLP_after_Fcs_self_MHA:
        pairs += _mlp(d + "mlp.")
    pairs += _mlp("decoder.last_MLP.")
    # sigma/mu attention (:358-406): no Wq (use_q_proj=False); both value projections start from the block's v
    pairs += [(m + "relative_position_bias_table", "table"), (m + "relative_position_index", "index"),
              (m + "Wk.weight", "k.weight"), (m + "Wv_scale.weight", "v.weight"), (m + "Wv_shift.weight", "v.weight"),
              (m + "proj.weight", "proj.weight")]
    if decoder_qkv_bias:
        pairs += [(m + "Wk.bias", "k.bias"), (m + "Wv_scale.bias", "v.bias"), (m + "Wv_shift.bias", "v.bias")]
    if decoder_proj_bias:
        pairs.append((m + "proj.bias", "proj.bias"))
    return pairs


def load_block_into_state_dict(state_dict: Dict[str, torch.Tensor], block: Dict[str, torch.Tensor], *, encoder_dim: int = 256,
                               decoder_dim: int = 256, encoder_mlp_ratio: float = 4, decoder_mlp_ratio: float = 4,
                               encoder_window_size=(7, 7), decoder_window_size=(7, 7), **flags) -> Dict[str, torch.Tensor]:
    """state_dict (of a StyleTransformer) with every mapped entry replaced by the block's tensor; same checks as the reference:
    AssertionError for a configuration the block does not fit (:86-92), ValueError for a shape / dtype mismatch or a missing
    destination (:433-668; a norm2 / mlp destination that the model does not have is the reference's KeyError)."""
    assert encoder_dim == 256, "encoder_dim should be 256 for pre-trained weight loading"
    assert decoder_dim == 256, "decoder_dim should be 256 for pre-trained weight loading"
    assert encoder_mlp_ratio == 4, "encoder_mlp_ratio should be 4 for pre-trained weight loading"
    assert decoder_mlp_ratio == 4, "decoder_mlp_ratio should be 4 for pre-trained weight loading"
    assert list(encoder_window_size) == [7, 7], "encoder_window_size should be [7, 7] for pre-trained weight loading"
    assert list(decoder_window_size) == [7, 7], "decoder_window_size should be [7, 7] for pre-trained weight loading"
    src = split_block(block)
    for dest, name in mapping(**flags):
        have, new = state_dict[dest], src[name]
        if have.shape != new.shape:
            raise ValueError(f"shape mismatch for {dest} (original: {tuple(have.shape)}, new: {tuple(new.shape)})")
        if have.dtype != new.dtype:
            raise ValueError(f"dtype mismatch for {dest} (original: {have.dtype}, new: {new.dtype})")
        state_dict[dest] = new
    return state_dict


def load_block_into_style_transformer(style_transformer: torch.nn.Module, block_path: Optional[str], **config) -> List[str]:
    """What MasterStyleTransferModel.load_pretained_weights_to_style_transformer does (full_model.py:160-212): read the block,
    map it, load_state_dict, and report the entries that did NOT change (the reference prints them; relative-position entries
    are exempt, :199-201).  Returns that list (empty = loaded correctly)."""
    if block_path is None:
        raise ValueError("Please provide the path of the pretrained weights")
    before = {k: v.detach().clone() for k, v in style_transformer.state_dict().items()}
    block = torch.load(block_path, map_location="cpu", weights_only=False)
    new = load_block_into_state_dict(dict(style_transformer.state_dict()), block, **config)
    style_transformer.load_state_dict(new)
    return [k for k in before if "relative_position" not in k and torch.equal(before[k], new[k].to(before[k].device))]
