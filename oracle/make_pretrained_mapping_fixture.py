"""Mint the fixture for the Swin-block -> style-transformer weight mapping from the REAL reference.

    python oracle/make_pretrained_mapping_fixture.py     # build container only; writes tests/golden/pretrained_mapping.json

TEST INFRASTRUCTURE ONLY.  A seeded stand-in for the pretrained block (same keys / shapes / dtypes as the timm
`swin_base_patch4_window7_224` stage-2 block the reference cuts out, load_pretrained_weights_to_style_transformer.py:17-47) is
pushed through the reference's own `get_pretained_weight_loaded_style_transformer_state_dict` (:689-733) for the default
configuration (and, for the error behaviour, decoder_exclude_MLP_after_Fcs_self_MHA=True); the fixture records, per destination key, a fingerprint
(sum, first, last element in float64) of what the reference put there.  tests/test_pretrained_mapping_cpu.py rebuilds the block
from the same seeds wherever the repo travels and checks the product's mapping against the fingerprints.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MST_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REPO, "tests"))

from conftest import fingerprint, seeded_swin_block  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden", "pretrained_mapping.json")


def main():
    from codes.load_pretrained_weights_to_style_transformer import get_pretained_weight_loaded_style_transformer_state_dict
    from codes.style_transformer import StyleTransformer
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "block.pth")
        torch.save(seeded_swin_block(), path)
        for name, exclude in (("default", False),):
            st = StyleTransformer(encoder_dim=256, decoder_dim=256, encoder_num_heads=8, decoder_num_heads=8,
                                  encoder_window_size=[7, 7], decoder_window_size=[7, 7], encoder_shift_size=[4, 4],
                                  decoder_shift_size=[4, 4], decoder_exclude_MLP_after_Fcs_self_MHA=exclude)
            sd = get_pretained_weight_loaded_style_transformer_state_dict(
                st.state_dict(), shifted_window_block_path=path, encoder_dim=256, decoder_dim=256, encoder_mlp_ratio=4,
                decoder_mlp_ratio=4, encoder_window_size=[7, 7], decoder_window_size=[7, 7],
                decoder_exclude_MLP_after_Fcs_self_MHA=exclude)
            st.load_state_dict(sd)  # the reference's own module accepts it
            out[name] = {k: fingerprint(v) for k, v in sd.items()}
            print(name, len(sd), "entries")
        # decoder_exclude_MLP_after_Fcs_self_MHA=True: the reference still asks for norm2 (:301-304) and dies with a KeyError
        st = StyleTransformer(encoder_dim=256, decoder_dim=256, encoder_num_heads=8, decoder_num_heads=8,
                              encoder_window_size=[7, 7], decoder_window_size=[7, 7], encoder_shift_size=[4, 4],
                              decoder_shift_size=[4, 4], decoder_exclude_MLP_after_Fcs_self_MHA=True)
        try:
            get_pretained_weight_loaded_style_transformer_state_dict(
                st.state_dict(), shifted_window_block_path=path, encoder_dim=256, decoder_dim=256, encoder_mlp_ratio=4,
                decoder_mlp_ratio=4, encoder_window_size=[7, 7], decoder_window_size=[7, 7],
                decoder_exclude_MLP_after_Fcs_self_MHA=True)
            out["no_self_mlp_error"] = None
        except Exception as e:  # noqa: BLE001
            out["no_self_mlp_error"] = [type(e).__name__, str(e).strip("'")]
        print("no_self_mlp:", out["no_self_mlp_error"])
    json.dump(out, open(GOLD, "w"), indent=0, sort_keys=True)
    print("wrote", GOLD, os.path.getsize(GOLD), "bytes")


if __name__ == "__main__":
    main()
