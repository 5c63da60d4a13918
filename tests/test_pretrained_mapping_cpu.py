"""CPU suite: Swin-block -> style-transformer weight mapping (reference
codes/load_pretrained_weights_to_style_transformer.py:16-733, full_model.py:144-212) against the fingerprints minted from the
real reference (oracle/make_pretrained_mapping_fixture.py)."""
import json
import os

import pytest
import torch

import mastermetastyletransfer_b200 as mst
from conftest import fingerprint, seeded_swin_block
from mastermetastyletransfer_b200 import pretrained_weights as pw

KW = dict(encoder_dim=256, decoder_dim=256, encoder_num_heads=8, decoder_num_heads=8, encoder_window_size=[7, 7],
          decoder_window_size=[7, 7], encoder_shift_size=[4, 4], decoder_shift_size=[4, 4])


@pytest.fixture(scope="module")
def gold(golden_dir):
    return json.load(open(os.path.join(golden_dir, "pretrained_mapping.json")))


def test_mapping_matches_the_reference(gold):
    st = mst.StyleTransformer(**KW)
    sd = pw.load_block_into_state_dict(dict(st.state_dict()), seeded_swin_block())
    assert set(sd) == set(gold["default"]) and len(sd) == 54
    for k, v in sd.items():
        assert fingerprint(v) == pytest.approx(gold["default"][k], rel=1e-12, abs=1e-12), k
    st.load_state_dict(sd)
    # every parameter comes from the block: fused qkv split in thirds, the sigma/mu attention shares k / v / proj with it
    b = seeded_swin_block()
    assert torch.equal(st.encoder.shared_MHA_without_MLP.attn.Wk.weight, b["1.qkv.weight"][256:512])
    assert torch.equal(st.decoder.decoder_MHA_for_sigma_and_mu.Wv_shift.bias, b["1.qkv.bias"][512:])
    assert torch.equal(st.decoder.MHA_self_attn.norm2.weight, b["3.weight"])
    assert torch.equal(st.encoder.encoder_MLP_Scale[3].weight, b["4.fc2.weight"])
    assert torch.equal(st.decoder.MHA_self_attn.attn.relative_position_index, b["1.relative_position_index"].flatten())


def test_error_behaviour_matches_the_reference(gold):
    block = seeded_swin_block()
    with pytest.raises(AssertionError):  # :86-92: the block only fits 7x7 windows, dim 256, ratio 4
        pw.load_block_into_state_dict(dict(mst.StyleTransformer(**{**KW, "encoder_window_size": [8, 8], "decoder_window_size": [8, 8]}).state_dict()),
                                      block, encoder_window_size=[8, 8], decoder_window_size=[8, 8])
    bad = dict(block)
    bad["1.proj.weight"] = torch.zeros(128, 256)
    with pytest.raises(ValueError):  # shape mismatch (:433-668)
        pw.load_block_into_state_dict(dict(mst.StyleTransformer(**KW).state_dict()), bad)
    bad = dict(block)
    bad["4.fc1.bias"] = block["4.fc1.bias"].double()
    with pytest.raises(ValueError):  # dtype mismatch
        pw.load_block_into_state_dict(dict(mst.StyleTransformer(**KW).state_dict()), bad)
    # without the decoder's MLP the reference still asks for norm2 and dies with a KeyError on it (:301-304): same here
    kind, key = gold["no_self_mlp_error"]
    st = mst.StyleTransformer(**KW, decoder_exclude_MLP_after_Fcs_self_MHA=True)
    with pytest.raises(KeyError) as e:
        pw.load_block_into_state_dict(dict(st.state_dict()), block, decoder_exclude_MLP_after_Fcs_self_MHA=True)
    assert kind == "KeyError" and key in str(e.value)


def test_model_constructor_loads_the_block(tmp_path, gold):
    """MasterStyleTransferModel(style_transformer_load_pretrained_weights=True, ...) as train.py:160-196 builds it."""
    path = tmp_path / "model_basic_layer_1_module_list_shifted_window_block_state_dict.pth"
    torch.save(seeded_swin_block(), path)
    m = mst.MasterStyleTransferModel(style_encoder_window_size=[7, 7], style_decoder_window_size=[7, 7],
                                     style_transformer_load_pretrained_weights=True,
                                     style_transformer_pretrained_weights_path=str(path))
    for k, v in m.style_transformer.state_dict().items():
        assert fingerprint(v) == pytest.approx(gold["default"][k], rel=1e-12, abs=1e-12), k
    assert m.load_pretained_weights_to_style_transformer(str(path)) != []  # second load changes nothing: reported, as the reference prints
    with pytest.raises(ValueError):
        mst.MasterStyleTransferModel(style_encoder_window_size=[7, 7], style_decoder_window_size=[7, 7],
                                     style_transformer_load_pretrained_weights=True)  # no path (full_model.py:162-163)
    with pytest.raises(AssertionError):  # default 8x8 windows do not fit the block
        mst.MasterStyleTransferModel(style_transformer_load_pretrained_weights=True, style_transformer_pretrained_weights_path=str(path))
