"""CPU suite: the drop-in boundary (constructors, state_dict layout, module protocol, error behaviour) and
the C-ABI shared library (loads, exports every symbol include/mst_b200.h declares)."""
import copy
import ctypes
import json
import os
import re

import pytest
import torch
from torch import nn

import mastermetastyletransfer_b200 as mst
from mastermetastyletransfer_b200 import _lib, synthetic

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    return mst.MasterStyleTransferModel()


def test_state_dict_layout_matches_reference(model, golden_dir):
    layout = json.load(open(os.path.join(golden_dir, "state_dict_layout.json")))["model"]
    sd = model.state_dict()
    assert list(sd.keys()) == sorted(sd.keys(), key=list(sd.keys()).index)  # stable
    assert set(sd.keys()) == set(layout.keys())
    assert len(sd) == 135
    for k, v in sd.items():
        assert [list(v.shape), str(v.dtype)] == layout[k], k
    st = model.style_transformer.state_dict()
    assert len(st) == 54 and sum(v.dtype == torch.int64 for v in st.values()) == 3
    assert sorted(model.decoder.state_dict()) == sorted(f"decoder.{i}.{p}" for i in (0, 3, 5, 7, 9, 12, 14, 17, 19) for p in ("weight", "bias"))
    assert sum(p.numel() for p in model.parameters()) == 6419603


def test_module_protocol_used_by_reference_scripts(model):
    # train.py:306-309 deepcopy, :428-431 load_state_dict every iteration, :524-534 in-place param.data +=
    omega = copy.deepcopy(model.style_transformer)
    omega.load_state_dict(model.style_transformer.state_dict())
    with torch.no_grad():
        for (n1, p1), (n2, p2) in zip(model.style_transformer.named_parameters(), omega.named_parameters()):
            assert n1 == n2
            p1.data += 1e-4 * (p2.data - p1.data)
    # train.py:201,259-266 .apply(init) with isinstance checks
    seen = {"linear": 0, "ln": 0}

    def init(m):
        if isinstance(m, nn.Linear):
            seen["linear"] += 1
        elif isinstance(m, nn.LayerNorm):
            seen["ln"] += 1
    model.apply(init)
    assert seen["linear"] >= 30 and seen["ln"] >= 10
    # attribute paths the scripts reach into
    for path in ("swin_encoder", "style_transformer.encoder", "style_transformer.decoder", "decoder"):
        obj = model
        for part in path.split("."):
            obj = getattr(obj, part)
        assert isinstance(obj, nn.Module)
    torch.optim.Adam([{"params": model.style_transformer.parameters()}, {"params": model.decoder.parameters()}], lr=1e-4)


def test_reference_index_buffer_is_bit_exact(model, golden_dir):
    import numpy as np
    gold = np.load(os.path.join(golden_dir, "maps.npz"))
    idx = model.style_transformer.encoder.shared_MHA_without_MLP.attn.relative_position_index
    assert idx.dtype == torch.int64 and np.array_equal(idx.numpy(), gold["relidx_8"].astype(np.int64))


def test_error_behaviour(model):
    with pytest.raises(ValueError):  # codes/style_transformer.py:192-193
        mst.ShiftedWindowAttention(256, 8, [8], [4, 4])
    with pytest.raises(AssertionError):  # codes/decoder.py:21
        mst.Decoder(initializer="nope")
    content, style = synthetic.synthetic_images(1, 64)
    with pytest.raises(RuntimeError):  # no CPU fallback
        model.eval()(content, style, 1)
    with pytest.raises(RuntimeError):
        model.style_transformer.eval()(torch.zeros(1, 8, 8, 256), torch.zeros(1, 8, 8, 256))


def test_alternate_configurations_gate():
    """the three re-ordering flags run in inference AND in the training step; affine InstanceNorm and the regular-MHA
    tail are accepted for inference and refused (loudly) by the training step; dropout (active even in eval in the reference) has no
    kernels and is refused everywhere."""
    kw = dict(encoder_dim=256, decoder_dim=256, encoder_num_heads=8, decoder_num_heads=8, encoder_window_size=[8, 8],
              decoder_window_size=[8, 8], encoder_shift_size=[4, 4], decoder_shift_size=[4, 4])
    default = mst.StyleTransformer(**kw)
    default._check_config()
    default._check_config(training=True)
    assert default.engine_flags() == dict(processed_key=True, key_in_after_linear=True, exclude_mlp=False)
    for flag, key, value in (("encoder_if_use_processed_Key_in_Scale_and_Shift_calculation", "processed_key", False),
                             ("decoder_use_Key_instance_norm_after_linear_transformation", "key_in_after_linear", False),
                             ("decoder_exclude_MLP_after_Fcs_self_MHA", "exclude_mlp", True)):
        st = mst.StyleTransformer(**kw, **{flag: value if key == "exclude_mlp" else False})
        st._check_config()
        assert st.engine_flags()[key] == value
        st._check_config(training=True)
    assert len(mst.StyleTransformer(**kw, decoder_exclude_MLP_after_Fcs_self_MHA=True).state_dict()) == 48
    for flag in ("decoder_use_instance_norm_with_affine", "decoder_use_regular_MHA_instead_of_Swin_at_the_end"):
        st = mst.StyleTransformer(**kw, **{flag: True})
        st._check_config()
        with pytest.raises(NotImplementedError):
            st._check_config(training=True)
    with pytest.raises(NotImplementedError):
        mst.StyleTransformer(**kw, decoder_dropout=0.1)._check_config()


def test_seeded_fill_is_name_keyed_and_deterministic():
    a, b = mst.Decoder(), mst.Decoder()
    synthetic.fill_state_dict_(a, 3)
    synthetic.fill_state_dict_(b, 3)
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    synthetic.fill_state_dict_(b, 4)
    assert not torch.equal(a.decoder[0].weight, b.decoder[0].weight)


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "mst_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mst_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 14
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/mst_b200.h but not exported"
    assert set(declared) == set(_lib.SYMBOLS), set(declared) ^ set(_lib.SYMBOLS)
    lib = _lib.lib()
    assert lib.mst_sm_arch() == 100 and lib.mst_version() >= 100
    assert lib.mst_error_string(-1) == b"bad argument"
    assert lib.mst_gemm_tile_n(768) == 256 and lib.mst_gemm_tile_n(384) == 128 and lib.mst_gemm_tile_n(16) == 16
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.mst_gemm(None, None) == -1
    g = _lib.MstGemm()
    assert lib.mst_gemm(ctypes.byref(g), None) == -1
    assert lib.mst_layernorm(None, None, None, None, 1, 256, None) == -1
    a = _lib.MstWindowAttn()
    assert lib.mst_window_attention(ctypes.byref(a), None) == -1


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "mastermetastyletransfer_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
