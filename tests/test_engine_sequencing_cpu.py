"""CPU suite: host sequencing of the inference style transformer.  engine.style_transformer_forward runs with the C-ABI
wrappers replaced by torch-CPU restatements of each kernel's contract (tests/engine_ops_mock.py, bf16 buffers kept bf16) and is
compared with the oracle for the default configuration and the reference's alternate orderings, on 8x8 windows
and on 7x7 windows with the reference's in-attention zero padding.  This checks which kernel runs on which buffer in which
order -- not the kernels (those are checked on the B200, tests/test_gpu_*.py, tests/test_zz_gpu_alternates.py)."""
import pytest
import torch

from conftest import ALTERNATE_CONFIGS, alternate_inputs, alternate_style_transformer
from mastermetastyletransfer_b200 import engine
from oracle import master_oracle as O

import engine_ops_mock

FEAT_TOL = 3e-2  # of the feature map's range, as on the device (tests/test_gpu_path.py)
CONFIGS = dict(default=({}, {}), **ALTERNATE_CONFIGS)


@pytest.fixture(scope="module")
def feats():
    return alternate_inputs()


def _build(name, ws):
    if name == "default":
        from mastermetastyletransfer_b200 import StyleTransformer, synthetic
        m = StyleTransformer(encoder_dim=256, decoder_dim=256, encoder_num_heads=8, decoder_num_heads=8,
                             encoder_window_size=[ws, ws], decoder_window_size=[ws, ws], encoder_shift_size=[4, 4],
                             decoder_shift_size=[4, 4])
        return synthetic.fill_state_dict_(m, 0).eval()
    return alternate_style_transformer(name, ws)


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("ws", [8, 7])
def test_engine_sequencing_matches_oracle(monkeypatch, feats, name, ws):
    engine_ops_mock.install(monkeypatch)
    fc, fs = feats
    B, H, W, C = fc.shape
    m = _build(name, ws)
    m._check_config()
    sd = {n: v.detach().clone() for n, v in m.state_dict().items()}
    okw = CONFIGS[name][1]
    for fused in (True, False):  # the fused projection+MLP kernels and the separate-kernel sequence (MST_FUSE_PROJ_MLP=0)
        monkeypatch.setattr(engine, "FUSE_PROJ_MLP", fused)
        for k in (1, 2):
            with torch.no_grad():
                w = engine.StyleTransformerWeights(sd)
                out = torch.empty(B, H, W, C)
                engine.style_transformer_forward(w, fc, fs, k, engine.Workspace(torch.device("cpu")), B, H, W, ws, 4, 8, out,
                                                 **m.engine_flags())
                ref = O.style_transformer(sd, fc, fs, k, ws=ws, sh=4, heads=8, **okw)
                err = ((out - ref).abs().max() / (ref.max() - ref.min())).item()
                assert err <= FEAT_TOL, (name, ws, k, fused, err)
                if okw and not okw.get("exclude_mlp"):  # closer to its own configuration than to the default ordering
                    other = O.style_transformer(sd, fc, fs, k, ws=ws, sh=4, heads=8)
                    assert (out - ref).norm().item() < (out - other).norm().item(), (name, ws, k)


def test_engine_refuses_a_state_dict_that_does_not_match_the_flag(monkeypatch, feats):
    engine_ops_mock.install(monkeypatch)
    fc, fs = feats
    sd = {n: v.detach().clone() for n, v in _build("default", 8).state_dict().items()}
    w = engine.StyleTransformerWeights(sd)
    with pytest.raises(ValueError):
        engine.style_transformer_forward(w, fc, fs, 1, engine.Workspace(torch.device("cpu")), 2, 16, 16, 8, 4, 8,
                                         torch.empty(2, 16, 16, 256), exclude_mlp=True)
