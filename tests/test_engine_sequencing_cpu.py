"""CPU suite: host sequencing of the inference style transformer.  engine.style_transformer_forward runs with the C-ABI
wrappers replaced by torch-CPU restatements of each kernel's contract (tests/engine_ops_mock.py, bf16 buffers kept bf16) and is
compared with the oracle for the default configuration and the reference's alternate orderings, on 8x8 windows
and on 7x7 windows with the reference's in-attention zero padding.  This checks which kernel runs on which buffer in which
order -- not the kernels (those are checked on the B200, tests/test_gpu_*.py, tests/test_zz_gpu_alternates.py)."""
import pytest
import torch

from conftest import ALTERNATE_CONFIGS, VARIANT_CONFIGS, alternate_inputs, alternate_style_transformer
from mastermetastyletransfer_b200 import engine
from oracle import master_oracle as O

import engine_ops_mock

FEAT_TOL = 3e-2  # of the feature map's range, as on the device (tests/test_gpu_path.py)
CONFIGS = dict(default=({}, {}), **ALTERNATE_CONFIGS, **VARIANT_CONFIGS)


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
                if okw and not okw.get("exclude_mlp") and name not in VARIANT_CONFIGS:  # closer to its own configuration than to the default ordering
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


@pytest.mark.parametrize("name", ["default", "all_three"])
def test_full_model_forward_passes_the_configuration_to_the_engine(monkeypatch, name):
    """MasterStyleTransferModel.forward (inference branch) with the Swin encoder and the CNN decoder replaced by the oracle and
    the style transformer's kernels by their stand-ins: the module -> engine hand-over (buffers, layer count, window / shift /
    heads, the alternate-ordering flags) against O.full_forward."""
    import mastermetastyletransfer_b200 as mst
    from mastermetastyletransfer_b200 import full_model, synthetic
    engine_ops_mock.install(monkeypatch)
    flags = {"style_" + k: v for k, v in CONFIGS[name][0].items()}
    m = synthetic.fill_state_dict_(mst.MasterStyleTransferModel(**flags), 0).eval()
    sd = {n: v.detach().clone() for n, v in m.state_dict().items()}

    class Raw:  # weight holders of the two stubbed stages: just the state_dict
        def __init__(self, sd_):
            self.sd = sd_

    def swin_encode(w, imgs, ws_, S, out32, out16, u8_norm=None):
        assert u8_norm is None  # fp32 tensors in: no uint8 conversion asked of the encoder
        out32.copy_(torch.cat([O.swin_encoder(w.sd, i, "") for i in imgs], 0))
        if out16 is not None:  # the encoder also hands the style transformer the bf16 copy of the features
            out16.copy_(out32)

    def cnn_decoder_forward(w, x16, ws_, B, H, W, out):
        out.copy_(O.cnn_decoder(w.sd, x16.float().view(B, H, W, 256).permute(0, 3, 1, 2), "decoder."))

    monkeypatch.setattr(engine, "SwinEncoderWeights", Raw)
    monkeypatch.setattr(engine, "CnnDecoderWeights", Raw)
    monkeypatch.setattr(engine, "swin_encode", swin_encode)
    monkeypatch.setattr(engine, "cnn_decoder_forward", cnn_decoder_forward)
    monkeypatch.setattr(full_model, "require_cuda", lambda *t: None)
    content, style = synthetic.synthetic_images(2, 128, seed=0)
    for k in (1, 2):
        with torch.no_grad():
            out = m(content, style, k)
            fc, fs = O.swin_encoder(sd, content, "swin_encoder."), O.swin_encoder(sd, style, "swin_encoder.")
            st = {n[len("style_transformer."):]: t for n, t in sd.items() if n.startswith("style_transformer.")}
            ref = O.cnn_decoder(sd, O.style_transformer(st, fc, fs, k, **CONFIGS[name][1]).permute(0, 3, 1, 2), "decoder.decoder.")
        assert out.shape == ref.shape == (2, 3, 128, 128)
        assert ((out - ref).abs().max() / (ref.max() - ref.min())).item() <= 2e-2


def test_style_transformer_module_forward_passes_the_configuration_to_the_engine(monkeypatch, feats):
    from mastermetastyletransfer_b200 import style_transformer as st_mod
    engine_ops_mock.install(monkeypatch)
    monkeypatch.setattr(st_mod, "require_cuda", lambda *t: None)
    fc, fs = feats
    for name in ("default", "unprocessed_key", "all_three"):
        m = _build(name, 7)
        sd = {n: v.detach().clone() for n, v in m.state_dict().items()}
        with torch.no_grad():
            out = m(fc, fs, 2)
            ref = O.style_transformer(sd, fc, fs, 2, ws=7, sh=4, heads=8, **CONFIGS[name][1])
        assert ((out - ref).abs().max() / (ref.max() - ref.min())).item() <= FEAT_TOL, name


@pytest.mark.parametrize("ws,size", [(8, 128), (7, 128), (8, 64)])
def test_whole_inference_forward_sequencing(monkeypatch, ws, size):
    """The complete inference path of MasterStyleTransferModel.forward -- Swin encoder (7x7 windows on zero-padded maps, patch
    merging), style transformer, CNN decoder (reflect padding, folded and materialised upsampling, NCHW output) -- as the engine
    sequences it, with every kernel replaced by its stand-in, against O.full_forward."""
    import mastermetastyletransfer_b200 as mst
    from mastermetastyletransfer_b200 import full_model, synthetic
    engine_ops_mock.install(monkeypatch)
    monkeypatch.setattr(full_model, "require_cuda", lambda *t: None)
    m = mst.MasterStyleTransferModel(style_encoder_window_size=[ws, ws], style_decoder_window_size=[ws, ws])
    synthetic.fill_state_dict_(m, 0).eval()
    sd = {n: v.detach().clone() for n, v in m.state_dict().items()}
    content, style = synthetic.synthetic_images(2, size, seed=0)
    with torch.no_grad():
        out = m(content, style, 1)
        ref = O.full_forward(sd, content, style, 1, ws=ws, sh=4)
    assert out.shape == ref.shape == (2, 3, size, size)
    err = ((out - ref).abs().max() / (ref.max() - ref.min())).item()
    assert err <= 2e-2, err


@pytest.mark.parametrize("squared", [False, True])
def test_perceptual_loss_sequencing(monkeypatch, squared):
    """engine.perceptual_loss_forward (VGG-19 taps of content | style | output as one batch, tap statistics, content term,
    finalize) with the kernels replaced by their stand-ins, against the oracle's get_overall_loss (loss.py:201-262).
    Tolerance: 1e-2 relative (bf16 taps; BASELINE's 1e-3 is for the device kernels' own accumulation order, checked on the B200)."""
    from mastermetastyletransfer_b200 import synthetic
    engine_ops_mock.install(monkeypatch)
    vgg = synthetic.build_vgg19_to_relu5_1()
    synthetic.fill_state_dict_(vgg, 0, prefix="vgg.")
    sd = {k: v.detach().clone() for k, v in vgg.state_dict().items()}
    content, style = synthetic.synthetic_images(2, 64, seed=3)
    output, _ = synthetic.synthetic_images(2, 64, seed=4)
    with torch.no_grad():
        w = engine.VggWeights(sd, prefix="")
        out3 = engine.perceptual_loss_forward(w, content, style, output, 10.0, squared, squared, engine.Workspace(torch.device("cpu")))
        ref = torch.stack(O.overall_loss(sd, content, style, output, 10.0, squared, squared))
    assert torch.allclose(out3, ref, rtol=1e-2), (out3, ref)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_perceptual_loss_vgg_bn_sequencing(monkeypatch, mode):
    """engine.perceptual_loss_forward_bn (VGG-19-BN extractor, three separate passes, batch statistics in train mode /
    running statistics in eval mode, BatchNorm + ReLU in place) with the kernels replaced by their stand-ins, against the oracle."""
    from conftest import seeded_vgg19_bn
    from mastermetastyletransfer_b200 import synthetic
    engine_ops_mock.install(monkeypatch)
    sd = {k: v.detach().clone() for k, v in seeded_vgg19_bn().state_dict().items()}
    content, style = synthetic.synthetic_images(2, 64, seed=3)
    output, _ = synthetic.synthetic_images(2, 64, seed=4)
    with torch.no_grad():
        w = engine.VggBnWeights(sd, prefix="")
        out3 = engine.perceptual_loss_forward_bn(w, content, style, output, 10.0, False, False, engine.Workspace(torch.device("cpu")),
                                                 training=mode == "train")
        ref = torch.stack(O.overall_loss(sd, content, style, output, 10.0, batchnorm=mode))
    assert torch.allclose(out3, ref, rtol=2e-2), (out3, ref)
