"""CPU suite: the oracle restatement against the golden vectors minted from the real reference
(oracle/make_golden.py).  Integer maps bit-exact, fp32 tensors to 1e-4 of their magnitude."""
import json
import os

import numpy as np
import pytest
import torch

from mastermetastyletransfer_b200 import synthetic
from oracle import master_oracle as O


@pytest.fixture(scope="module")
def model_sd():
    from mastermetastyletransfer_b200 import MasterStyleTransferModel
    m = MasterStyleTransferModel()
    synthetic.fill_state_dict_(m, 0)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


@pytest.fixture(scope="module")
def vgg_sd():
    vgg = synthetic.build_vgg19_to_relu5_1()
    synthetic.fill_state_dict_(vgg, 0, prefix="vgg.")
    return {k: v.detach().clone() for k, v in vgg.state_dict().items()}


def close(a, gold, tol=1e-4):
    gold = torch.from_numpy(np.asarray(gold)).float()
    return (a.float() - gold).abs().max().item() <= tol * max(1.0, gold.abs().max().item())


@pytest.mark.parametrize("H,ws,s", [(32, 8, 4), (64, 8, 4), (16, 8, 4), (8, 8, 4), (32, 7, 4), (64, 7, 4), (32, 7, 3), (64, 7, 3), (16, 7, 3)])
def test_integer_maps_bit_exact(golden_dir, H, ws, s):
    gold = np.load(os.path.join(golden_dir, "maps.npz"))
    Hp, Wp = O.padded_dims(H, H, ws)
    g = O.window_gather_map(H, H, ws, s)
    y, x = g // Wp, g % Wp
    mine = torch.where((y < H) & (x < H), y * H + x, torch.full_like(g, -1))
    assert np.array_equal(mine.numpy(), gold[f"gather_{H}_{ws}_{s}"])
    m = O.shift_mask(H, H, ws, s)
    key = f"mask_{H}_{ws}_{s}"
    assert (m is None) == (key not in gold)
    if m is not None:
        assert np.array_equal((m != 0).numpy().astype(np.uint8), gold[key])
        assert set(m.unique().tolist()) <= {0.0, -100.0}
    assert np.array_equal(O.relative_position_index(ws).numpy(), gold[f"relidx_{ws}"].astype(np.int64))


def test_gather_map_is_a_permutation_of_the_padded_grid():
    for H, ws, s in [(32, 8, 4), (32, 7, 3), (20, 7, 3), (24, 8, 4)]:
        Hp, Wp = O.padded_dims(H, H, ws)
        g = O.window_gather_map(H, H, ws, s).reshape(-1)
        assert sorted(g.tolist()) == list(range(Hp * Wp))


@pytest.mark.parametrize("ws,s", [(8, 4), (7, 4), (7, 3)])
def test_window_attention_component(golden_dir, ws, s):
    gold = np.load(os.path.join(golden_dir, "path_128.npz"))
    from mastermetastyletransfer_b200 import ShiftedWindowAttention
    mod = ShiftedWindowAttention(256, 8, [ws, ws], [s, s])
    synthetic.fill_state_dict_(mod, 0, prefix=f"unit{ws}.")
    g = torch.Generator().manual_seed(77 + ws + s)
    xq, xk, xv = (torch.randn(2, 16, 16, 256, generator=g) for _ in range(3))
    out = O.window_attention(xq, xk, xv, *O._attn_weights(mod.state_dict(), ""), ws, s, 8)
    assert close(out[:, ::2, ::2, ::4], gold[f"wattn_{ws}_{s}"])


def test_path_128_against_reference_goldens(golden_dir, model_sd):
    gold = np.load(os.path.join(golden_dir, "path_128.npz"))
    content, style = synthetic.synthetic_images(2, 128, seed=0)
    st = {n[len("style_transformer."):]: t for n, t in model_sd.items() if n.startswith("style_transformer.")}
    with torch.no_grad():
        fc = O.swin_encoder(model_sd, content, "swin_encoder.")
        fs = O.swin_encoder(model_sd, style, "swin_encoder.")
        assert close(fc[:, ::2, ::2, ::4], gold["fc"])
        key, scale, shift = O.style_encoder(st, fs, fs, fs, 8, 4, 8)
        assert close(key[:, ::2, ::2, ::4], gold["enc_key"]) and close(scale[:, ::2, ::2, ::4], gold["enc_scale"])
        assert close(shift[:, ::2, ::2, ::4], gold["enc_shift"])
        out1 = O.style_transformer(st, fc, fs, 1)
        assert close(out1[:, ::2, ::2, ::4], gold["st_k1"])
        assert close(O.style_transformer(st, fc, fs, 2)[:, ::2, ::2, ::4], gold["st_k2"])
        img = O.cnn_decoder(model_sd, out1.permute(0, 3, 1, 2), "decoder.decoder.")
        assert close(img, gold["img_k1"])
        assert close(O.cnn_decoder(model_sd, fc.permute(0, 3, 1, 2), "decoder.decoder.")[:, :, ::2, ::2], gold["cnn_dec"])


def test_config1_256_against_reference_goldens(golden_dir, model_sd):
    gold = np.load(os.path.join(golden_dir, "path_256.npz"))
    content, style = synthetic.synthetic_images(1, 256, seed=1)
    with torch.no_grad():
        img = O.full_forward(model_sd, content, style, 1)
    assert close(img[:, :, ::4, ::4], gold["img_k1"])
    stats = gold["img_k1_stats"]
    assert abs(img.mean().item() - stats[0]) < 1e-5 and abs(img.std().item() - stats[1]) < 1e-5


def test_loss_against_reference_goldens(golden_dir, model_sd, vgg_sd):
    gold = np.load(os.path.join(golden_dir, "path_128.npz"))
    content, style = synthetic.synthetic_images(2, 128, seed=0)
    img = torch.from_numpy(gold["img_k1"])
    with torch.no_grad():
        taps = O.vgg_taps(vgg_sd, img)
        for i, t in enumerate(taps):
            assert close(t.mean(dim=(2, 3)), gold[f"tap{i}_mean"]) and close(t.std(dim=(2, 3)), gold[f"tap{i}_std"])
        tot, lc, ls = O.overall_loss(vgg_sd, content, style, img, lam=10.0)
        np.testing.assert_allclose([tot.item(), lc.item(), ls.item()], gold["loss"], rtol=1e-5)
        tot2, lc2, ls2 = O.overall_loss(vgg_sd, content, style, img, 10.0, True, True)
        np.testing.assert_allclose([tot2.item(), lc2.item(), ls2.item()], gold["loss_squared"], rtol=1e-5)
    with pytest.raises(AssertionError):
        O.overall_loss(vgg_sd, content, style[:1], img)  # codes/loss.py:212 shape assertion


def test_instance_norm_twice_is_not_idempotent():
    """IN(IN(x)) != IN(x) at eps=1e-5 -- the reference applies both, so does the oracle."""
    x = torch.randn(1, 8, 8, 4, generator=torch.Generator().manual_seed(0)) * 0.01
    once = O.instance_norm_bhwc(x)
    assert (O.instance_norm_bhwc(once) - once).abs().max() > 1e-4


def test_u8_boundary_oracle_matches_torchvision_transforms():
    """The uint8 boundary restatement against the reference's own dependency: transforms.ToTensor() + transforms.Normalize
    (test_model.py:39-48) and the numpy clip / uint8 cast of test_model.py:207 -- bit for bit."""
    tvt = pytest.importorskip("torchvision.transforms")
    from oracle import master_oracle as O
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (2, 24, 20, 3), generator=g, dtype=torch.uint8)
    img[0, 0, :, 0] = torch.arange(20, dtype=torch.uint8) * 13  # include 0 and 247
    img[1, 1, :16, 1] = torch.arange(240, 256, dtype=torch.uint8)
    tf = tvt.Compose([tvt.ToPILImage(), tvt.ToTensor()])
    norm = tvt.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    plain = torch.stack([tf(i.numpy()) for i in img])
    assert torch.equal(O.images_u8_to_tensor(img, mean=None), plain)
    assert torch.equal(O.images_u8_to_tensor(img), torch.stack([norm(p) for p in plain]))
    x = torch.randn(2, 3, 8, 12, generator=g) * 0.7 + 0.5
    x[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 254.999 / 255, 1.0 / 255])
    want = np.stack([np.clip(xi.permute(1, 2, 0).numpy() * 255, 0, 255).astype(np.uint8) for xi in x])
    assert np.array_equal(O.tensor_to_images_u8(x).numpy(), want)


@pytest.mark.parametrize("name", ["unprocessed_key", "no_self_mlp", "key_in_before", "all_three",
                                  "affine_in", "affine_in_key_before", "regular_mha", "regular_mha_key_before"])
@pytest.mark.parametrize("ws", [8, 7])
def test_alternate_configurations_against_reference_goldens(golden_dir, name, ws):
    """the reference's alternate orderings (codes/style_transformer.py:883-909, :470-472, :389-392), oracle vs
    the outputs of the real reference built with the same flags and seeded weights (oracle/make_alternates_golden.py)."""
    from conftest import ALTERNATE_CONFIGS, ORACLE_ONLY_CONFIGS, alternate_inputs, alternate_style_transformer
    gold = np.load(os.path.join(golden_dir, "alternates.npz"))
    m = alternate_style_transformer(name, ws)
    m._check_config()  # every one of these configurations has an inference path
    if name in ORACLE_ONLY_CONFIGS:  # the training step has the default ordering and the three alternate orderings only
        with pytest.raises(NotImplementedError):
            m._check_config(training=True)
    else:
        m._check_config(training=True)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert sorted(sd.keys()) == list(gold[f"{name}_ws{ws}_keys"])  # same state_dict layout as the reference's module
    fc, fs = alternate_inputs()
    for k in (1, 2):
        with torch.no_grad():
            out = O.style_transformer(sd, fc, fs, k, ws=ws, sh=4, heads=8, **{**ALTERNATE_CONFIGS, **ORACLE_ONLY_CONFIGS}[name][1])
        assert close(out[:, ::2, ::2, ::4], gold[f"{name}_ws{ws}_k{k}"]), (name, ws, k)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_vgg_bn_loss_variant_against_reference_fixture(golden_dir, mode):
    """use_vgg19_with_batchnorm (codes/loss.py:41-63).  The reference's scripts leave the loss module in train
    mode, so BatchNorm uses the statistics of each batch it is given."""
    from conftest import seeded_vgg19_bn
    import mastermetastyletransfer_b200 as mst
    gold = json.load(open(os.path.join(golden_dir, "vgg_bn_loss.json")))
    sd = {k: v.detach().clone() for k, v in seeded_vgg19_bn().state_dict().items()}
    assert len(sd) == gold["keys"]
    content, style = synthetic.synthetic_images(2, 64, seed=3)
    output, _ = synthetic.synthetic_images(2, 64, seed=4)
    with torch.no_grad():
        got = [t.item() for t in O.overall_loss(sd, content, style, output, 10.0, batchnorm=mode)]
    assert got == pytest.approx(gold[mode], rel=1e-4)
    m = mst.custom_loss("/nonexistent", use_vgg19_with_batchnorm=True)  # builds the reference's module tree (kernels: tests/test_gpu_path.py)
    assert sorted(m.feature_extractor_model.features.state_dict().keys()) == sorted(sd.keys())


def test_pil_resize_coefficients_reproduce_pillow_bit_for_bit():
    """the fixed-point resampling coefficients the GPU transform uploads (data.pil_resize_coeffs, a restatement of
    Pillow's precompute_coeffs / normalize_coeffs_8bpc) drive a plain integer two-pass resample that equals PIL.Image.resize
    (BILINEAR) exactly -- shrinking (antialiased, many taps), enlarging and mixed cases."""
    from PIL import Image
    from mastermetastyletransfer_b200.data import PRECISION_BITS, pil_resize_coeffs
    rng = np.random.default_rng(0)
    for (H, W, oh, ow) in [(480, 640, 512, 512), (333, 500, 512, 512), (700, 1024, 512, 512), (100, 120, 512, 512), (37, 1200, 64, 48)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        xm, xc, xk = pil_resize_coeffs(W, ow)
        ym, yc, yk = pil_resize_coeffs(H, oh)
        src = img.astype(np.int64)
        tmp = np.zeros((H, ow, 3), np.int64)
        for x in range(ow):
            acc = np.full((H, 3), 1 << (PRECISION_BITS - 1), np.int64)
            for t in range(xc[x]):
                acc += src[:, xm[x] + t, :] * int(xk[x, t])
            tmp[:, x, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
        out = np.zeros((oh, ow, 3), np.uint8)
        for y in range(oh):
            acc = np.full((ow, 3), 1 << (PRECISION_BITS - 1), np.int64)
            for t in range(yc[y]):
                acc += tmp[ym[y] + t] * int(yk[y, t])
            out[y] = np.clip(acc >> PRECISION_BITS, 0, 255)
        assert np.array_equal(out, ref), (H, W, oh, ow)


def test_oracle_train_transform_is_the_reference_compose():
    """oracle.train_transform == the reference's transforms.Compose (codes/get_dataloader.py:30-36) when RandomCrop draws the same
    window: the Compose is run under a fixed seed, the drawn (top, left) recovered with RandomCrop.get_params under the same seed."""
    from torchvision import transforms
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (300, 420, 3), dtype=np.uint8)
    compose = transforms.Compose([transforms.ToPILImage(), transforms.Resize((512, 512)), transforms.RandomCrop((256, 256)),
                                  transforms.ToTensor(), transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    torch.manual_seed(123)
    ref = compose(img)
    torch.manual_seed(123)
    from mastermetastyletransfer_b200.data import GpuTrainTransform
    top, left = GpuTrainTransform("cpu").crop_params()
    assert torch.equal(O.train_transform(img, top, left), ref)
