"""Path-level parity on the B200: the drop-in modules (CUDA kernels through the C ABI) against the CPU
oracle on the same seeded weights and inputs, and against the golden vectors minted from the real
reference (oracle/make_golden.py).

Tolerances (BASELINE.json north_star): stylised images max-abs <= 2e-2 on [0,1] images for bf16.
The seeded random-init model produces images with a dynamic range of about 1.3-1.6, so the bound is
applied to the error divided by the reference image's range (max - min).  Intermediate feature maps are
checked at 3e-2 of their range (they feed further bf16 GEMMs).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

IMG_TOL = 2e-2
FEAT_TOL = 3e-2


@pytest.fixture(scope="module")
def model():
    from mastermetastyletransfer_b200 import MasterStyleTransferModel, synthetic
    m = MasterStyleTransferModel()
    synthetic.fill_state_dict_(m, 0)
    return m.eval().cuda()


@pytest.fixture(scope="module")
def sd(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def rel_err(out, ref):
    ref = ref.float()
    return ((out.float().cpu() - ref).abs().max() / (ref.max() - ref.min())).item()


def test_swin_encoder_vs_oracle(model, sd):
    from mastermetastyletransfer_b200 import synthetic
    from oracle import master_oracle as O
    content, _ = synthetic.synthetic_images(2, 128, seed=0)
    with torch.no_grad():
        out = model.swin_encoder(content.cuda())
        ref = O.swin_encoder(sd, content, "swin_encoder.")
    assert out.shape == ref.shape == (2, 16, 16, 256)
    assert rel_err(out, ref) <= FEAT_TOL, rel_err(out, ref)


@pytest.mark.parametrize("k", [1, 2])
def test_style_transformer_vs_oracle(model, sd, k):
    from mastermetastyletransfer_b200 import synthetic
    from oracle import master_oracle as O
    content, style = synthetic.synthetic_images(2, 128, seed=0)
    st = {n[len("style_transformer."):]: t for n, t in sd.items() if n.startswith("style_transformer.")}
    with torch.no_grad():
        fc, fs = O.swin_encoder(sd, content, "swin_encoder."), O.swin_encoder(sd, style, "swin_encoder.")
        ref = O.style_transformer(st, fc, fs, k)
        out = model.style_transformer(fc.cuda(), fs.cuda(), k)
    assert rel_err(out, ref) <= FEAT_TOL, rel_err(out, ref)


def test_cnn_decoder_vs_oracle(model, sd):
    from oracle import master_oracle as O
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 16, 16, 256, generator=g)
    with torch.no_grad():
        ref = O.cnn_decoder(sd, x.permute(0, 3, 1, 2), "decoder.decoder.")
        out = model.decoder(x.cuda().permute(0, 3, 1, 2))
    assert out.shape == ref.shape == (2, 3, 128, 128)
    assert rel_err(out, ref) <= IMG_TOL, rel_err(out, ref)


@pytest.mark.parametrize("k", [1, 2])
def test_full_forward_vs_oracle_and_golden(model, sd, k, golden_dir):
    from mastermetastyletransfer_b200 import synthetic
    from oracle import master_oracle as O
    content, style = synthetic.synthetic_images(2, 128, seed=0)
    with torch.no_grad():
        out = model(content.cuda(), style.cuda(), k)
        ref = O.full_forward(sd, content, style, k)
    assert out.shape == (2, 3, 128, 128) and out.dtype == torch.float32
    e = rel_err(out, ref)
    assert e <= IMG_TOL, e
    gold = np.load(os.path.join(golden_dir, "path_128.npz"))
    g = torch.from_numpy(gold[f"img_k{k}"])
    mine = out.cpu() if k == 1 else out.cpu()[:, :, ::2, ::2]
    assert ((mine - g).abs().max() / (g.max() - g.min())).item() <= IMG_TOL


@pytest.mark.parametrize("k", [1, 3])
def test_config1_256_vs_golden(model, k, golden_dir):
    """BASELINE configs[0] shape: one 256x256 content + style pair, batch 1 (seeded weights: SURVEY 8c)."""
    from mastermetastyletransfer_b200 import synthetic
    content, style = synthetic.synthetic_images(1, 256, seed=1)
    with torch.no_grad():
        out = model(content.cuda(), style.cuda(), k).cpu()
    gold = np.load(os.path.join(golden_dir, "path_256.npz"))
    g = torch.from_numpy(gold[f"img_k{k}"])
    rng = float(gold[f"img_k{k}_stats"][3] - gold[f"img_k{k}_stats"][2])
    e = ((out[:, :, ::4, ::4] - g).abs().max() / rng).item()
    assert e <= (IMG_TOL if k == 1 else 2 * IMG_TOL), e  # three shared-weight layers compound the bf16 rounding


def test_batch_independence_and_determinism(model):
    """Size-independent properties at the bench shape: images are independent (batch split == whole batch)
    and reruns are bit-identical."""
    from mastermetastyletransfer_b200 import synthetic
    content, style = synthetic.synthetic_images(8, 256, seed=3)
    c, s = content.cuda(), style.cuda()
    with torch.no_grad():
        whole = model(c, s, 1).clone()
        again = model(c, s, 1).clone()
        halves = torch.cat([model(c[:4], s[:4], 1).clone(), model(c[4:], s[4:], 1).clone()])
    assert torch.equal(whole, again)
    assert torch.equal(whole, halves)


def test_no_cpu_fallback(model):
    from mastermetastyletransfer_b200 import synthetic
    content, style = synthetic.synthetic_images(1, 64, seed=0)
    with pytest.raises(RuntimeError):
        model(content, style, 1)  # CPU tensors: the product path refuses instead of falling back
