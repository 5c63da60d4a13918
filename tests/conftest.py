import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(REPO, "tests", "golden")


# The reference's alternate StyleTransformer configurations that are built: name -> (constructor flags of
# codes/style_transformer.py:1185-1189, keyword arguments of the oracle).  Goldens: oracle/make_alternates_golden.py.
ALTERNATE_CONFIGS = {
    "unprocessed_key": (dict(encoder_if_use_processed_Key_in_Scale_and_Shift_calculation=False), dict(processed_key=False)),
    "no_self_mlp": (dict(decoder_exclude_MLP_after_Fcs_self_MHA=True), dict(exclude_mlp=True)),
    "key_in_before": (dict(decoder_use_Key_instance_norm_after_linear_transformation=False), dict(key_in_after_linear=False)),
}
ALTERNATE_CONFIGS["all_three"] = ({k: v for c in list(ALTERNATE_CONFIGS.values()) for k, v in c[0].items()},
                                  {k: v for c in list(ALTERNATE_CONFIGS.values()) for k, v in c[1].items()})

# Variants with kernels of their own (round 2): affine InstanceNorm (gamma folded into the statistics kernel's scale, beta added
# by the apply kernel) and the regular-MHA decoder tail (joint (C,T) normalisation + one-head global attention sequenced from
# the tensor-core GEMM with packed keys / values + a row-softmax kernel).  Inference only.  Goldens from the real reference:
# oracle/make_alternates_golden.py (which knows them under the name ORACLE_ONLY).
VARIANT_CONFIGS = ORACLE_ONLY_CONFIGS = {
    "affine_in": (dict(decoder_use_instance_norm_with_affine=True), dict(affine_in=True)),
    "affine_in_key_before": (dict(decoder_use_instance_norm_with_affine=True, decoder_use_Key_instance_norm_after_linear_transformation=False),
                             dict(affine_in=True, key_in_after_linear=False)),
    "regular_mha": (dict(decoder_use_regular_MHA_instead_of_Swin_at_the_end=True), dict(regular_mha=True)),
    "regular_mha_key_before": (dict(decoder_use_regular_MHA_instead_of_Swin_at_the_end=True,
                                    decoder_use_Key_instance_norm_after_linear_transformation=False),
                               dict(regular_mha=True, key_in_after_linear=False)),
}


def alternate_style_transformer(name: str, ws: int):
    """The drop-in StyleTransformer built with one of ALTERNATE_CONFIGS and the seeded (name-keyed) weights."""
    from mastermetastyletransfer_b200 import StyleTransformer, synthetic
    m = StyleTransformer(encoder_dim=256, decoder_dim=256, encoder_num_heads=8, decoder_num_heads=8,
                         encoder_window_size=[ws, ws], decoder_window_size=[ws, ws], encoder_shift_size=[4, 4],
                         decoder_shift_size=[4, 4], **{**ALTERNATE_CONFIGS, **ORACLE_ONLY_CONFIGS}[name][0])
    synthetic.fill_state_dict_(m, 0)
    return m.eval()


def alternate_inputs():
    """Fc, Fs [2,16,16,256]: the oracle Swin encoder's features of the seeded 128x128 synthetic images."""
    import torch
    from mastermetastyletransfer_b200 import MasterStyleTransferModel, synthetic
    from oracle import master_oracle as O
    m = MasterStyleTransferModel()
    synthetic.fill_state_dict_(m, 0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    content, style = synthetic.synthetic_images(2, 128, seed=0)
    with torch.no_grad():
        return O.swin_encoder(sd, content, "swin_encoder."), O.swin_encoder(sd, style, "swin_encoder.")


def seeded_swin_block():
    """A stand-in for the pretrained Swin block of codes/load_pretrained_weights_to_style_transformer.py:17-47: the timm keys,
    shapes and dtypes of `swin_base_patch4_window7_224` stage 2 (dim 256, 8 heads, window 7), every tensor seeded by its name."""
    import torch
    from mastermetastyletransfer_b200 import synthetic
    shapes = {"0.weight": (256,), "0.bias": (256,), "1.relative_position_bias_table": (169, 8), "1.qkv.weight": (768, 256),
              "1.qkv.bias": (768,), "1.proj.weight": (256, 256), "1.proj.bias": (256,), "3.weight": (256,), "3.bias": (256,),
              "4.fc1.weight": (1024, 256), "4.fc1.bias": (1024,), "4.fc2.weight": (256, 1024), "4.fc2.bias": (256,)}
    block = {k: torch.randn(s, generator=synthetic._gen(7, "block." + k)) for k, s in shapes.items()}
    ys, xs = torch.meshgrid(torch.arange(7), torch.arange(7), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    block["1.relative_position_index"] = (ys[:, None] - ys[None, :] + 6) * 13 + (xs[:, None] - xs[None, :] + 6)  # [49,49] int64 (timm)
    return block


def fingerprint(t):
    t = t.detach().double().reshape(-1)
    return [float(t.sum()), float(t[0]), float(t[-1]), int(t.numel())]


def seeded_vgg19_bn():
    """vgg19_bn.features[:43] (what codes/utils.py:34-36 pickles for use_vgg19_with_batchnorm), random architecture with every
    floating-point state_dict entry seeded by its name; running_var made positive."""
    import torch
    from torchvision.models import vgg19_bn
    from mastermetastyletransfer_b200 import synthetic
    seq = torch.nn.Sequential(*list(vgg19_bn(weights=None).features)[0:43])
    synthetic.fill_state_dict_(seq, 0, prefix="vggbn.")
    with torch.no_grad():
        for m in seq:
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_var.copy_(m.running_var.abs() + 0.5)
    return seq


def condition_vgg_(features, images, shift: float = 1.0):
    """Data-dependent rescaling of a seeded VGG-19 feature stack (LSUV style), in place and deterministic: conv by conv, on the
    given image batch (the test's own content / style / fp32-oracle output), scale each output channel's weights so that its
    pre-activation has unit standard deviation and set its bias so that the mean sits `shift` standard deviations above zero.
    Every tap channel then carries signal on those images: the InstanceNorm inside the content loss (eps 1e-5) no longer
    amplifies the bf16 rounding noise of nearly dead channels by 1/sqrt(1e-5) -- a property of random VGG weights, not of the
    loss -- and end-to-end gradient parity can be gated tightly.  Test infrastructure."""
    import torch
    import torch.nn.functional as F
    x = images
    with torch.no_grad():
        for m in features:
            if isinstance(m, torch.nn.Conv2d):
                y = F.conv2d(x, m.weight, None, padding=1)
                std = y.std(dim=(0, 2, 3)).clamp_min(1e-6)
                m.weight.div_(std.view(-1, 1, 1, 1))
                y = y / std.view(1, -1, 1, 1)
                m.bias.copy_(shift - y.mean(dim=(0, 2, 3)))
                x = y + m.bias.view(1, -1, 1, 1)
            else:
                x = m(x)
    return features
