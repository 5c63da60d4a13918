"""Mint golden vectors by running the REAL reference (imported from /root/reference) on seeded
weights and inputs, check the oracle restatement against it, and write tests/golden/*.npz.

Run in the build container only (the reference cannot travel to the GPU box):

    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  Staging files (pickled random-init torchvision slices, SURVEY.md 8c
recipe) go to a temp dir outside the repo.
"""
from __future__ import annotations

import os
import sys
import tempfile

os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")

import numpy as np
import torch
from torch.overrides import TorchFunctionMode

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MST_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from mastermetastyletransfer_b200 import synthetic  # noqa: E402
from oracle import master_oracle as O  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
SEED = 0


def stage_reference(tmp: str):
    """SURVEY.md 8c recipe: pickle random-init slices so the reference's downloaders short-circuit."""
    os.makedirs(os.path.join(tmp, "weights"), exist_ok=True)
    torch.save(synthetic.build_swin_b_first_two_stages(), os.path.join(tmp, "weights", "swin_B_first_2_stages.pt"))
    torch.save(synthetic.build_vgg19_to_relu5_1(), os.path.join(tmp, "weights", "vgg_19_last_layer_is_relu_5_1_output.pt"))
    from codes.full_model import MasterStyleTransferModel
    from codes.loss import custom_loss

    def make_model(ws):
        m = MasterStyleTransferModel(project_absolute_path=tmp,
                                     swin_model_relative_path="weights/swin_B_first_2_stages.pt",
                                     style_encoder_window_size=[ws, ws], style_decoder_window_size=[ws, ws])
        synthetic.fill_state_dict_(m, SEED)
        return m.eval()

    loss = custom_loss(project_absolute_path=tmp,
                       feature_extractor_model_relative_path="weights/vgg_19_last_layer_is_relu_5_1_output.pt")
    synthetic.fill_state_dict_(loss.feature_extractor_model.features, SEED, prefix="vgg.")
    return make_model, loss.eval()


class Capture(TorchFunctionMode):
    """Record the windowed q-input (first F.linear input) and the {0,-100} mask the reference builds."""

    def __init__(self):
        super().__init__()
        self.linear_inputs = []
        self.masks = []

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        out = func(*args, **kwargs)
        if func is torch.nn.functional.linear:
            self.linear_inputs.append(args[0].detach().clone())
        if getattr(func, "__name__", "") == "masked_fill" and out.dim() == 3:
            self.masks.append(out.detach().clone())
        return out


def reference_maps(H, ws, s):
    """Drive codes/style_transformer.py:37-169 with a position-id tensor and capture its own maps."""
    from codes.style_transformer import shifted_window_attention

    W = H
    ids = (torch.arange(H * W, dtype=torch.float32) + 1).reshape(1, H, W, 1)
    one = torch.ones(1, 1)
    n = ws * ws
    with Capture() as cap:
        shifted_window_attention(ids, ids, ids, one, one, one, one, torch.zeros(1, 1, n, n), [ws, ws], 1, [s, s])
    gather = cap.linear_inputs[0].reshape(-1, n).round().to(torch.int64) - 1  # -1 = zero padding
    mask = cap.masks[-1] if cap.masks else None
    return gather, mask


def check_maps(out):
    from codes.style_transformer import ShiftedWindowAttention

    for ws in (7, 8):
        ref_idx = ShiftedWindowAttention(32, 1, [ws, ws], [ws // 2, ws // 2]).relative_position_index
        assert torch.equal(ref_idx, O.relative_position_index(ws)), f"rel idx ws={ws}"
        out[f"relidx_{ws}"] = ref_idx.numpy().astype(np.int16)
    for (H, ws, s) in [(32, 8, 4), (64, 8, 4), (16, 8, 4), (8, 8, 4), (32, 7, 4), (64, 7, 4), (32, 7, 3), (64, 7, 3), (16, 7, 3)]:
        g_ref, m_ref = reference_maps(H, ws, s)
        Hp, Wp = O.padded_dims(H, H, ws)
        g = O.window_gather_map(H, H, ws, s)
        y, x = g // Wp, g % Wp
        mine = torch.where((y < H) & (x < H), y * H + x, torch.full_like(g, -1))
        assert torch.equal(mine, g_ref), f"gather map {(H, ws, s)}"
        m = O.shift_mask(H, H, ws, s)
        assert (m is None) == (m_ref is None), f"mask presence {(H, ws, s)}"
        if m is not None:
            assert torch.equal(m, m_ref), f"mask {(H, ws, s)}"
            out[f"mask_{H}_{ws}_{s}"] = (m_ref != 0).numpy().astype(np.uint8)  # 1 <=> -100
        out[f"gather_{H}_{ws}_{s}"] = g_ref.numpy().astype(np.int32)
    print("integer maps: bit-exact vs reference")


def sub(t: torch.Tensor, *steps):
    idx = tuple(slice(None, None, st) for st in steps)
    return t[idx].contiguous().numpy()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out_maps, out_small, out_256 = {}, {}, {}
    with tempfile.TemporaryDirectory() as tmp, torch.no_grad():
        make_model, loss = stage_reference(tmp)
        check_maps(out_maps)

        model = make_model(8)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        import json
        layout = {k: [list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()}
        layout_vgg = {k: [list(v.shape), str(v.dtype)] for k, v in loss.feature_extractor_model.state_dict().items()}
        with open(os.path.join(GOLD, "state_dict_layout.json"), "w") as fh:
            json.dump({"model": layout, "loss.feature_extractor_model": layout_vgg}, fh, indent=0, sort_keys=True)
        vgg_sd = {k: v.clone() for k, v in loss.feature_extractor_model.features.state_dict().items()}
        st_sd = {k[len("style_transformer."):]: v for k, v in sd.items() if k.startswith("style_transformer.")}
        worst = {}

        def cmp(name, ref, mine, tol):
            d = (ref - mine).abs().max().item()
            scale = ref.abs().max().item()
            worst[name] = d
            print(f"{name:38s} max|ref|={scale:9.4f} max-abs diff={d:.3e}")
            assert d <= tol * max(1.0, scale), name

        # ---- component level: window attention with padding (7) and without (8) ----
        from codes.style_transformer import ShiftedWindowAttention
        for ws, s, H in [(8, 4, 16), (7, 4, 16), (7, 3, 16)]:
            g = torch.Generator().manual_seed(77 + ws + s)
            mod = ShiftedWindowAttention(256, 8, [ws, ws], [s, s])
            synthetic.fill_state_dict_(mod, SEED, prefix=f"unit{ws}.")
            xq, xk, xv = (torch.randn(2, H, H, 256, generator=g) for _ in range(3))
            ref = mod(xq, xk, xv)
            msd = mod.state_dict()
            mine = O.window_attention(xq, xk, xv, *O._attn_weights(msd, ""), ws, s, 8)
            cmp(f"window_attention ws={ws} s={s}", ref, mine, 2e-5)
            out_small[f"wattn_{ws}_{s}"] = sub(ref, 1, 2, 2, 4)

        # ---- full path at 128^2, B=2 ----
        content, style = synthetic.synthetic_images(2, 128, seed=SEED)
        fc_ref, fs_ref = model.swin_encoder(content), model.swin_encoder(style)
        fc = O.swin_encoder(sd, content, "swin_encoder.")
        cmp("swin_encoder 128^2", fc_ref, fc, 2e-5)
        out_small["fc"] = sub(fc_ref, 1, 2, 2, 4)
        for k in (1, 2):
            st_ref = model.style_transformer(fc_ref, fs_ref, k)
            st = O.style_transformer(st_sd, fc_ref, fs_ref, k)
            cmp(f"style_transformer k={k}", st_ref, st, 5e-5)
            out_small[f"st_k{k}"] = sub(st_ref, 1, 2, 2, 4)
            img_ref = model(content, style, k)
            img = O.full_forward(sd, content, style, k)
            cmp(f"full forward 128^2 k={k}", img_ref, img, 5e-5)
            out_small[f"img_k{k}"] = img_ref.numpy() if k == 1 else sub(img_ref, 1, 1, 2, 2)
        # encoder/decoder halves
        key_r, sc_r, sh_r = model.style_transformer.encoder(fs_ref, fs_ref, fs_ref)
        key_o, sc_o, sh_o = O.style_encoder(st_sd, fs_ref, fs_ref, fs_ref, 8, 4, 8)
        cmp("style_encoder Key", key_r, key_o, 5e-5)
        cmp("style_encoder Scale", sc_r, sc_o, 5e-5)
        cmp("style_encoder Shift", sh_r, sh_o, 5e-5)
        out_small["enc_key"], out_small["enc_scale"], out_small["enc_shift"] = (sub(t, 1, 2, 2, 4) for t in (key_r, sc_r, sh_r))
        dec_r = model.decoder(fc_ref.permute(0, 3, 1, 2))
        cmp("cnn_decoder", dec_r, O.cnn_decoder(sd, fc_ref.permute(0, 3, 1, 2), "decoder.decoder."), 5e-5)
        out_small["cnn_dec"] = sub(dec_r, 1, 1, 2, 2)

        # ---- loss at 128^2 ----
        img_ref = model(content, style, 1)
        taps_ref = loss.feature_extractor_model(img_ref)
        taps = O.vgg_taps(vgg_sd, img_ref)
        for i, (a, b) in enumerate(zip(taps_ref, taps)):
            cmp(f"vgg tap {i}", a, b, 5e-5)
            out_small[f"tap{i}_mean"] = a.mean(dim=(2, 3)).numpy()
            out_small[f"tap{i}_std"] = a.std(dim=(2, 3)).numpy()
        tot_r, lc_r, ls_r = loss(content, style, img_ref, output_content_and_style_loss=True)
        tot, lc, ls = O.overall_loss(vgg_sd, content, style, img_ref, lam=10.0)
        for n, a, b in (("total", tot_r, tot), ("content", lc_r, lc), ("style", ls_r, ls)):
            rel = abs(a.item() - b.item()) / abs(a.item())
            print(f"loss {n:8s} ref={a.item():.6f} oracle={b.item():.6f} rel={rel:.2e}")
            assert rel < 1e-5
        out_small["loss"] = np.array([tot_r.item(), lc_r.item(), ls_r.item()], dtype=np.float64)
        # squared distances variant (codes/loss.py:108-130)
        tot2, lc2, ls2 = O.overall_loss(vgg_sd, content, style, img_ref, 10.0, True, True)
        out_small["loss_squared"] = np.array([tot2.item(), lc2.item(), ls2.item()], dtype=np.float64)

        # ---- config 1 shape: B=1, 256^2, k=1 and k=3 ----
        c256, s256 = synthetic.synthetic_images(1, 256, seed=SEED + 1)
        for k in (1, 3):
            img_ref = model(c256, s256, k)
            cmp(f"full forward 256^2 k={k}", img_ref, O.full_forward(sd, c256, s256, k), 5e-5)
            out_256[f"img_k{k}"] = sub(img_ref, 1, 1, 4, 4)
            out_256[f"img_k{k}_stats"] = np.array([img_ref.mean().item(), img_ref.std().item(),
                                                   img_ref.min().item(), img_ref.max().item()])
        t, a, b = loss(c256, s256, model(c256, s256, 1), output_content_and_style_loss=True)
        out_256["loss"] = np.array([t.item(), a.item(), b.item()], dtype=np.float64)

        # ---- 7x7 style-transformer windows (CLI default train.py:703-711) ----
        model7 = make_model(7)
        sd7 = {k: v.clone() for k, v in model7.state_dict().items()}
        img7 = model7(content, style, 1)
        cmp("full forward 128^2 ws=7", img7, O.full_forward(sd7, content, style, 1, ws=7, sh=4), 5e-5)
        out_small["img_ws7"] = sub(img7, 1, 1, 2, 2)

        out_small["oracle_vs_reference_max_abs"] = np.array(sorted(worst.values())[-1])

    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "maps.npz"), **out_maps)
    np.savez_compressed(os.path.join(GOLD, "path_128.npz"), **out_small)
    np.savez_compressed(os.path.join(GOLD, "path_256.npz"), **out_256)
    for f in ("maps.npz", "path_128.npz", "path_256.npz"):
        print(f, os.path.getsize(os.path.join(GOLD, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
