"""Mint tests/golden/path_512.npz: BASELINE configs[4] geometry (512x512 images, 64x64 feature maps, 64 8x8 windows / 100 padded
7x7 windows per image) run through the REAL reference (imported from /root/reference) on the seeded weights, with the oracle
restatement checked against it on the way.

Run in the build container only (the reference cannot travel to the GPU box):

    python oracle/make_golden_512.py

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import tempfile

os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import make_golden as MG  # noqa: E402  (puts /root/reference on sys.path)
from oracle import master_oracle as O  # noqa: E402
from mastermetastyletransfer_b200 import synthetic  # noqa: E402

SEED_512 = 2  # image seed of the 512^2 cases (tests/test_gpu_path.py uses the same)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out = {}
    with tempfile.TemporaryDirectory() as tmp, torch.no_grad():
        make_model, loss = MG.stage_reference(tmp)
        content, style = synthetic.synthetic_images(1, 512, seed=SEED_512)
        for ws in (8, 7):
            model = make_model(ws)
            sd = {k: v.clone() for k, v in model.state_dict().items()}
            for k in ((1, 2) if ws == 8 else (1,)):
                ref = model(content, style, k)
                mine = O.full_forward(sd, content, style, k, ws=ws, sh=4)
                d = (ref - mine).abs().max().item()
                print(f"full forward 512^2 ws={ws} k={k}: max|ref|={ref.abs().max().item():.4f} oracle-vs-reference max-abs {d:.3e}")
                assert d <= 5e-5 * max(1.0, ref.abs().max().item())
                name = f"img_ws{ws}_k{k}"
                out[name] = ref[:, :, ::8, ::8].contiguous().numpy()
                out[name + "_stats"] = np.array([ref.mean().item(), ref.std().item(), ref.min().item(), ref.max().item()])
                if ws == 8 and k == 1:
                    t, a, b = loss(content, style, ref, output_content_and_style_loss=True)
                    to, ao, bo = O.overall_loss({k_: v.clone() for k_, v in loss.feature_extractor_model.features.state_dict().items()},
                                                content, style, ref, lam=10.0)
                    for x, y in ((t, to), (a, ao), (b, bo)):
                        assert abs(x.item() - y.item()) <= 1e-5 * abs(x.item())
                    out["loss"] = np.array([t.item(), a.item(), b.item()], dtype=np.float64)
    path = os.path.join(MG.GOLD, "path_512.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
