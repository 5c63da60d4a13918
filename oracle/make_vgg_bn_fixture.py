"""Pin the oracle's VGG-19-BN loss variant (codes/loss.py:41-63 + use_vgg19_with_batchnorm) to the REAL
reference and write tests/golden/vgg_bn_loss.json.  Build container only; TEST INFRASTRUCTURE ONLY.

    python oracle/make_vgg_bn_fixture.py

The reference never puts custom_loss in eval mode in its scripts, so "train" (batch statistics) is the mode that matters;
"eval" (running statistics, seeded here) is recorded too.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
import torch  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MST_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REPO, "tests"))

from conftest import seeded_vgg19_bn  # noqa: E402
from mastermetastyletransfer_b200 import synthetic  # noqa: E402
from oracle import master_oracle as O  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden", "vgg_bn_loss.json")


def main():
    from codes.loss import custom_loss
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "weights"))
        torch.save(seeded_vgg19_bn(), os.path.join(tmp, "weights", "vgg_bn.pt"))
        loss = custom_loss(project_absolute_path=tmp, feature_extractor_model_relative_path="weights/vgg_bn.pt",
                           use_vgg19_with_batchnorm=True)
        sd = {k: v.detach().clone() for k, v in loss.feature_extractor_model.features.state_dict().items()}
        content, style = synthetic.synthetic_images(2, 64, seed=3)
        output, _ = synthetic.synthetic_images(2, 64, seed=4)
        for mode in ("eval", "train"):  # eval first: a train-mode forward moves the running statistics
            loss.train(mode == "train")
            with torch.no_grad():
                ref = [t.item() for t in loss(content, style, output, output_content_and_style_loss=True)]
                mine = [t.item() for t in O.overall_loss(sd, content, style, output, 10.0, batchnorm=mode)]
            rel = max(abs(a - b) / abs(a) for a, b in zip(ref, mine))
            print(mode, ref, mine, f"rel {rel:.2e}")
            assert rel <= 1e-5, rel
            out[mode] = ref
        out["keys"] = len(sd)
    json.dump(out, open(GOLD, "w"), indent=0, sort_keys=True)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
