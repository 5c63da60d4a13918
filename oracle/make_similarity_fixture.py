"""Pin the `output_similarity_loss=True` return value of the REAL reference (imported from /root/reference; build container only):
get_similarity_loss (codes/loss.py:321-336) hands the CONTENT features to both arguments of every term, so the value is
identically 0 -- recorded with the other returned scalars in tests/golden/similarity_loss.json, which the drop-in's tests replay.

    python oracle/make_similarity_fixture.py

TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402  (sets sys.path for the reference and the repo)
import torch  # noqa: E402
from mastermetastyletransfer_b200 import synthetic  # noqa: E402


def main():
    with tempfile.TemporaryDirectory() as tmp, torch.no_grad():
        _, loss = G.stage_reference(tmp)
        content, style = synthetic.synthetic_images(2, 64, seed=5)
        out, _ = synthetic.synthetic_images(2, 64, seed=6)
        four = loss(content, style, out, output_content_and_style_loss=True, output_similarity_loss=True)
        two = loss(content, style, out, output_similarity_loss=True)
        assert len(four) == 4 and len(two) == 2
        assert four[3].item() == 0.0 and two[1].item() == 0.0 and four[3].dtype == torch.float32 and four[3].dim() == 0
        # The paper's form (content features vs OUTPUT features), computed with the reference's own term function
        # (loss.similarity_loss_each_term -> codes/utils.py:105-133) on the reference's own VGG taps, at 128x128 / batch 1 (the
        # reference broadcasts a [B, N, N, C] tensor: 1 GB at N = 1024) -- what the product's tensor-core kernels are checked against.
        from oracle import master_oracle as O
        c128, _ = synthetic.synthetic_images(1, 128, seed=7)
        o128, _ = synthetic.synthetic_images(1, 128, seed=8)
        o128 = (0.6 * c128 + 0.4 * o128)  # an "output" correlated with the content, as a stylised image is
        tc, to = loss.feature_extractor_model(c128), loss.feature_extractor_model(o128)
        paper = {}
        for dist in ("euclidian", "euclidian_squared"):
            from codes.loss import custom_loss as RefLoss
            term = RefLoss.__new__(RefLoss)
            torch.nn.Module.__init__(term)
            from codes.utils import get_scaled_self_cosine_distance_map_lower_triangle as gmap
            f = (lambda a, b: torch.mean(torch.square(gmap(a) - gmap(b)))) if dist == "euclidian_squared" else (lambda a, b: torch.mean(torch.abs(gmap(a) - gmap(b))))
            val = (f(tc[1], to[1]) + f(tc[2], to[2])).item()
            mine = O.similarity_loss(tc, to, dist == "euclidian_squared").item()
            assert abs(val - mine) <= 1e-5 * abs(val), (val, mine)
            paper[dist] = val
        rec = {"inputs": "synthetic_images(2, 64, seed=5) content/style, synthetic_images(2, 64, seed=6)[0] as the output image",
               "content_vs_output": {"inputs": "content = synthetic_images(1, 128, seed=7)[0], output = 0.6*content + 0.4*synthetic_images(1, 128, seed=8)[0]",
                                      **paper},
               "total_content_style_similarity": [t.item() for t in four], "total_similarity": [t.item() for t in two],
               "similarity_dtype": str(four[3].dtype), "similarity_shape": list(four[3].shape)}
    path = os.path.join(G.GOLD, "similarity_loss.json")
    with open(path, "w") as fh:
        json.dump(rec, fh, indent=1)
    print(path, rec)


if __name__ == "__main__":
    main()
