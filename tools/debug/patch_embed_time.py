"""patch_embed timing: tcgen05 kernel vs the mma.sync kernel (MST_PATCH_EMBED_TC=0) on the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mastermetastyletransfer_b200 import ops
B, S = 32, 256
img = torch.randn(B, 3, S, S, device="cuda")
w, b = torch.randn(128, 3, 4, 4, device="cuda") / 7, torch.randn(128, device="cuda") * 0.05
g, be, g1, be1 = (torch.randn(128, device="cuda") * 0.1 + 1 for _ in range(4))
out = torch.empty(B, S // 4, S // 4, 128, device="cuda")
y16 = torch.empty(B, S // 4, S // 4, 128, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.patch_embed(img, w, b, g, be, out, B, S, gamma1=g1, beta1=be1, y16=y16)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    ops.patch_embed(img, w, b, g, be, out, B, S, gamma1=g1, beta1=be1, y16=y16)
e1.record()
torch.cuda.synchronize()
print("20 back-to-back launches without an L2 flush in between: %.1f us each" % (e0.elapsed_time(e1) * 1e3 / 20))
print("MST_PATCH_EMBED_TC=%s us per launch:" % os.environ.get("MST_PATCH_EMBED_TC", "1"), [round(t, 1) for t in ts], "checksum", float(out.double().sum()), float(y16.double().sum()))
