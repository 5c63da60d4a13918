"""One conv shape through the TMA-fed implicit GEMM, checked against torch (separate process per shape: a trap is sticky)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from mastermetastyletransfer_b200 import ops
B, H, W, Cin, Cout = map(int, sys.argv[1:6]); pad = sys.argv[6]
torch.manual_seed(0)
x = torch.randn(B, H, W, Cin).bfloat16()
wt = torch.randn(Cout, Cin, 3, 3) * (9 * Cin) ** -0.5
bias = torch.randn(Cout)
pm = ops.pack_conv3x3(wt.cuda(), bias.cuda())
out = torch.empty(B * H * W, Cout, device="cuda")
ops.gemm(x.cuda(), pm, B * H * W, act=ops.ACT_RELU, out_f32=out, conv=dict(H=H, W=W, Cin=Cin, pad_mode=1 if pad == "reflect" else 0, upsample=False, impl="gather"))
torch.cuda.synchronize()
xi = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect" if pad == "reflect" else "constant")
ref = torch.relu(F.conv2d(xi, wt.bfloat16().float(), bias)).permute(0, 2, 3, 1).reshape(B * H * W, Cout)
print(sys.argv[1:], "max err", (out.cpu() - ref).abs().max().item())
