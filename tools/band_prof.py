"""Phase breakdown of conv_band_kernel (build with MST_NVCC_EXTRA=-DMST_BAND_PROF)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import ops
B, H, Cin, Cout = 32, 64, 128, 128
x = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
pm = ops.pack_conv3x3(torch.randn(Cout, Cin, 3, 3, device="cuda") / (9 * Cin) ** 0.5, torch.randn(Cout, device="cuda"))
out = torch.empty(B * H * H, pm.n_pad, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm(x, pm, B * H * H, act=1, out_bf16=out, conv=dict(H=H, W=H, Cin=Cin, pad_mode=1, upsample=False, impl="band"))
    torch.cuda.synchronize()
