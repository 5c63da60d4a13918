"""Role-level cycle breakdown of conv_rows_kernel (MST_ROWS_MODE bit 3 turns on the in-kernel clock64 counters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import ops
for (B, H, Cin, Cout, up) in [(32, 256, 32, 32, True), (32, 256, 32, 16, False), (32, 128, 64, 64, False), (32, 128, 64, 32, False)]:
    hs = H // 2 if up else H
    x = torch.randn(B, hs, hs, Cin, device="cuda").bfloat16()
    pm = ops.pack_conv3x3(torch.randn(Cout, Cin, 3, 3, device="cuda") / (9 * Cin) ** 0.5, torch.randn(Cout, device="cuda"))
    M = B * H * H
    out = torch.empty(M, pm.n_pad, device="cuda", dtype=torch.bfloat16)
    print(f"--- B={B} H={H} Cin={Cin} Cout={Cout} up={up}", flush=True)
    for _ in range(2):
        ops.gemm(x, pm, M, act=1, out_bf16=out, conv=dict(H=H, W=H, Cin=Cin, pad_mode=1, upsample=up, impl="rows"))
        torch.cuda.synchronize()
