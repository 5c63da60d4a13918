"""Wait-time breakdown of the MMA warp of conv_cm_kernel (MST_CM_PROF=1 turns on the in-kernel clock64 counters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import ops
for (B, H, Cin, Cout) in [(32, 64, 128, 128), (32, 32, 256, 128), (24, 128, 128, 128), (24, 64, 256, 256), (24, 32, 512, 512)]:
    x = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
    pm = ops.pack_conv3x3(torch.randn(Cout, Cin, 3, 3, device="cuda") / (9 * Cin) ** 0.5, torch.randn(Cout, device="cuda"))
    M = B * H * H
    out = torch.empty(M, pm.n_pad, device="cuda", dtype=torch.bfloat16)
    print(f"--- B={B} H={H} Cin={Cin} Cout={Cout}", flush=True)
    for impl in ("cm", "cm", "gather"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(x, pm, M, act=1, out_bf16=out, conv=dict(H=H, W=H, Cin=Cin, pad_mode=1, impl=impl))
        e1.record()
        torch.cuda.synchronize()
        print(f"   {impl}: {e0.elapsed_time(e1) * 1e3:.1f} us  {2.0 * M * Cout * 9 * Cin / e0.elapsed_time(e1) / 1e9:.0f} TF/s", flush=True)
