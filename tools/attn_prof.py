"""In-kernel phase clocks of attn_fused_kernel (build with MST_NVCC_EXTRA=-DMST_AF_PROF): one launch per bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import ops
dev = "cuda"
for (B, H, C, ws, shift) in [(64, 64, 128, 7, 3), (64, 32, 256, 7, 3), (32, 32, 256, 8, 4)]:
    heads = C // 32
    T = B * H * H
    x16 = torch.randn(T, C, device=dev).bfloat16()
    w = [torch.randn(C, C, device=dev) * (0.7 / C ** 0.5) for _ in range(3)]
    b = [torch.randn(C, device=dev) * 0.2 for _ in range(3)]
    table = torch.randn((2 * ws - 1) ** 2, heads, device=dev) * 0.5
    pk = ops.pack_attn_qkv(*w, *b, heads)
    o = torch.zeros(T, C, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        ops.attn_block(x16, pk, table, o, B, H, H, ws, shift)
        torch.cuda.synchronize()
    print("----", flush=True)
