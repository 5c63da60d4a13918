"""Small driver for ncu: a handful of representative launches of each hot kernel (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
dev = "cuda"
if which in ("gemm", "all"):
    M = 32768
    A = torch.randn(M, 256, device=dev).bfloat16()
    pm = ops.pack_linear(torch.randn(256, 256, device=dev) / 16, torch.randn(256, device=dev))
    pm4 = ops.pack_linear(torch.randn(1024, 256, device=dev) / 16, torch.randn(1024, device=dev))
    o16 = torch.empty(M, 256, device=dev, dtype=torch.bfloat16)
    h16 = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(A, pm, M, out_bf16=o16)                      # plain projection
        ops.gemm(A, pm4, M, act=ops.ACT_GELU, out_bf16=h16)   # fc1 + GELU
    x = torch.randn(8, 128, 128, 64, device=dev).bfloat16()
    pc = ops.pack_conv3x3(torch.randn(64, 64, 3, 3, device=dev) / 24, torch.randn(64, device=dev))
    oc = torch.empty(8 * 128 * 128, 64, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(x, pc, 8 * 128 * 128, act=ops.ACT_RELU, out_bf16=oc, conv=dict(H=128, W=128, Cin=64, pad_mode=1, upsample=False))
if which in ("attn", "all"):
    B, H, C = 32, 32, 256
    T = B * H * H
    q, k, v = (torch.randn(T, C, device=dev).bfloat16() for _ in range(3))
    table = torch.randn(225, 8, device=dev)
    o = torch.empty(T, C, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.window_attention(q, k, v, o, table, B, H, H, 8, 8, 4, C, C, C, C)
torch.cuda.synchronize()
print("ok")
if which == "conv128":
    x = torch.randn(32, 64, 64, 128, device=dev).bfloat16()
    pc = ops.pack_conv3x3(torch.randn(128, 128, 3, 3, device=dev) / 34, torch.randn(128, device=dev))
    oc = torch.empty(32 * 64 * 64, 128, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(x, pc, 32 * 64 * 64, act=ops.ACT_RELU, out_bf16=oc, conv=dict(H=64, W=64, Cin=128, pad_mode=1, upsample=False, impl="gather"))
    torch.cuda.synchronize()
    print("ok")
