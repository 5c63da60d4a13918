"""HBM throughput of the loss reductions (tap statistics, content term) at the training-step sizes (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import ops

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3

peak = 6545.9
for B in (8, 32):
    for (hw, c) in ((128, 128), (64, 256), (32, 512), (16, 512)):
        T = hw * hw
        x = torch.randn(3 * B, T, c, device="cuda").bfloat16()
        mean, var = torch.empty(3 * B, c, device="cuda"), torch.empty(3 * B, c, device="cuda")
        us = timeit(lambda: ops.tap_stats(x, mean, var, 3 * B, T, c))
        gb = x.numel() * 2 / us / 1e3
        part = torch.empty(592, device="cuda")
        xv = x.view(3 * B, T * c)
        us2 = timeit(lambda: ops.content_term(xv[:B], xv[2 * B:], mean[:B], var[:B], mean[2 * B:], var[2 * B:], B, T, c, False, part))
        gb2 = 2 * B * T * c * 2 / us2 / 1e3
        print(f"B={B:3d} tap [{3*B},{hw}x{hw},{c}]: tap_stats {us:7.1f} us {gb:7.0f} GB/s ({100*gb/peak:4.1f}% of HBM)   content_term {us2:7.1f} us {gb2:7.0f} GB/s ({100*gb2/peak:4.1f}%)")
