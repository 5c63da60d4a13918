"""Time individual GEMM / conv shapes with CUDA events (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import ops

dev = "cuda"
torch.manual_seed(0)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us

def plain(M, N, K, o32=False, o16=True, res=False, act=0, ext=False):
    A = torch.randn(M, K, device=dev).bfloat16()
    pm = ops.pack_linear(torch.randn(N, K, device=dev) / K ** 0.5, torch.randn(N, device=dev))
    out16 = torch.empty(M, pm.n_pad, device=dev, dtype=torch.bfloat16) if o16 else None
    out32 = torch.empty(M, pm.n_pad, device=dev) if o32 else None
    r = torch.randn(M, pm.n_pad, device=dev) if res else None
    rs = torch.ones(M // 128 + 1, device=dev) if ext else None
    us = timeit(lambda: ops.gemm(A, pm, M, act=act, res=r, out_f32=out32, out_bf16=out16, row_scale=rs, rows_per_scale=128))
    by = M * K * 2 + (M * N * 2 if o16 else 0) + (M * N * 4 if o32 else 0) + (M * N * 4 if res else 0)
    print(f"plain ext={int(ext)} M={M:8d} N={N:4d} K={K:4d} o32={int(o32)} o16={int(o16)} res={int(res)} act={act}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TF/s  {by/us/1e3:7.1f} GB/s(alg)")

def conv(B, H, Cin, Cout, up=False, reflect=True, impl='auto'):
    hs = H // 2 if up else H
    x = torch.randn(B, hs, hs, Cin, device=dev).bfloat16()
    pm = ops.pack_conv3x3(torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5, torch.randn(Cout, device=dev))
    M = B * H * H
    out = torch.empty(M, pm.n_pad, device=dev, dtype=torch.bfloat16)
    try:
        us = timeit(lambda: ops.gemm(x, pm, M, act=1, out_bf16=out, conv=dict(H=H, W=H, Cin=Cin, pad_mode=1 if reflect else 0, upsample=up, impl=impl)))
    except ValueError as e:
        print(f'conv[{impl}] B={B} H={H} Cin={Cin} Cout={Cout}: unsupported'); return
    by = x.numel() * 2 + M * Cout * 2
    print(f"conv[{impl}] B={B} H={H} Cin={Cin} Cout={Cout} up={int(up)}: {us:8.1f} us  {2*M*Cout*9*Cin/us/1e6:7.1f} TF/s  {by/us/1e3:7.1f} GB/s(alg)")

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("plain", "all"):
    plain(32768, 256, 256)
    plain(32768, 256, 256, ext=True)
    plain(262144, 384, 128, ext=True)
    plain(32768, 1024, 256, act=2)
    plain(32768, 256, 1024, o32=True, res=True)
    plain(262144, 384, 128)
    plain(262144, 128, 128, o32=True, o16=False, res=True)
    plain(2097152, 32, 320)
    plain(524288, 64, 576)
    plain(131072, 128, 1152)
    plain(8192, 256, 8192)  # compute-heavy sanity: tensor-pipe ceiling of this kernel
if which in ("conv", "all"):
    conv(32, 256, 32, 32)
    conv(32, 256, 32, 32, up=True)
    conv(32, 128, 64, 64)
    conv(32, 64, 128, 128)
    conv(32, 32, 256, 128)
    conv(32, 64, 128, 128, reflect=False)
if which == "rows":
    for impl in ("rows", "band"):
        conv(32, 256, 32, 32, up=True, impl=impl)
        conv(32, 256, 32, 16, impl=impl)
        conv(32, 128, 64, 64, up=True, impl=impl)
        conv(32, 128, 64, 32, impl=impl)
        conv(96, 256, 64, 64, reflect=False, impl=impl)
if which in ("band", "all"):
    for impl in ("band", "gather"):
        conv(32, 256, 32, 32, impl=impl)
        conv(32, 256, 32, 16, impl=impl)
        conv(32, 128, 64, 64, up=True, impl=impl)
        conv(32, 128, 64, 32, impl=impl)
        conv(32, 64, 128, 128, impl=impl)
        conv(32, 64, 128, 64, impl=impl)
        conv(32, 32, 256, 128, impl=impl)
        conv(96, 256, 64, 64, reflect=False, impl=impl)    # vgg conv1_2 on 3x32 images
        conv(96, 128, 128, 128, reflect=False, impl=impl)  # vgg conv2_2
        conv(96, 64, 256, 256, reflect=False, impl=impl)   # vgg conv3_x
        conv(96, 32, 512, 512, reflect=False, impl=impl)   # vgg conv4_x
