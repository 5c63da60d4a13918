"""Kernel-level parity on the B200: every C-ABI entry point against a plain fp32 PyTorch statement of
the same op on the same (bf16-rounded) operands, and the integer maps bit-exactly against the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from mastermetastyletransfer_b200 import ops
    return ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 16, 64), (256, 32, 128), (300, 64, 256), (1024, 384, 128),
                                   (4096, 256, 256), (2048, 1024, 256), (2048, 256, 1024), (777, 768, 256), (128, 512, 512), (8192, 256, 256), (20000, 512, 128)])
def test_gemm_plain(M, N, K):
    ops = _ops()
    A = _rand(M, K, seed=1).bfloat16().cuda()
    W = _rand(N, K, seed=2, scale=K ** -0.5)
    bias = _rand(N, seed=3)
    pm = ops.pack_linear(W.cuda(), bias.cuda())
    out32 = torch.empty(M, N, device="cuda")
    out16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, pm, M, out_f32=out32, out_bf16=out16)
    ref = A.float().cpu() @ W.bfloat16().float().T + bias
    torch.cuda.synchronize()
    assert torch.allclose(out32.cpu(), ref, atol=2e-3, rtol=2e-3), (out32.cpu() - ref).abs().max()
    assert torch.allclose(out16.float().cpu(), ref, atol=3e-2, rtol=1e-2)


def test_gemm_epilogues():
    ops = _ops()
    M, N, K = 512, 256, 256
    A = _rand(M, K, seed=4).bfloat16().cuda()
    W = _rand(N, K, seed=5, scale=K ** -0.5)
    bias = _rand(N, seed=6)
    res = _rand(M, N, seed=7)
    mul = _rand(M, N, seed=8)
    pm = ops.pack_linear(W.cuda(), bias.cuda())
    acc = A.float().cpu() @ W.bfloat16().float().T + bias
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A, pm, M, act=ops.ACT_GELU, out_f32=out)
    assert torch.allclose(out.cpu(), F.gelu(acc), atol=2e-3, rtol=2e-3)
    ops.gemm(A, pm, M, act=ops.ACT_RELU, out_f32=out)
    assert torch.allclose(out.cpu(), torch.relu(acc), atol=2e-3, rtol=2e-3)
    r = res.cuda()
    ops.gemm(A, pm, M, res=r, out_f32=r)  # in-place residual
    assert torch.allclose(r.cpu(), acc + res, atol=2e-3, rtol=2e-3)
    ops.gemm(A, pm, M, res=res.cuda(), mul=mul.cuda(), out_f32=out)
    assert torch.allclose(out.cpu(), res * mul + acc, atol=2e-3, rtol=2e-3)
    # strided A (slice of a fused buffer) and strided output
    big = _rand(M, 3 * K, seed=9).bfloat16().cuda()
    wide = torch.zeros(M, 3 * N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(big[:, K:], pm, M, lda=3 * K, out_bf16=wide[:, N:], ld_out16=3 * N)
    ref = big[:, K:2 * K].float().cpu() @ W.bfloat16().float().T + bias
    assert torch.allclose(wide[:, N:2 * N].float().cpu(), ref, atol=3e-2, rtol=1e-2)
    assert wide[:, :N].abs().max().item() == 0 and wide[:, 2 * N:].abs().max().item() == 0


@pytest.mark.parametrize("B,H,W,Cin,Cout,pad,up,relu", [
    (2, 16, 16, 64, 64, "reflect", False, True), (1, 32, 32, 256, 128, "reflect", False, True),
    (2, 16, 16, 128, 128, "reflect", True, True), (1, 32, 32, 32, 32, "reflect", True, True),
    (2, 16, 16, 64, 128, "zeros", False, True), (1, 8, 8, 512, 512, "zeros", False, True),
    (3, 12, 20, 32, 64, "zeros", False, False), (1, 128, 128, 64, 64, "reflect", True, True),
    (1, 256, 256, 32, 32, "reflect", False, True), (2, 30, 34, 64, 256, "zeros", False, True),
    (3, 128, 128, 64, 32, "zeros", False, True), (2, 256, 256, 32, 32, "zeros", True, True), (2, 6, 128, 64, 64, "reflect", False, False),
    (5, 3, 512, 32, 32, "reflect", False, True), (40, 8, 128, 32, 16, "zeros", False, False), (2, 64, 256, 64, 64, "reflect", True, True),
    # TMA-fed implicit GEMM (Cin % 64 == 0, whole image rows per tile): the CNN decoder's reflect layers and VGG's zero-padded ones
    (2, 64, 64, 128, 128, "reflect", False, True), (3, 32, 32, 256, 128, "reflect", False, True), (1, 64, 64, 128, 64, "reflect", False, True),
    (3, 32, 32, 512, 512, "zeros", False, True), (2, 16, 16, 512, 512, "zeros", False, True), (1, 256, 256, 64, 64, "zeros", False, True)])
@pytest.mark.parametrize("impl", ["gather", "band", "rows"])
def test_conv3x3(B, H, W, Cin, Cout, pad, up, relu, impl):
    """H, W are the conv's output size; with up=True the stored input is [B,H/2,W/2,Cin].
    impl: gathered implicit GEMM (gemm_tc.cu) or the halo-band kernel (conv_band.cu)."""
    ops = _ops()
    if impl == "band" and not ops.band_supported(Cout, Cin, H, W):
        pytest.skip("no band plan for this shape (halo + weight ring exceed shared memory)")
    if impl == "rows" and not ops.rows_supported(Cout, Cin, H, W):
        pytest.skip("row-streaming kernel: Cin in {32,64}, Cout <= 64, W % 128 == 0 only")
    hs, ws = (H // 2, W // 2) if up else (H, W)
    x = _rand(B, hs, ws, Cin, seed=10).bfloat16()
    wt = _rand(Cout, Cin, 3, 3, seed=11, scale=(9 * Cin) ** -0.5)
    bias = _rand(Cout, seed=12)
    pm = ops.pack_conv3x3(wt.cuda(), bias.cuda())
    out = torch.empty(B * H * W, Cout, device="cuda")
    ops.gemm(x.cuda(), pm, B * H * W, act=ops.ACT_RELU if relu else ops.ACT_NONE, out_f32=out,
             conv=dict(H=H, W=W, Cin=Cin, pad_mode=ops.PAD_REFLECT if pad == "reflect" else ops.PAD_ZERO, upsample=up, impl=impl))
    xi = x.float().permute(0, 3, 1, 2)
    if up:
        xi = xi.repeat_interleave(2, 2).repeat_interleave(2, 3)
    xi = F.pad(xi, (1, 1, 1, 1), mode="reflect" if pad == "reflect" else "constant")
    ref = F.conv2d(xi, wt.bfloat16().float(), bias)
    if relu:
        ref = torch.relu(ref)
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    assert torch.allclose(out.cpu(), ref, atol=3e-3, rtol=3e-3), (out.cpu() - ref).abs().max()


@pytest.mark.parametrize("B,H,W,Cin,Cout,pad,relu", [
    (2, 64, 64, 128, 128, "reflect", True),    # decoder.py:29-35 at 64x64
    (3, 32, 32, 256, 128, "reflect", True),    # decoder.py:25: two channel slices, 8 rows per unit
    (1, 128, 128, 64, 128, "zeros", True),     # VGG conv2_1: one 64-channel plane, 2 rows per unit
    (2, 128, 128, 128, 128, "zeros", False),   # VGG conv2_2 shape, no activation: 64-channel slices because W > 64
    (2, 64, 64, 256, 256, "zeros", True),      # VGG conv3_x: two output-channel tiles out of a 256-wide packed weight
    (5, 64, 64, 512, 512, "zeros", True),      # four slices, four channel tiles, units not a multiple of the SM count
    (1, 8, 64, 64, 128, "reflect", False), (7, 16, 64, 128, 384, "zeros", True),
    (3, 64, 64, 128, 64, "reflect", True), (2, 64, 64, 256, 64, "zeros", False)])   # 64 output channels: half-filled weight stages (decoder.py:37)
def test_conv3x3_channel_major(B, H, W, Cin, Cout, pad, relu):
    """conv_cm.cu (output channels = MMA M, one image row = MMA N, row-shifted taps) vs F.conv2d on the bf16-rounded operands;
    'auto' must pick it for these shapes, and its result must equal the gathered implicit GEMM's to bf16 rounding."""
    ops = _ops()
    assert ops.cm_supported(Cout, Cin, H, W)
    x = _rand(B, H, W, Cin, seed=20).bfloat16()
    wt = _rand(Cout, Cin, 3, 3, seed=21, scale=(9 * Cin) ** -0.5)
    bias = _rand(Cout, seed=22)
    pm = ops.pack_conv3x3(wt.cuda(), bias.cuda())
    conv = dict(H=H, W=W, Cin=Cin, pad_mode=ops.PAD_REFLECT if pad == "reflect" else ops.PAD_ZERO)
    act = ops.ACT_RELU if relu else ops.ACT_NONE
    outs = {}
    for impl in ("cm", "auto", "gather"):
        out = torch.full((B * H * W, pm.n_pad), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.gemm(x.cuda(), pm, B * H * W, act=act, out_bf16=out, conv=dict(conv, impl=impl))
        outs[impl] = out.float().cpu()[:, :Cout]
    xi = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect" if pad == "reflect" else "constant")
    ref = F.conv2d(xi, wt.bfloat16().float(), bias)
    if relu:
        ref = torch.relu(ref)
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    assert torch.equal(outs["cm"], outs["auto"])
    assert torch.allclose(outs["cm"], ref, atol=1e-2, rtol=1e-2), (outs["cm"] - ref).abs().max()
    assert torch.allclose(outs["cm"], outs["gather"], atol=1e-2, rtol=1e-2)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 128, 128), (3, 128, 128, 128, 128), (1, 32, 32, 256, 128)])
def test_conv3x3_channel_major_folded_upsample(B, H, W, Cin, Cout):
    """decoder.py:27-29: nearest-x2 upsample + reflect-padded conv; conv_cm.cu folds the upsample into its row fetch (the stored
    input is [B, H/2, W/2, Cin]).  Same result as the gathered GEMM on the materialised upsample, to bf16 rounding, and as F.conv2d."""
    ops = _ops()
    x = _rand(B, H // 2, W // 2, Cin, seed=23).bfloat16()
    wt = _rand(Cout, Cin, 3, 3, seed=24, scale=(9 * Cin) ** -0.5)
    bias = _rand(Cout, seed=25)
    pm = ops.pack_conv3x3(wt.cuda(), bias.cuda())
    out = torch.full((B * H * W, pm.n_pad), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(x.cuda(), pm, B * H * W, act=ops.ACT_RELU, out_bf16=out, conv=dict(H=H, W=W, Cin=Cin, pad_mode=ops.PAD_REFLECT, upsample=True, impl="cm"))
    xi = x.float().permute(0, 3, 1, 2).repeat_interleave(2, 2).repeat_interleave(2, 3)
    ref = torch.relu(F.conv2d(F.pad(xi, (1, 1, 1, 1), mode="reflect"), wt.bfloat16().float(), bias)).permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    assert torch.allclose(out.float().cpu()[:, :Cout], ref, atol=1e-2, rtol=1e-2), (out.float().cpu()[:, :Cout] - ref).abs().max()


@pytest.mark.parametrize("impl", ["gather", "band", "rows"])
def test_conv3x3_nchw_out(impl):
    ops = _ops()
    B, H, W, Cin = (3, 20, 256, 32) if impl == "rows" else (2, 32, 32, 32)
    x = _rand(B, H, W, Cin, seed=13).bfloat16()
    wt = _rand(3, Cin, 3, 3, seed=14, scale=(9 * Cin) ** -0.5)
    bias = _rand(3, seed=15)
    pm = ops.pack_conv3x3(wt.cuda(), bias.cuda())
    out = torch.empty(B, 3, H, W, device="cuda")
    ops.gemm(x.cuda(), pm, B * H * W, out_f32=out,
             conv=dict(H=H, W=W, Cin=Cin, pad_mode=ops.PAD_REFLECT, upsample=False, out_nchw=True, n_real=3, impl=impl))
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect"), wt.bfloat16().float(), bias)
    assert torch.allclose(out.cpu(), ref, atol=3e-3, rtol=3e-3)


@pytest.mark.parametrize("up", [False, True])
def test_conv3x3_u8_image_out(up):
    """MST_OUT_IMAGE_U8: the row-streaming conv storing (uint8) clip(x * 255, 0, 255) as an [B,H,W,3] image == the same conv's fp32
    NCHW output through mst_images_nchw_to_u8, bit for bit (values spread over < 0, [0, 1] and > 1); the other conv kernels refuse."""
    ops = _ops()
    B, H, W, Cin = 3, 20, 256, 32
    hs, wsz = (H // 2, W // 2) if up else (H, W)
    x = _rand(B, hs, wsz, Cin, seed=13).bfloat16().cuda()
    wt = _rand(3, Cin, 3, 3, seed=14, scale=(9 * Cin) ** -0.5)
    bias = torch.tensor([0.5, -0.2, 1.0])
    pm = ops.pack_conv3x3(wt.cuda(), bias.cuda())
    conv = dict(H=H, W=W, Cin=Cin, pad_mode=ops.PAD_REFLECT, upsample=up, n_real=3)
    out32 = torch.empty(B, 3, H, W, device="cuda")
    ops.gemm(x, pm, B * H * W, out_f32=out32, conv=dict(conv, out_nchw=True, impl="rows"))
    want = torch.empty(B, H, W, 3, dtype=torch.uint8, device="cuda")
    ops.images_nchw_to_u8(out32, want)
    got = torch.zeros(B, H, W, 3, dtype=torch.uint8, device="cuda")
    ops.gemm(x, pm, B * H * W, out_u8=got, conv=conv)
    assert torch.equal(got, want)
    frac0, frac255 = (want == 0).float().mean().item(), (want == 255).float().mean().item()
    assert 0.02 < frac0 < 0.9 and 0.02 < frac255 < 0.9  # both clip branches exercised
    with pytest.raises(ValueError):
        ops.gemm(x, pm, B * H * W, out_u8=got, out_f32=out32, conv=conv)
    from mastermetastyletransfer_b200 import _lib
    import ctypes as C
    g = _lib.MstGemm()
    g.A, g.Wt, g.bias, g.out_f32 = x.data_ptr(), pm.w.data_ptr(), pm.bias.data_ptr(), got.data_ptr()
    g.M, g.N, g.K, g.k_pad, g.lda = B * H * W, pm.n_pad, pm.K, pm.k_pad, pm.K
    g.ld_out32 = g.ld_out16 = g.ld_res = pm.n_pad
    g.a_mode, g.H, g.W, g.Cin, g.pad_mode, g.upsample, g.out_nchw, g.n_real = 1, H, W, Cin, 1, int(up), 2, 3
    st = torch.cuda.current_stream().cuda_stream
    assert _lib.lib().mst_gemm(C.byref(g), st) == -2 and _lib.lib().mst_conv3x3_band(C.byref(g), st) == -2  # MST_ERR_UNSUPPORTED


@pytest.mark.parametrize("H,ws,s", [(32, 8, 4), (64, 8, 4), (16, 8, 4), (8, 8, 4), (32, 7, 4), (64, 7, 4), (32, 7, 3), (64, 7, 3), (16, 7, 3), (32, 7, 0)])
def test_window_maps_bit_exact(H, ws, s, golden_dir):
    """Partition / shift / mask indexing of the CUDA kernels == oracle == reference-derived goldens, bit for bit."""
    import os
    from oracle import master_oracle as O
    ops = _ops()
    gather, labels, relidx = (t.cpu().long() for t in ops.window_maps(H, H, ws, s))
    Hp, Wp = O.padded_dims(H, H, ws)
    g = O.window_gather_map(H, H, ws, s)
    y, x = g // Wp, g % Wp
    mine = torch.where((y < H) & (x < H), y * H + x, torch.full_like(g, -1))
    assert torch.equal(gather, mine)
    assert torch.equal(relidx, O.relative_position_index(ws))
    lab = O.region_labels(H, H, ws, s)
    if lab is not None:
        # only label *differences* are observable (they decide the -100 mask): compare the masks
        assert torch.equal(labels.unsqueeze(1) != labels.unsqueeze(2), lab.unsqueeze(1) != lab.unsqueeze(2))
    gold = np.load(os.path.join(golden_dir, "maps.npz"))
    key = f"gather_{H}_{ws}_{s}"
    if key in gold:
        assert np.array_equal(gather.numpy(), gold[key])
        if f"mask_{H}_{ws}_{s}" in gold:
            assert np.array_equal((labels.unsqueeze(1) != labels.unsqueeze(2)).numpy().astype(np.uint8), gold[f"mask_{H}_{ws}_{s}"])


@pytest.mark.parametrize("ws,s,H,heads,dual", [(8, 4, 16, 8, False), (8, 4, 32, 8, True), (7, 3, 16, 8, False), (7, 0, 32, 4, False), (7, 3, 32, 4, False), (8, 4, 8, 8, False)])
def test_window_attention_core(ws, s, H, heads, dual):
    """Attention core on already-projected q/k/v against the oracle's softmax path (identity projections)."""
    from oracle import master_oracle as O
    ops = _ops()
    B, C = 2, heads * 32
    T = B * H * H
    q, k, v, v2 = (_rand(T, C, seed=20 + i).bfloat16() for i in range(4))
    table = _rand((2 * ws - 1) ** 2, heads, seed=30, scale=0.5)
    padv = [_rand(C, seed=40 + i, scale=0.3) for i in range(4)]
    out = torch.zeros(T, C, device="cuda", dtype=torch.bfloat16)
    out2 = torch.zeros(T, C, device="cuda", dtype=torch.bfloat16) if dual else None
    ops.window_attention(q.cuda(), k.cuda(), v.cuda(), out, table.cuda(), B, H, H, heads, ws, s, C, C, C, C,
                         v2=v2.cuda() if dual else None, out2=out2,
                         pad_q=padv[0].cuda(), pad_k=padv[1].cuda(), pad_v=padv[2].cuda(), pad_v2=padv[3].cuda() if dual else None)
    eye, zero = torch.eye(C), torch.zeros(C)

    def ref_one(vv, pv):
        # zero-padded tokens take the projection bias (= pad vectors): emulate with x - pad, bias = pad
        sh = lambda t, p: (t.float() - p).reshape(B, H, H, C)
        return O.window_attention(sh(q, padv[0]), sh(k, padv[1]), sh(vv, pv), eye, padv[0], eye, padv[1], eye, pv,
                                  eye, zero, table, ws, s, heads).reshape(T, C)
    ref = ref_one(v, padv[2])
    assert torch.allclose(out.float().cpu(), ref, atol=2e-2, rtol=2e-2), (out.float().cpu() - ref).abs().max()
    if dual:
        ref2 = ref_one(v2, padv[3])
        assert torch.allclose(out2.float().cpu(), ref2, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("ws,s,B,H,W,heads", [(8, 4, 2, 32, 32, 8), (8, 0, 1, 16, 32, 8), (7, 3, 2, 28, 28, 8), (8, 4, 1, 24, 24, 4),
                                              (7, 3, 3, 14, 21, 4), (8, 4, 3, 64, 64, 8)])
def test_window_attention_dual_tcgen05(ws, s, B, H, W, heads):
    """The dual (one softmax, two value tensors) passes on maps the windows tile run on attn_core.cu (TMA window loads, S / PV on
    tcgen05, P as a TMEM operand); against the oracle's attention with identity projections.  Odd window counts, rectangular maps,
    unshifted and shifted windows, strided (fused-buffer) operands."""
    from oracle import master_oracle as O
    ops = _ops()
    C = heads * 32
    T = B * H * W
    big = _rand(T, 4 * C, seed=60).bfloat16()  # q | k | v | v2 side by side: leading dimension 4C
    table = _rand((2 * ws - 1) ** 2, heads, seed=61, scale=0.5)
    bigc = big.cuda()
    out = torch.full((T, 2 * C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.window_attention(bigc[:, :C], bigc[:, C:2 * C], bigc[:, 2 * C:3 * C], out[:, :C], table.cuda(), B, H, W, heads, ws, s, 4 * C, 4 * C, 4 * C, 2 * C,
                         v2=bigc[:, 3 * C:], out2=out[:, C:])
    eye, zero = torch.eye(C), torch.zeros(C)
    q, k, v, v2 = (big[:, i * C:(i + 1) * C].float().reshape(B, H, W, C) for i in range(4))
    for vv, o in ((v, out[:, :C]), (v2, out[:, C:])):
        ref = O.window_attention(q, k, vv, eye, zero, eye, zero, eye, zero, eye, zero, table, ws, s, heads).reshape(T, C)
        assert torch.allclose(o.float().cpu(), ref, atol=2e-2, rtol=2e-2), (o.float().cpu() - ref).abs().max()


@pytest.mark.parametrize("C", [128, 256, 512])
def test_layernorm(C):
    ops = _ops()
    rows = 1000
    x, g, b = _rand(rows, C, seed=50, scale=2.0) + 0.5, 1 + 0.1 * _rand(C, seed=51), 0.1 * _rand(C, seed=52)
    y = torch.empty(rows, C, device="cuda", dtype=torch.bfloat16)
    ops.layernorm(x.cuda(), g.cuda(), b.cuda(), y, rows, C)
    ref = F.layer_norm(x, (C,), g, b)
    assert torch.allclose(y.float().cpu(), ref, atol=2e-2, rtol=1e-2)


def test_patch_merge_layernorm():
    ops = _ops()
    B, H, W, C = 2, 8, 12, 128
    x, g, b = _rand(B, H, W, C, seed=53), 1 + 0.1 * _rand(4 * C, seed=54), 0.1 * _rand(4 * C, seed=55)
    y = torch.empty(B * H * W // 4, 4 * C, device="cuda", dtype=torch.bfloat16)
    ops.patch_merge_layernorm(x.cuda(), g.cuda(), b.cuda(), y, B, H, W, C)
    cat = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    ref = F.layer_norm(cat, (4 * C,), g, b).reshape(-1, 4 * C)
    assert torch.allclose(y.float().cpu(), ref, atol=2e-2, rtol=1e-2)


@pytest.mark.parametrize("twice", [False, True])
def test_instance_norm(twice):
    from oracle import master_oracle as O
    ops = _ops()
    B, H, C = 3, 16, 256
    x = _rand(B, H, H, C, seed=56, scale=1.7) + 0.3
    mean, rstd = torch.empty(B, C, device="cuda"), torch.empty(B, C, device="cuda")
    y32 = torch.empty(B, H, H, C, device="cuda")
    y16 = torch.empty(B, H, H, C, device="cuda", dtype=torch.bfloat16)
    xc = x.cuda()
    ops.instnorm_stats(xc, mean, rstd, B, H * H, C, twice=twice)
    ops.instnorm_apply(xc, mean, rstd, B, H * H, C, y16=y16, y32=y32)
    ref = O.instance_norm_bhwc(x)
    if twice:
        ref = O.instance_norm_bhwc(ref)
    assert torch.allclose(y32.cpu(), ref, atol=2e-5, rtol=1e-5), (y32.cpu() - ref).abs().max()
    assert torch.allclose(y16.float().cpu(), ref, atol=2e-2, rtol=1e-2)


@pytest.mark.parametrize("B,T,C,twice,affine,n_pad", [(3, 1024, 256, True, False, 0), (2, 1024, 256, False, True, 0), (2, 256, 128, True, True, 0),
                                                      (3, 576, 256, False, False, 49), (2, 1024, 256, False, True, 64), (1, 1600, 64, False, False, 0),
                                                      (2, 49, 256, True, False, 0), (1, 1021, 64, False, True, 0), (1, 4096, 256, False, False, 0)])
def test_instnorm_one_call(B, T, C, twice, affine, n_pad):
    """mst_instnorm (statistics + application; csrc/instnorm_fused.cu: the slice through TMA into shared memory, read once) == the
    two-kernel sequence mst_instnorm_stats_affine -> mst_instnorm_apply_affine bit for bit: mean, rstd, the normalised padding value
    and the bf16 result; and against nn.InstanceNorm semantics in fp32.  Includes box heights other than 256 (576 = 3 x 192,
    1600 = 8 x 200, 49 = one box), shapes the fused kernel does not take (T = 1021, a prime: no box height; T = 4096: too large), affine
    and padded variants."""
    ops = _ops()
    from mastermetastyletransfer_b200 import _lib
    x = (_rand(B, T, C, seed=156, scale=1.7) + 0.3).cuda()
    g = (1 + 0.2 * _rand(C, seed=157)).cuda() if affine else None
    be = (0.3 * _rand(C, seed=158)).cuda() if affine else None
    pad_val = (0.5 * _rand(C, seed=159)).cuda() if n_pad else None
    outs = []
    for fused in (False, True):
        mean, rstd = torch.full((B, C), float("nan"), device="cuda"), torch.full((B, C), float("nan"), device="cuda")
        pad_norm = torch.full((B, C), float("nan"), device="cuda") if n_pad else None
        y16 = torch.full((B, T, C), float("nan"), device="cuda", dtype=torch.bfloat16)
        if fused:
            ops.instnorm(x, mean, rstd, y16, B, T, C, twice=twice, gamma=g, beta=be, n_pad=n_pad, pad_val=pad_val, pad_norm=pad_norm)
        else:
            if n_pad:
                ops.instnorm_stats_padded(x, mean, rstd, B, T, C, n_pad, pad_val, pad_norm=pad_norm, gamma=g, beta=be)
            else:
                ops.instnorm_stats(x, mean, rstd, B, T, C, twice=twice, gamma=g)
            ops.instnorm_apply(x, mean, rstd, B, T, C, y16=y16, beta=be)
        outs.append((mean, rstd, y16, pad_norm))
    for name, a, b_ in zip(("mean", "rstd", "y16", "pad_norm"), outs[0], outs[1]):
        assert (a is None and b_ is None) or torch.equal(a, b_), name
    assert bool(_lib.lib().mst_instnorm_fused_supported(T, C)) == (T not in (1021, 4096))
    # semantics (fp32): biased variance over the T tokens plus n_pad tokens of value pad_val
    xd = x.double()
    if n_pad:
        xd = torch.cat([xd, pad_val.double().view(1, 1, C).expand(B, n_pad, C)], 1)
    m, var = xd.mean(1, keepdim=True), xd.var(1, unbiased=False, keepdim=True)
    gd = g.double() if affine else 1.0
    bd = be.double() if affine else 0.0
    y = (xd - m) / torch.sqrt(var + 1e-5) * gd + bd
    if twice:
        m2, v2 = y.mean(1, keepdim=True), y.var(1, unbiased=False, keepdim=True)
        y = (y - m2) / torch.sqrt(v2 + 1e-5) * gd + bd
    assert torch.allclose(outs[1][2].double(), y[:, :T], atol=3e-2, rtol=1e-2)


@pytest.mark.parametrize("S,exact", [(64, False), (128, False), (64, True), (40, False)])
def test_patch_embed(S, exact):
    """tv swin features[0] + (fused) norm1 of the first block.  The tensor-core path (S % 64 == 0) rounds the conv weights to
    bf16 like every other layer; exact / other sizes run the fp32 kernel."""
    ops = _ops()
    B = 3
    img = _rand(B, 3, S, S, seed=57)
    w, b = _rand(128, 3, 4, 4, seed=58, scale=48 ** -0.5), 0.05 * _rand(128, seed=59)
    g, beta = 1 + 0.1 * _rand(128, seed=60), 0.1 * _rand(128, seed=61)
    g1, beta1 = 1 + 0.1 * _rand(128, seed=62), 0.1 * _rand(128, seed=63)
    out = torch.empty(B, S // 4, S // 4, 128, device="cuda")
    y16 = torch.empty(B, S // 4, S // 4, 128, device="cuda", dtype=torch.bfloat16)
    ops.patch_embed(img.cuda(), w.cuda(), b.cuda(), g.cuda(), beta.cuda(), out, B, S, gamma1=g1.cuda(), beta1=beta1.cuda(), y16=y16, exact=exact)
    tc = S % 64 == 0 and not exact
    ref = F.layer_norm(F.conv2d(img, w.bfloat16().float() if tc else w, b, stride=4).permute(0, 2, 3, 1), (128,), g, beta)
    assert torch.allclose(out.cpu(), ref, atol=2e-4, rtol=1e-4), (out.cpu() - ref).abs().max()
    ref1 = F.layer_norm(ref, (128,), g1, beta1)
    assert torch.allclose(y16.float().cpu(), ref1, atol=2e-2, rtol=1e-2)
    if tc:  # distance to the un-rounded fp32 layer: bf16 weight rounding only
        full = F.layer_norm(F.conv2d(img, w, b, stride=4).permute(0, 2, 3, 1), (128,), g, beta)
        assert (out.cpu() - full).abs().max() < 3e-2
    out2 = torch.empty_like(out)
    ops.patch_embed(img.cuda(), w.cuda(), b.cuda(), g.cuda(), beta.cuda(), out2, B, S, exact=exact)  # without the fused norm1
    assert torch.equal(out2, out)


@pytest.mark.parametrize("S,normalize", [(64, True), (48, True), (128, False)])
def test_patch_embed_u8_images(S, normalize):
    """mst_patch_embed_ln_u8: uint8 [B,S,S,3] images with ToTensor + Normalize folded into the image loads == mst_images_u8_to_nchw
    followed by mst_patch_embed_ln, bit for bit (same kernel, same fp32 image values); S % 16 != 0 is refused."""
    ops = _ops()
    B = 3
    gen = torch.Generator().manual_seed(71)
    img8 = torch.randint(0, 256, (B, S, S, 3), generator=gen, dtype=torch.uint8).cuda()
    mean = ops.IMAGENET_MEAN if normalize else None
    img32 = torch.empty(B, 3, S, S, device="cuda")
    ops.images_u8_to_nchw(img8, img32, mean)
    w, b = _rand(128, 3, 4, 4, seed=58, scale=48 ** -0.5).cuda(), (0.05 * _rand(128, seed=59)).cuda()
    g, beta = (1 + 0.1 * _rand(128, seed=60)).cuda(), (0.1 * _rand(128, seed=61)).cuda()
    g1, beta1 = (1 + 0.1 * _rand(128, seed=62)).cuda(), (0.1 * _rand(128, seed=63)).cuda()
    outs = []
    for img in (img32, img8):
        out = torch.empty(B, S // 4, S // 4, 128, device="cuda")
        y16 = torch.empty(B, S // 4, S // 4, 128, device="cuda", dtype=torch.bfloat16)
        ops.patch_embed(img, w, b, g, beta, out, B, S, gamma1=g1, beta1=beta1, y16=y16,
                        u8_mean=mean, u8_std=ops.IMAGENET_STD if normalize else None)
        outs.append((out, y16))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert outs[1][0].abs().max() > 0.1
    assert not ops.patch_embed_u8_supported(40)
    with pytest.raises(Exception):
        ops.patch_embed(torch.zeros(1, 40, 40, 3, dtype=torch.uint8, device="cuda"), w, b, g, beta, torch.empty(1, 10, 10, 128, device="cuda"), 1, 40)


def test_upsample2x_nhwc():
    ops = _ops()
    B, H, W, C = 3, 5, 7, 128
    x = _rand(B, H, W, C, seed=91).bfloat16().cuda()
    y = torch.empty(B, 2 * H, 2 * W, C, device="cuda", dtype=torch.bfloat16)
    ops.upsample2x_nhwc(x, y, B, H, W, C)
    assert torch.equal(y, x.repeat_interleave(2, 1).repeat_interleave(2, 2))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_token_map_pad_crop(dtype):
    """mst_token_map_copy: F.pad(x, (0,0,0,pad_r,0,pad_b)) / x[:, :H, :W, :] of style_transformer.py:77-87, :230-232 -- bit-exact."""
    from mastermetastyletransfer_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, H, W, Hp, Wp, C = 3, 16, 18, 21, 21, 64
    x = torch.randn(B, H, W, C, generator=g).to(dtype).cuda()
    xp = torch.full((B, Hp, Wp, C), 7.0, dtype=dtype, device="cuda")
    ops.token_map_copy(x, xp, B, H, W, Hp, Wp)
    assert torch.equal(xp, F.pad(x, (0, 0, 0, Wp - W, 0, Hp - H)))
    y = torch.randn(B, Hp, Wp, C, generator=g).to(dtype).cuda()
    yc = torch.empty(B, H, W, C, dtype=dtype, device="cuda")
    ops.token_map_copy(y, yc, B, Hp, Wp, H, W)
    assert torch.equal(yc, y[:, :H, :W])
    if dtype == torch.float32:
        acc = torch.randn(B, H, W, C, generator=g).cuda()
        want = acc + y[:, :H, :W]
        ops.token_map_copy(y, acc, B, Hp, Wp, H, W, accumulate=True)
        assert torch.equal(acc, want)
    else:
        with pytest.raises(TypeError):
            ops.token_map_copy(y, yc, B, Hp, Wp, H, W, accumulate=True)


@pytest.mark.parametrize("normalize", [True, False])
def test_u8_image_boundary_bit_exact(normalize):
    """mst_images_u8_to_nchw / mst_images_nchw_to_u8 against the oracle's restatement of ToTensor + Normalize and of
    np.clip(x*255, 0, 255).astype(uint8) (test_model.py:39-48, :207): bit-exact, every uint8 value, out-of-range floats."""
    from mastermetastyletransfer_b200 import ops
    from oracle import master_oracle as O
    g = torch.Generator().manual_seed(1)
    B, H, W = 3, 20, 256
    img = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    img[0, 0, :, :] = torch.arange(256, dtype=torch.uint8).view(256, 1)
    out = torch.empty(B, 3, H, W, device="cuda")
    ops.images_u8_to_nchw(img.cuda(), out, ops.IMAGENET_MEAN if normalize else None)
    assert torch.equal(out.cpu(), O.images_u8_to_tensor(img, mean=O.IMAGENET_MEAN if normalize else None))
    x = torch.randn(B, 3, H, W, generator=g) * 0.6 + 0.5
    x[0, 0, 0, :256] = torch.arange(256, dtype=torch.float32) / 255
    x[0, 1, 0, :256] = (torch.arange(256, dtype=torch.float32) + 0.999) / 255
    x[0, 2, 0, :4] = torch.tensor([-1e9, 1e9, -0.0, 2.0])
    back = torch.empty(B, H, W, 3, dtype=torch.uint8, device="cuda")
    ops.images_nchw_to_u8(x.cuda(), back)
    assert torch.equal(back.cpu(), O.tensor_to_images_u8(x))
    with pytest.raises(ValueError):
        ops.images_nchw_to_u8(x.cuda(), torch.empty(B, H, W, 4, dtype=torch.uint8, device="cuda"))


def test_bad_arguments_raise():
    ops = _ops()
    A = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    pm = ops.pack_linear(torch.zeros(16, 64, device="cuda"))
    with pytest.raises(ValueError):
        ops.gemm(A, pm, 128)  # no output
    with pytest.raises(TypeError):
        ops.gemm(A.float(), pm, 128, out_f32=torch.empty(128, 16, device="cuda"))
    with pytest.raises(RuntimeError):
        ops.gemm(A.cpu(), pm, 128, out_f32=torch.empty(128, 16, device="cuda"))


@pytest.mark.parametrize("M,C", [(128, 256), (1000, 256), (4096, 128), (333, 128), (40000, 256)])
def test_mlp_fused(M, C):
    """fc1 + GELU(erf) + fc2 + residual in one kernel, against fp32 torch on the operands as the kernel rounds them: bf16 input and
    W1, and fp16 for the hidden activation and W2 (the GELU runs in packed fp16 and the second GEMM takes fp16 operands)."""
    ops = _ops()
    A = _rand(M, C, seed=70).bfloat16()
    w1, b1 = _rand(4 * C, C, seed=71, scale=C ** -0.5), 0.1 * _rand(4 * C, seed=72)
    w2, b2 = _rand(C, 4 * C, seed=73, scale=(4 * C) ** -0.5), 0.1 * _rand(C, seed=74)
    res = _rand(M, C, seed=75)
    pm = ops.pack_mlp(w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda())
    out32 = torch.empty(M, C, device="cuda")
    out16 = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    ops.mlp_fused(A.cuda(), pm, M, res=res.cuda(), out_f32=out32, out_bf16=out16)
    h = F.gelu(A.float() @ w1.bfloat16().float().T + b1).half().float()
    ref = res + h @ w2.half().float().T + b2
    torch.cuda.synchronize()
    assert torch.allclose(out32.cpu(), ref, atol=4e-3, rtol=4e-3), (out32.cpu() - ref).abs().max()
    assert torch.allclose(out16.float().cpu(), ref, atol=3e-2, rtol=1e-2)
    r = res.cuda()
    ops.mlp_fused(A.cuda(), pm, M, res=r, out_f32=r)  # in-place residual stream
    assert torch.allclose(r.cpu(), ref, atol=4e-3, rtol=4e-3)


@pytest.mark.parametrize("M,C,ln,mul", [(128, 256, True, False), (1000, 256, False, False), (4096, 128, True, False), (333, 128, False, True),
                                       (40000, 256, True, False), (40000, 128, True, False), (20000, 256, False, True), (19201, 128, False, False)])
def test_proj_mlp_fused(M, C, ln, mul):
    """Attention-output half of a transformer block in one kernel (MstMlp::pre): x1 = res (*mul) + A.Wp^T + bp;
    out = x1 + mlp([LN](x1)).  Reference: fp32 torch on bf16-rounded operands, with the MLP input and the hidden activation
    rounded to bf16 where the kernel rounds them."""
    ops = _ops()
    A = _rand(M, C, seed=80).bfloat16()
    wp, bp = _rand(C, C, seed=81, scale=C ** -0.5), 0.1 * _rand(C, seed=82)
    w1, b1 = _rand(4 * C, C, seed=83, scale=C ** -0.5), 0.1 * _rand(4 * C, seed=84)
    w2, b2 = _rand(C, 4 * C, seed=85, scale=(4 * C) ** -0.5), 0.1 * _rand(C, seed=86)
    res = _rand(M, C, seed=87) + 0.5  # non-zero row means: exercises the pairwise mean/M2 merge of the fused LayerNorm
    m = (1 + 0.3 * _rand(M, C, seed=88)) if mul else None
    g, be = (1 + 0.1 * _rand(C, seed=89), 0.1 * _rand(C, seed=90)) if ln else (None, None)
    pm = ops.pack_mlp(w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), wpre=wp.cuda(), bpre=bp.cuda())
    x = res.cuda()
    out16 = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    ops.mlp_fused(A.cuda(), pm, M, res=x, out_f32=x, out_bf16=out16, pre=True, mul=m.cuda() if mul else None,
                  ln_g=g.cuda() if ln else None, ln_b=be.cuda() if ln else None)  # in-place residual stream
    v = A.float() @ wp.bfloat16().float().T + bp
    x1 = res * m + v if mul else res + v
    xin = F.layer_norm(x1, (C,), g, be) if ln else x1
    h = F.gelu(xin.bfloat16().float() @ w1.bfloat16().float().T + b1).half().float()  # fp16 hidden activation and W2 (see test_mlp_fused)
    ref = x1 + h @ w2.half().float().T + b2
    torch.cuda.synchronize()
    err = (x.cpu() - ref).abs().max()
    assert torch.allclose(x.cpu(), ref, atol=6e-3, rtol=6e-3), err
    assert torch.allclose(out16.float().cpu(), ref, atol=4e-2, rtol=1e-2)
    # out of place (the style transformer's first layer reads the encoder's feature map as the residual and writes its own buffer):
    # same bits as in place, and the residual source is left untouched
    res_dev, y = res.cuda(), torch.full((M, C), float("nan"), device="cuda")
    ops.mlp_fused(A.cuda(), pm, M, res=res_dev, out_f32=y, pre=True, mul=m.cuda() if mul else None,
                  ln_g=g.cuda() if ln else None, ln_b=be.cuda() if ln else None)
    assert torch.equal(y, x) and torch.equal(res_dev.cpu(), res)
    # MstMlp::lnn_g / lnn_b / lnn_rows: out_bf16 = LayerNorm(out) for the next block on the first lnn_rows rows (the others keep the
    # bf16 copy), from the tile-end epilogue.  out_f32 keeps its bits; the bf16 tensor matches a separate LayerNorm of it (different
    # summation order: one bf16 ulp of slack).
    assert ops.mlp_next_ln_supported(C)
    g2, be2 = (1 + 0.1 * _rand(C, seed=91)).cuda(), (0.1 * _rand(C, seed=92)).cuda()
    for inplace, rows in ((True, None), (False, M // 3), (True, 128 * (M // 256))):
        src = res.cuda()
        y2 = src if inplace else torch.empty(M, C, device="cuda")
        ln16 = torch.full((M, C), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.mlp_fused(A.cuda(), pm, M, res=src, out_f32=y2, out_bf16=ln16, pre=True, mul=m.cuda() if mul else None,
                      ln_g=g.cuda() if ln else None, ln_b=be.cuda() if ln else None, next_ln=(g2, be2) if rows is None else (g2, be2, rows))
        assert torch.equal(y2, x)
        n_ln = M if rows is None or rows == 0 else rows
        want = F.layer_norm(x[:n_ln], (C,), g2, be2)
        assert torch.allclose(ln16[:n_ln].float(), want, atol=2e-2, rtol=1e-2), (ln16[:n_ln].float() - want).abs().max()
        assert torch.equal(ln16[n_ln:], out16[n_ln:])   # plain bf16 copy of out beyond lnn_rows
    with pytest.raises(ValueError):
        ops.mlp_fused(A.cuda(), pm, M, res=x, out_f32=x, pre=True, next_ln=(g2, be2))   # needs out_bf16


@pytest.mark.parametrize("B,H,C,ws,shift", [(1, 8, 256, 8, 0), (2, 32, 256, 8, 4), (1, 24, 256, 8, 4), (2, 32, 256, 7, 4), (1, 16, 256, 7, 3),
                                            (2, 32, 128, 7, 0), (1, 64, 128, 7, 3), (1, 16, 128, 8, 4), (3, 64, 256, 7, 4)])
def test_attn_block_fused_vs_oracle(B, H, C, ws, shift):
    """csrc/attn_fused.cu: the three projections + shifted-window attention in one kernel (q, k, v stay on chip) against the fp32
    oracle restatement of codes/style_transformer.py:77-155 on the bf16-rounded operands, and its projected q | k | v (test hook)
    against a plain matmul.  Geometries: unshifted / shifted (wrap-around windows take the cp.async path, the others one TMA box
    per k-block), 8x8 and zero-padded 7x7 windows, an odd number of windows (24x24 map: nine), both channel counts."""
    from oracle import master_oracle as O
    ops = _ops()
    heads = C // 32
    g = torch.Generator().manual_seed(100 * H + ws + shift)
    x = torch.randn(B, H, H, C, generator=g)
    wq, wk, wv = (torch.randn(C, C, generator=g) * (0.7 / C ** 0.5) for _ in range(3))
    bq, bk, bv = (torch.randn(C, generator=g) * 0.2 for _ in range(3))
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5
    T = B * H * H
    x16 = x.cuda().bfloat16().view(T, C).contiguous()
    pk = ops.pack_attn_qkv(wq.cuda(), wk.cuda(), wv.cuda(), bq.cuda(), bk.cuda(), bv.cuda(), heads)
    out = torch.zeros(T, C, dtype=torch.bfloat16, device="cuda")
    dbg = torch.zeros(T, 3 * C, dtype=torch.bfloat16, device="cuda")
    ops.attn_block(x16, pk, table.cuda(), out, B, H, H, ws, shift, dbg_qkv=dbg)
    torch.cuda.synchronize()
    xr = x16.float().cpu()
    r = lambda t: t.bfloat16().float()
    qkv_ref = torch.cat([xr @ r(wq).t() + bq, xr @ r(wk).t() + bk, xr @ r(wv).t() + bv], 1)
    assert (dbg.float().cpu() - qkv_ref).abs().max().item() <= 0.03  # bf16 rounding of O(1) values
    xw = O._to_windows(xr.view(B, H, H, C), ws, shift)
    q, k, v = (r(torch.nn.functional.linear(xw, r(w_), b_)) for w_, b_ in ((wq, bq), (wk, bk), (wv, bv)))
    p = O._softmax_probs(q, k, heads, O._bias_from_table(table, ws), O.shift_mask(H, H, ws, shift), B)
    ref = O._from_windows(O._apply_probs(p, v, heads), B, H, H, ws, shift).reshape(T, C)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 1.5e-2 * max(1.0, ref.abs().max().item()), err


def test_attn_block_equals_unfused_kernels():
    """The fused kernel and the kernels it replaces (QKV GEMM + window_attn_kernel) agree to bf16 rounding on a benched shape."""
    ops = _ops()
    B, H, C, ws, shift, heads = 4, 32, 256, 8, 4, 8
    g = torch.Generator().manual_seed(3)
    T = B * H * H
    x16 = torch.randn(T, C, generator=g).cuda().bfloat16()
    w3 = [(torch.randn(C, C, generator=g) * (0.7 / 16)).cuda() for _ in range(3)]
    b3 = [(torch.randn(C, generator=g) * 0.2).cuda() for _ in range(3)]
    table = (torch.randn(225, heads, generator=g) * 0.5).cuda()
    pm = ops.pack_linear(torch.cat(w3, 0), torch.cat(b3, 0))
    qkv = torch.empty(T, 3 * C, dtype=torch.bfloat16, device="cuda")
    o_ref = torch.zeros(T, C, dtype=torch.bfloat16, device="cuda")
    ops.gemm(x16, pm, T, out_bf16=qkv)
    ops.window_attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o_ref, table, B, H, H, heads, ws, shift, 3 * C, 3 * C, 3 * C, C)
    o = torch.zeros(T, C, dtype=torch.bfloat16, device="cuda")
    ops.attn_block(x16, ops.pack_attn_qkv(*w3, *b3, heads), table, o, B, H, H, ws, shift)
    torch.cuda.synchronize()
    assert (o.float() - o_ref.float()).abs().max().item() <= 1e-2


@pytest.mark.parametrize("H,W", [(480, 640), (333, 500), (700, 1024), (100, 120), (512, 512)])
def test_gpu_train_transform_bit_exact_vs_reference_pipeline(H, W):
    """the reference's training-image transform (codes/get_dataloader.py:30-36: ToPILImage -> Resize((512,512)) ->
    RandomCrop((256,256)) -> ToTensor -> Normalize) as ONE kernel on the decoded uint8 image, against the torchvision / Pillow
    pipeline itself (oracle.train_transform) with the same crop window: BIT-EXACT (integer resample arithmetic restated from
    Pillow, IEEE fp32 division / subtraction in torchvision's order).  Shrinking (antialiased), enlarging, identity sizes; crop
    windows at the borders and drawn like RandomCrop does."""
    from mastermetastyletransfer_b200.data import GpuTrainTransform
    from oracle import master_oracle as O
    rng = np.random.default_rng(H * 7 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    tf = GpuTrainTransform("cuda")
    torch.manual_seed(H + W)
    windows = [(0, 0), (256, 256), (0, 256), tf.crop_params(), tf.crop_params()]
    for top, left in windows:
        out = tf(img, top_left=(top, left)).cpu()
        ref = O.train_transform(img, top, left)
        assert out.shape == ref.shape == (3, 256, 256)
        assert torch.equal(out, ref), (H, W, top, left, (out - ref).abs().max().item())
    # the random path consumes the generator exactly like the reference's Compose
    from torchvision import transforms
    compose = transforms.Compose([transforms.ToPILImage(), transforms.Resize((512, 512)), transforms.RandomCrop((256, 256)),
                                  transforms.ToTensor(), transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    torch.manual_seed(5)
    ref = compose(img)
    torch.manual_seed(5)
    assert torch.equal(tf(img).cpu(), ref)
