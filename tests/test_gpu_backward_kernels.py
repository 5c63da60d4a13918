"""Kernel-level parity of the training-step (backward) entry points on the B200: each against torch autograd
of a plain fp32 statement of the same op on the same bf16-rounded operands."""
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


def _close(a, b, rtol=2e-2, name=""):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-12
    assert err <= rtol * ref, f"{name}: max err {err:.4e} vs ref max {ref:.4e}"


@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (128, 128, 256), (1000, 256, 256), (8192, 256, 256), (4096, 1024, 256),
                                   (4096, 256, 1024), (3000, 768, 256), (512, 32, 128), (700, 8, 64), (2048, 256, 72)])
def test_wgrad_linear(M, N, K):
    ops = _ops()
    dY = _rand(M, N, seed=1).bfloat16().cuda()
    X = _rand(M, K, seed=2).bfloat16().cuda()
    dW = torch.zeros(N, K, device="cuda")
    ops.wgrad(dY, X, dW, M, N, K)
    ops.wgrad(dY, X, dW, M, N, K)  # accumulates
    ref = 2 * dY.float().T @ X.float()
    _close(dW, ref, 2e-3, "wgrad")
    db = torch.zeros(N, device="cuda")
    ops.colsum(dY, M, N, db)
    _close(db, dY.float().sum(0), 2e-3, "colsum")


@pytest.mark.parametrize("B,H,W,Cin,Cout,pad,up", [
    (2, 16, 16, 64, 64, "reflect", False), (1, 32, 32, 256, 128, "reflect", False), (2, 16, 16, 128, 128, "reflect", True),
    (1, 32, 32, 32, 32, "reflect", True), (2, 16, 16, 64, 128, "zeros", False), (2, 24, 40, 32, 8, "reflect", False),
    (1, 64, 64, 8, 32, "zeros", False)])
def test_wgrad_conv(B, H, W, Cin, Cout, pad, up):
    ops = _ops()
    hs, ws = (H // 2, W // 2) if up else (H, W)
    x = _rand(B, hs, ws, Cin, seed=3).bfloat16()
    dy = _rand(B, H, W, Cout, seed=4).bfloat16()
    w = torch.zeros(Cout, Cin, 3, 3, requires_grad=True)
    xi = x.float().permute(0, 3, 1, 2)
    if up:
        xi = F.interpolate(xi, scale_factor=2, mode="nearest")
    xi = F.pad(xi, (1, 1, 1, 1), mode="reflect" if pad == "reflect" else "constant")
    y = F.conv2d(xi, w)
    y.backward(dy.float().permute(0, 3, 1, 2))
    dW = torch.zeros(Cout, Cin, 3, 3, device="cuda")
    n_real = 3 if Cout == 8 else 0
    ops.wgrad(dy.cuda().view(-1, Cout), x.cuda(), dW, B * H * W, Cout, 9 * Cin,
              conv=dict(H=H, W=W, Cin=Cin, pad_mode=1 if pad == "reflect" else 0, upsample=up), n_real=n_real)
    ref = w.grad.clone()
    if n_real:
        ref[n_real:] = 0
    _close(dW, ref, 3e-3, "conv wgrad")


def test_gemm_training_epilogues():
    ops = _ops()
    M, N, K = 640, 256, 256
    A = _rand(M, K, seed=4).bfloat16().cuda()
    W = _rand(N, K, seed=5, scale=K ** -0.5)
    bias = _rand(N, seed=6)
    gate = _rand(M, N, seed=7).bfloat16()
    add = _rand(M, N, seed=8).bfloat16()
    pm = ops.pack_linear(W.cuda(), bias.cuda())
    acc = A.float().cpu() @ W.bfloat16().float().T + bias
    out = torch.empty(M, N, device="cuda")
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, pm, M, act=ops.ACT_GELU, out_f32=out, out_pre16=pre)
    _close(out, F.gelu(acc), 2e-3, "gelu")
    _close(pre, acc, 1e-2, "pre-activation copy")
    ops.gemm(A, pm, M, out_f32=out, gate=gate.cuda(), gate_mode=ops.GATE_RELU, add16=add.cuda())
    _close(out, acc * (gate.float() > 0) + add.float(), 2e-3, "relu gate + add")
    g = gate.float().clone().requires_grad_(True)
    F.gelu(g).sum().backward()
    ops.gemm(A, pm, M, out_f32=out, gate=gate.cuda(), gate_mode=ops.GATE_GELU)
    _close(out, acc * g.grad, 2e-3, "gelu' gate")
    rs = torch.tensor([0.0, 1.25, 1.25, 0.0, 1.25], device="cuda")
    res = _rand(M, N, seed=9)
    ops.gemm(A, pm, M, out_f32=out, res=res.cuda(), row_scale=rs, rows_per_scale=128)
    ref = res + acc * rs.cpu().repeat_interleave(128)[:, None]
    _close(out, ref, 2e-3, "row scale")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (1, 32, 32, 128, 256), (2, 12, 20, 32, 32), (1, 64, 64, 8, 32)])
def test_conv_dgrad_reflect(B, H, W, Cin, Cout):
    """Data gradient of a reflect-padded 3x3 conv = conv_full with flipped/transposed weights + reflect fold."""
    ops = _ops()
    w = _rand(Cout, Cin, 3, 3, seed=1, scale=(9 * Cin) ** -0.5).bfloat16().float()
    dy = _rand(B, H, W, Cout, seed=2).bfloat16()
    x = torch.zeros(B, Cin, H, W, requires_grad=True)
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = x.grad.permute(0, 2, 3, 1)
    wt = w.permute(1, 0, 2, 3).flip(2, 3).contiguous()  # [Cin, Cout, 3, 3]
    pm = ops.pack_conv3x3(wt.cuda(), None)
    Hp, Wp = H + 2, W + 2
    dxp = torch.empty(B * Hp * Wp, pm.n_pad, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy.cuda(), pm, B * Hp * Wp, out_bf16=dxp, conv=dict(H=Hp, W=Wp, Cin=Cout, pad_mode=0, full=True))
    dx = torch.empty(B, H, W, pm.n_pad, device="cuda", dtype=torch.bfloat16)
    ops.reflect_fold(dxp, None, dx, B, H, W, pm.n_pad)
    _close(dx[..., :Cin], ref, 2e-2, "reflect dgrad")


def test_reflect_fold_upsample_gate():
    ops = _ops()
    B, H, W, C = 2, 16, 24, 32
    dxp = _rand(B, H + 2, W + 2, C, seed=1).bfloat16()
    gate = _rand(B, H // 2, W // 2, C, seed=2).bfloat16()
    x = torch.zeros(B, C, H // 2, W // 2, requires_grad=True)
    y = F.pad(F.interpolate(torch.relu(x + gate.float().permute(0, 3, 1, 2)), scale_factor=2, mode="nearest"), (1, 1, 1, 1), mode="reflect")
    y.backward(dxp.float().permute(0, 3, 1, 2))
    out = torch.empty(B, H // 2, W // 2, C, device="cuda", dtype=torch.bfloat16)
    ops.reflect_fold(dxp.cuda(), gate.cuda(), out, B, H, W, C, upsample=True)
    _close(out, x.grad.permute(0, 2, 3, 1), 1e-2, "fold+upsample+relu")


def test_maxpool_bwd():
    ops = _ops()
    B, H, W, C = 2, 16, 12, 64
    pre = _rand(B, H, W, C, seed=1).bfloat16()
    dy = _rand(B, H // 2, W // 2, C, seed=2).bfloat16()
    x = pre.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.max_pool2d(torch.relu(x), 2).backward(dy.float().permute(0, 3, 1, 2))
    dx = torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    ops.maxpool2x2_bwd(torch.relu(pre.float()).bfloat16().cuda(), dy.cuda(), dx, B, H, W, C)
    _close(dx, x.grad.permute(0, 2, 3, 1), 1e-6, "maxpool bwd")


@pytest.mark.parametrize("C", [128, 256])
def test_layernorm_bwd(C):
    ops = _ops()
    rows = 1000
    x = _rand(rows, C, seed=1) * 2 + 0.5
    gamma, beta = _rand(C, seed=2) + 1, _rand(C, seed=3)
    dy = _rand(rows, C, seed=4).bfloat16()
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (C,), gr, br, 1e-5).backward(dy.float())
    acc0 = _rand(rows, C, seed=5)
    dx = acc0.clone().cuda()
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.layernorm_bwd(x.cuda(), gamma.cuda(), dy.cuda(), dx, dg, db, rows, C)
    _close(dx.cpu() - acc0, xr.grad, 1e-3, "ln dx")
    _close(dg, gr.grad, 1e-3, "ln dgamma")
    _close(db, br.grad, 1e-3, "ln dbeta")


@pytest.mark.parametrize("twice", [False, True])
@pytest.mark.parametrize("f32", [False, True])
def test_instnorm_bwd(twice, f32):
    ops = _ops()
    B, T, C = 2, 256, 64
    x = _rand(B, T, C, seed=1) * 1.5 + 0.3
    dy = _rand(B, T, C, seed=2)
    dy = dy if f32 else dy.bfloat16()
    xr = x.clone().requires_grad_(True)

    def inorm(t):
        return F.instance_norm(t.permute(0, 2, 1), eps=1e-5).permute(0, 2, 1)

    y = inorm(inorm(xr)) if twice else inorm(xr)
    y.backward(dy.float())
    coef = torch.empty(B, C, 4, device="cuda")
    dx16 = torch.empty(B, T, C, device="cuda", dtype=torch.bfloat16)
    acc = torch.zeros(B, T, C, device="cuda")
    ops.instnorm_bwd(x.cuda(), dy.cuda(), coef, B, T, C, twice=twice, dx_accum=acc, dx16=dx16)
    _close(acc, xr.grad, 2e-3, "in dx")
    _close(dx16, xr.grad, 1e-2, "in dx16")


def _attn_ref(q, k, v, table, B, H, W, heads, ws, shift, v2=None):
    """fp32 torch statement of the window-attention core on token-major [B*H*W, C] tensors (autograd-able)."""
    from oracle import master_oracle as O
    C = q.shape[1]
    hd = C // heads
    gm = O.window_gather_map(H, W, ws, shift)  # [nW, N]
    nW, N = gm.shape
    mask = O.shift_mask(H, W, ws, shift)
    idx = O.relative_position_index(ws)
    bias = table[idx].view(N, N, heads).permute(2, 0, 1)

    def win(t):
        return t.view(B, H * W, C)[:, gm.reshape(-1)].view(B * nW, N, heads, hd).permute(0, 2, 1, 3)

    qw, kw, vw = win(q) * hd ** -0.5, win(k), win(v)
    s = qw @ kw.transpose(-1, -2) + bias[None]
    if mask is not None:
        s = s.view(B, nW, heads, N, N) + mask.view(1, nW, 1, N, N)
        s = s.view(B * nW, heads, N, N)
    p = s.softmax(-1)

    def unwin(o):
        o = o.permute(0, 2, 1, 3).reshape(B, nW * N, C)
        out = torch.zeros(B, H * W, C, dtype=o.dtype)
        out[:, gm.reshape(-1)] = o
        return out.view(B * H * W, C)

    if v2 is None:
        return unwin(p @ vw)
    return unwin(p @ vw), unwin(p @ win(v2))


@pytest.mark.parametrize("ws", [8, 7])
@pytest.mark.parametrize("shift", [0, 4])
@pytest.mark.parametrize("dual", [False, True])
def test_window_attention_bwd(shift, dual, ws):
    """ws = 7: 49-token windows in 64 slots on a map that is a multiple of 7 (the training path pads the map itself)."""
    ops = _ops()
    B, H, W, heads, C = 2, 2 * ws, 3 * ws, 8, 256
    T = B * H * W
    q, k, v, v2 = (_rand(T, C, seed=s).bfloat16() for s in (1, 2, 3, 4))
    do, do2 = _rand(T, C, seed=5).bfloat16(), _rand(T, C, seed=6).bfloat16()
    table = _rand((2 * ws - 1) ** 2, heads, seed=7, scale=0.5)
    qr, kr, vr, v2r = (t.float().clone().requires_grad_(True) for t in (q, k, v, v2))
    tr = table.clone().requires_grad_(True)
    if dual:
        o, o2 = _attn_ref(qr, kr, vr, tr, B, H, W, heads, ws, shift, v2r)
        (o * do.float()).sum().add((o2 * do2.float()).sum()).backward()
    else:
        o = _attn_ref(qr, kr, vr, tr, B, H, W, heads, ws, shift)
        (o * do.float()).sum().backward()
    # forward parity of the reference statement itself (guards the test's own gather/mask logic)
    of = torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
    of2 = torch.empty(T, C, device="cuda", dtype=torch.bfloat16) if dual else None
    ops.window_attention(q.cuda(), k.cuda(), v.cuda(), of, table.cuda(), B, H, W, heads, ws, shift, C, C, C, C,
                         v2=v2.cuda() if dual else None, out2=of2)
    _close(of, o.detach(), 2e-2, "fwd")
    dq, dk, dv = (torch.empty(T, C, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    dv2 = torch.empty(T, C, device="cuda", dtype=torch.bfloat16) if dual else None
    dtab = torch.zeros((2 * ws - 1) ** 2, heads, device="cuda")
    ops.window_attention_bwd(q.cuda(), k.cuda(), v.cuda(), do.cuda(), dq, dk, dv, table.cuda(), dtab, B, H, W, heads, ws, shift,
                             C, C, C, C, C, C, C, v2=v2.cuda() if dual else None, dout2=do2.cuda() if dual else None, dv2=dv2)
    _close(dq, qr.grad, 3e-2, "dq")
    _close(dk, kr.grad, 3e-2, "dk")
    _close(dv, vr.grad, 3e-2, "dv")
    if dual:
        _close(dv2, v2r.grad, 3e-2, "dv2")
    _close(dtab, tr.grad, 3e-2, "dtable")


def test_blend_and_add_cast():
    ops = _ops()
    n = 4096 * 4
    gy, sg, qy = _rand(n, seed=1), _rand(n, seed=2), _rand(n, seed=3)
    gq = torch.empty(n, device="cuda")
    gs, gm = (torch.empty(n, device="cuda", dtype=torch.bfloat16) for _ in range(2))
    ops.blend_bwd(gy.cuda(), sg.cuda(), qy.cuda(), gq, gs, gm)
    _close(gq, gy * sg, 1e-6, "gquery")
    _close(gs, gy * qy, 1e-2, "gsigma")
    _close(gm, gy, 1e-2, "gmu")
    o32 = torch.empty(n, device="cuda")
    o16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    ops.add_cast(gy.cuda(), sg.cuda(), o32, o16)
    _close(o32, gy + sg, 1e-6, "add")
    _close(o16, gy + sg, 1e-2, "add16")


@pytest.mark.parametrize("sq_c,sq_s", [(False, False), (True, True)])
def test_loss_bwd(sq_c, sq_s):
    ops = _ops()
    B, T, C = 2, 320, 128
    fc, fs, fo = (torch.relu(_rand(B, T, C, seed=s) + 0.3).bfloat16() for s in (1, 2, 3))
    w = torch.tensor([0.7, 3.0])

    def stats(t):
        return t.float().mean(1), t.float().var(1, unbiased=False)

    fr = fo.float().clone().requires_grad_(True)
    tin = lambda t: F.instance_norm(t.permute(0, 2, 1), eps=1e-5)
    d = tin(fc.float()) - tin(fr)
    content = (d * d).mean() if sq_c else d.abs().mean()
    dm = fs.float().mean(1) - fr.mean(1)
    ds = fs.float().std(1) - fr.std(1)
    style = (dm * dm).mean() + (ds * ds).mean() if sq_s else dm.abs().mean() + ds.abs().mean()
    (w[0] * content + w[1] * style).backward()
    ref = fr.grad * (fo.float() > 0)
    dev = lambda t: t.contiguous().cuda()
    (mc, vc), (ms, vs), (mo, vo) = stats(fc), stats(fs), stats(fo)
    s = torch.empty(B, C, 2, device="cuda")
    dfo = torch.empty(B, T, C, device="cuda", dtype=torch.bfloat16)
    ops.loss_bwd(dev(fc), dev(fo), dev(mc), dev(vc), dev(mo), dev(vo), dev(ms), dev(vs), s, dev(w), B, T, C, sq_c, sq_s, dfo)
    _close(dfo, ref, 2e-2, "loss bwd")
    ops.nchw3_to_nhwc8  # exported


def test_nchw3_to_nhwc8():
    ops = _ops()
    g = _rand(2, 3, 8, 12, seed=1)
    out = torch.empty(2, 8, 12, 8, device="cuda", dtype=torch.bfloat16)
    ops.nchw3_to_nhwc8(g.cuda(), out, 2, 8, 12)
    _close(out[..., :3], g.permute(0, 2, 3, 1), 1e-2, "nhwc8")
    assert out[..., 3:].abs().max().item() == 0
