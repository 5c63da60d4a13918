import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch, torch.nn.functional as F
from mastermetastyletransfer_b200 import custom_loss, synthetic, train_engine as te, engine
from mastermetastyletransfer_b200.style_transformer import packed_weights, workspace_of
from oracle import master_oracle as O
from test_gpu_train import rb
sq = len(sys.argv) > 1
dist = "euclidian_squared" if sq else "euclidian"
loss = custom_loss("/nonexistent", distance_content=dist, distance_style=dist)
synthetic.fill_state_dict_(loss, 1); loss = loss.cuda()
vsd = {k[len("feature_extractor_model.features."):]: v.detach() for k, v in loss.state_dict().items() if k.startswith("feature_extractor_model.features.")}
content, style = synthetic.synthetic_images(2, 64, seed=2)
out_img, _ = synthetic.synthetic_images(2, 64, seed=3)
content, style, out_img = content.cuda(), style.cuda(), out_img.cuda()
pre = {}
def emu_taps(x, keep=False):
    taps = [None] * 4
    for idx in O.VGG_CONVS:
        if idx in O.VGG_POOL_BEFORE: x = F.max_pool2d(x, 2)
        w = vsd[f"{idx}.weight"]
        z = F.conv2d(x, w if idx == 0 else rb(w), vsd[f"{idx}.bias"], padding=1)
        if keep: z.retain_grad(); pre[idx] = z
        x = rb(torch.relu(z))
        if idx in O.VGG_TAPS: taps[O.VGG_TAPS[idx]] = x
    return taps
o_ref = out_img.clone().requires_grad_(True)
tc, ts = [t.detach() for t in emu_taps(content)], [t.detach() for t in emu_taps(style)]
to = emu_taps(o_ref, keep=True)
lc = sum(((O._in_nchw(a) - O._in_nchw(b)).square().mean() if sq else (O._in_nchw(a) - O._in_nchw(b)).abs().mean()) for a, b in zip(tc, to))
ls = 0
for a, b in zip(ts, to):
    dm, ds = a.mean(dim=(2, 3)) - b.mean(dim=(2, 3)), a.std(dim=(2, 3)) - b.std(dim=(2, 3))
    ls = ls + ((dm.square().mean() + ds.square().mean()) if sq else (dm.abs().mean() + ds.abs().mean()))
(lc + 10 * ls).backward()
fe = loss.feature_extractor_model
w = packed_weights(fe, te.VggTrainWeights); wi = packed_weights(fe, engine.VggWeights)
ws = workspace_of(loss, out_img.device)
out3, saved = te.perceptual_loss_forward_train(w, wi, content, style, out_img, 10.0, sq, sq, ws)
print("loss mine", out3.tolist(), "emu", (lc + 10 * ls).item(), lc.item(), ls.item())
saved["debug"] = {}
coef2 = torch.tensor([1.0, 10.0], device="cuda")
dimg = te.perceptual_loss_backward(w, saved, coef2, ws)
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
for idx in reversed(O.VGG_CONVS):
    mine = saved["debug"][idx]
    ref = pre[idx].grad  # NCHW
    B, C, H, W = ref.shape
    mine = mine.view(B, H, W, -1)[..., :C].permute(0, 3, 1, 2)
    # forward act check
    a_m = saved["acts"][idx][0].view(B, H, W, -1)[..., :C].permute(0, 3, 1, 2).float()
    a_r = rb(torch.relu(pre[idx])).detach()
    print(idx, "grad rel", rel(mine, ref), "|ref|", ref.norm().item(), "act rel", rel(a_m, a_r))
print("dimg rel", rel(dimg, o_ref.grad))
print("---- per-tap kernel check on the pipeline's own tensors")
B = 2
for i, tp in enumerate(saved["taps"]):
    T, C = tp["T"], tp["C"]
    fc = tp["fc"].view(B, T, C).float(); fo = tp["fo"].view(B, T, C).float().clone().requires_grad_(True)
    fs_mean, fs_var = tp["mean_s"], tp["var_s"]
    tin = lambda t: F.instance_norm(t.permute(0, 2, 1), eps=1e-5)
    d = tin(fc) - tin(fo)
    cl = (d * d).mean() if sq else d.abs().mean()
    dm = fs_mean - fo.mean(1)
    ds = (fs_var * T / (T - 1)).sqrt() - fo.std(1)
    sl = (dm * dm).mean() + (ds * ds).mean() if sq else dm.abs().mean() + ds.abs().mean()
    (cl + 10 * sl).backward()
    ref = fo.grad * (fo.detach() > 0)
    s = torch.empty(B, C, 2, device="cuda"); dfo = torch.empty(B, T, C, device="cuda", dtype=torch.bfloat16)
    from mastermetastyletransfer_b200 import ops
    ops.loss_bwd(tp["fc"], tp["fo"], tp["mean_c"], tp["var_c"], tp["mean_o"], tp["var_o"], tp["mean_s"], tp["var_s"], s, coef2, B, T, C, sq, sq, dfo)
    print(i, T, C, "kernel vs torch rel", rel(dfo.view(B, T, C), ref), "|ref|", ref.norm().item(),
          "mean_o err", (tp["mean_o"] - fo.detach().mean(1)).abs().max().item(), "var_o err", (tp["var_o"] - fo.detach().var(1, unbiased=False)).abs().max().item(),
          "min var_o", tp["var_o"].min().item(), "min var_c", tp["var_c"].min().item())
