"""CPU experiment: how far do whole-step parameter gradients move when the CNN decoder and the VGG loss network round their
operands to bf16 (what the kernels do), against the pure fp32 oracle?  Decides the end-to-end gradient gate of
tests/test_gpu_train.py::test_full_training_step_vs_oracle.  Usage: python tools/debug/grad_gate_cpu.py [--cond]"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from mastermetastyletransfer_b200 import MasterStyleTransferModel, custom_loss, synthetic
from oracle import master_oracle as O
import test_gpu_train as T

cond = "--cond" in sys.argv
size = int(os.environ.get("SIZE", "64")); Bn = int(os.environ.get("BATCH", "1")); sq = os.environ.get("SQ", "0") == "1"
m = MasterStyleTransferModel(); synthetic.fill_state_dict_(m, 0)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
loss_fn = custom_loss("/nonexistent"); synthetic.fill_state_dict_(loss_fn, 1)
if cond:
    import conftest
    c_, s_ = synthetic.synthetic_images(Bn, size, seed=4)
    with torch.no_grad():
        o_ = O.full_forward(sd, c_, s_, 1)
    conftest.condition_vgg_(loss_fn.feature_extractor_model.features, torch.cat([c_, s_, o_]), float(os.environ.get("SHIFT", "1.0")))
vsd = {k[len("feature_extractor_model.features."):]: v.detach() for k, v in loss_fn.state_dict().items() if k.startswith("feature_extractor_model.features.")}
content, style = synthetic.synthetic_images(Bn, size, seed=4)

def run(emu):
    ps = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.startswith("swin_encoder.") else v.clone()) for k, v in sd.items()}
    if not emu:
        img = O.full_forward(ps, content, style, 1)
        t, c, s = O.overall_loss(vsd, content, style, img, 10.0, sq, sq)
    else:
        fc, fs = O.swin_encoder(ps, content, "swin_encoder."), O.swin_encoder(ps, style, "swin_encoder.")
        st = {n[len("style_transformer."):]: v for n, v in ps.items() if n.startswith("style_transformer.")}
        x = O.style_transformer(st, fc, fs, 1)
        img = T.emu_cnn_decoder(ps, x.permute(0, 3, 1, 2), "decoder.decoder.")
        tc, ts, to = T.emu_vgg_taps(vsd, content), T.emu_vgg_taps(vsd, style), T.emu_vgg_taps(vsd, img)
        c, s = O.content_loss(tc, to, sq), O.style_loss(ts, to, sq)
        t = c + 10.0 * s
    t.backward()
    return t.item(), c.item(), s.item(), {k: v.grad for k, v in ps.items() if v.requires_grad and v.grad is not None}

a = run(False); b = run(True)
print("loss fp32", a[:3], "emu", b[:3])
taps = O.vgg_taps(vsd, O.full_forward(sd, content, style, 1).detach())
for i, tp in enumerate(taps):
    v = tp.var(dim=(2, 3), unbiased=False)
    print(f"tap {i}: channel var min {v.min().item():.2e} median {v.median().item():.2e} dead(<1e-6) {(v < 1e-6).sum().item()}/{v.numel()}")
worst_c, worst_r = 1.0, 0.0
tot_a = torch.cat([g.flatten() for g in a[3].values()]); tot_b = torch.cat([b[3][k].flatten() for k in a[3]])
print("all params: rel", ((tot_a - tot_b).norm() / tot_a.norm()).item(), "cos", (torch.dot(tot_a, tot_b) / tot_a.norm() / tot_b.norm()).item())
scale = max(g.norm().item() for g in a[3].values())
for k in a[3]:
    ga, gb = a[3][k].flatten(), b[3][k].flatten()
    if ga.norm() < 1e-4 * scale: continue
    r = ((ga - gb).norm() / ga.norm()).item(); c = (torch.dot(ga, gb) / ga.norm() / gb.norm()).item()
    if c < worst_c: worst_c, wk = c, k
    worst_r = max(worst_r, r)
print("worst per-param cos", worst_c, wk, "worst rel", worst_r)
