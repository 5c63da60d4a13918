"""Error statistics of the CUDA path against the CPU oracle (mean / p99.9 / max of |diff| / range) for the test shapes."""
import sys, torch
sys.path.insert(0, ".")
from mastermetastyletransfer_b200 import MasterStyleTransferModel, synthetic
from oracle import master_oracle as O

def stats(out, ref):
    d = (out.cpu().float() - ref).abs().flatten() / (ref.max() - ref.min())
    return f"mean {d.mean().item():.5f} p99.9 {d.kthvalue(int(0.999 * d.numel())).values.item():.5f} max {d.max().item():.5f}"

for ws in (8, 7):
    kw = {} if ws == 8 else dict(style_encoder_window_size=[7, 7], style_decoder_window_size=[7, 7])
    m = MasterStyleTransferModel(**kw)
    synthetic.fill_state_dict_(m, 0)
    sd = {n: v.detach().cpu().clone() for n, v in m.state_dict().items()}
    m = m.eval().cuda()
    for size, k, seed in ((128, 1, 0), (128, 2, 0), (128, 3, 0), (256, 1, 0), (256, 2, 1), (256, 3, 1), (512, 1, 2), (512, 2, 2), (128, 1, 3), (128, 1, 4)):
        content, style = synthetic.synthetic_images(2, size, seed=seed)
        with torch.no_grad():
            out = m(content.cuda(), style.cuda(), k)
            ref = O.full_forward(sd, content, style, k, ws=ws, sh=4)
        print(f"ws {ws} size {size} k {k} seed {seed}: {stats(out, ref)}", flush=True)
