import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mastermetastyletransfer_b200 import MasterStyleTransferModel, synthetic
from oracle import master_oracle as O
m = MasterStyleTransferModel(); synthetic.fill_state_dict_(m, 0); m = m.cuda().eval()
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
g = torch.Generator().manual_seed(5)
x = torch.randn(2, 16, 16, 256, generator=g)
G = torch.randn(2, 3, 128, 128, generator=g)
ps = {k[len("decoder."):]: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("decoder.")}
xr = x.clone().requires_grad_(True)
ref = O.cnn_decoder(ps, xr.permute(0, 3, 1, 2), "decoder.")
(ref * G).sum().backward()
dec = m.decoder
xc = x.cuda().requires_grad_(True)
out = dec(xc.permute(0, 3, 1, 2))
(out * G.cuda()).sum().backward()
def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm()).item()
for n, p in dec.named_parameters():
    print(n, "rel", rel(p.grad, ps[n].grad))
e = (xc.grad.cpu() - xr.grad).abs()  # [2,16,16,256]
print("dx rel", rel(xc.grad, xr.grad))
print("err by row", e.mean(dim=(0, 2, 3)))
print("err by col", e.mean(dim=(0, 1, 3)))
print("ref by row", xr.grad.abs().mean(dim=(0, 2, 3)))
