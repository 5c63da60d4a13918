"""Per-kernel-family time of one eager training step (CUDA events around every launch of this library; torch's own kernels --
fills, copies -- show up as the difference to the step total)."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import MasterStyleTransferModel, custom_loss, ops, synthetic
from mastermetastyletransfer_b200.training import InnerLoopTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
model = MasterStyleTransferModel(); synthetic.fill_state_dict_(model, 0); model = model.to(dev)
loss_fn = custom_loss("/nonexistent"); synthetic.fill_state_dict_(loss_fn, 1); loss_fn = loss_fn.to(dev)
for m in (model.style_transformer.encoder, model.style_transformer.decoder):
    m.stochastic_depth.p = 0.0
model.style_transformer.encoder.encoder_stochastic_depth_prob = 0.0
content, style = synthetic.synthetic_images(B, 256, seed=0)
style = style[:1].repeat(B, 1, 1, 1)
content, style = content.to(dev), style.to(dev)
tr = InnerLoopTrainer(model, loss_fn, inner_lr=1e-4)
for _ in range(3):
    tr.step(content, style, 1)
torch.cuda.synchronize()
with ops.timing() as rec:
    tr.step(content, style, 1)
torch.cuda.synchronize()
fam = collections.OrderedDict()
detail = collections.OrderedDict()
for name, flops, nbytes, a, b, desc in rec:
    ms = a.elapsed_time(b)
    f = fam.setdefault(name, [0, 0.0, 0.0])
    f[0] += 1; f[1] += ms; f[2] += flops
    d = detail.setdefault((name, desc), [0, 0.0])
    d[0] += 1; d[1] += ms
tot = sum(v[1] for v in fam.values())
print(f"sum of library kernels {tot:.3f} ms over {sum(v[0] for v in fam.values())} launches")
for k, (n, ms, fl) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} {n:4d} {ms:8.3f} ms {100*ms/tot:5.1f}%  {fl/ms/1e9 if ms else 0:8.1f} TF/s")
print("--- heaviest launches")
for (k, desc), (n, ms) in sorted(detail.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:24s} x{n:2d} {ms:7.3f} ms  {desc}")
