"""Driver for profiling the training step (run under ncu / plain): warm-up steps, then ONE step whose launch count is printed."""
import os, sys
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
for _ in range(2):
    tr.step(content, style, 1)
torch.cuda.synchronize()
n0 = ops.launch_count
import time
t0 = time.perf_counter()
tr.step(content, style, 1)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"LAUNCHES_LAST_STEP {ops.launch_count - n0} host_enqueue_ms {t_host*1e3:.2f} wall_ms {t_all*1e3:.2f}")
