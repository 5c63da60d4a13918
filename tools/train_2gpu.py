"""Data-parallel training step across ranks (BASELINE configs[2]): every rank runs the reference's inner-loop
step on its own content batch; one NCCL all-reduce of the flat gradient; identical Adam update on every rank.
Checks: parameters stay bit-identical across ranks, and the averaged gradient equals the mean of the per-rank gradients.
Launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_2gpu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mastermetastyletransfer_b200 import MasterStyleTransferModel, custom_loss, synthetic
from mastermetastyletransfer_b200.training import InnerLoopTrainer, meta_iteration

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
model = MasterStyleTransferModel()
synthetic.fill_state_dict_(model, 0)
model = model.to(dev)
loss_fn = custom_loss("/nonexistent")
synthetic.fill_state_dict_(loss_fn, 1)
loss_fn = loss_fn.to(dev)
for m in (model.style_transformer.encoder, model.style_transformer.decoder):
    m.stochastic_depth.p = 0.0
model.style_transformer.encoder.encoder_stochastic_depth_prob = 0.0
B, S = 2, 128
content, style = synthetic.synthetic_images(B, S, seed=10 + rank)
content, style = content.to(dev), style.to(dev)

# 1) the all-reduced gradient is the mean of the per-rank gradients
tr = InnerLoopTrainer(model, loss_fn, inner_lr=1e-4, data_parallel=False)
with torch.no_grad():
    fc, fs = model.swin_encoder(content), model.swin_encoder(style)
out = tr.omega_dec(tr.omega_st(fc, fs, 1).permute(0, 3, 1, 2))
loss_fn(content, style, out).backward()
local = torch.cat([p.grad.reshape(-1) for p in tr.params]).clone()
gathered = [torch.empty_like(local) for _ in range(world)]
dist.all_gather(gathered, local)
expect = torch.stack(gathered).mean(0)
from mastermetastyletransfer_b200.training import allreduce_gradients
flat = allreduce_gradients(tr.params)
err = ((flat - expect).abs().max() / expect.abs().max()).item()
views_ok = all(p.grad.data_ptr() >= flat.data_ptr() and p.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4 for p in tr.params)

# 2) three DP steps keep the replicas bit-identical; 3) a meta iteration (one style task per rank) too
tr = InnerLoopTrainer(model, loss_fn, inner_lr=1e-4, data_parallel=True)
for _ in range(3):
    losses = tr.step(content, style, 1)
def same_everywhere(params):
    v = torch.cat([p.detach().reshape(-1) for p in params])
    ref = v.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([1.0 if torch.equal(v, ref) else 0.0], device=v.device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)  # rank 0 always equals itself: the verdict is the minimum over ranks
    return bool(same.item())
dp_same = same_everywhere(tr.params)
tm = InnerLoopTrainer(model, loss_fn, inner_lr=1e-4, data_parallel=False)
theta0 = torch.cat([p.detach().reshape(-1) for p in model.style_transformer.parameters()]).clone()
meta_iteration(tm, style, [content], 0.5, 1)
theta1 = torch.cat([p.detach().reshape(-1) for p in model.style_transformer.parameters()])
meta_same = same_everywhere(list(model.style_transformer.parameters()) + list(model.decoder.parameters()))
moved = (theta1 - theta0).abs().max().item() > 0
omega_differs = not same_everywhere(tm.params) if world > 1 else True  # each rank trained on its own task
dist.barrier()
if rank == 0:
    print(f"TRAIN_2GPU world={world} grad_mean_err={err:.3e} views_ok={views_ok} dp_identical={dp_same} meta_identical={meta_same} "
          f"theta_moved={moved} omega_differs={omega_differs} loss={[round(x, 4) for x in losses.tolist()]}")
assert err < 1e-5 and views_ok and dp_same and meta_same and moved and omega_differs
dist.destroy_process_group()
