"""Meta (outer-loop) update across ranks: each rank holds its own omega (one style task per GPU);
reptile_update all-reduces the deltas over NCCL and every rank ends with the same theta.
Launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/reptile_2gpu.py"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mastermetastyletransfer_b200 import StyleTransformer, synthetic
from mastermetastyletransfer_b200.optim import reptile_update

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
theta = StyleTransformer(256, 256, 8, 8, [8, 8], [8, 8], [4, 4], [4, 4])
synthetic.fill_state_dict_(theta, 0)
theta = theta.cuda()
omegas = []
for r in range(world):  # every rank can rebuild every rank's omega (seeded), to form the expected mean locally
    o = copy.deepcopy(theta)
    g = torch.Generator(device="cuda").manual_seed(100 + r)
    with torch.no_grad():
        for p in o.parameters():
            p.add_(torch.randn(p.shape, generator=g, device="cuda") * 0.02)
    omegas.append(o)
expect = copy.deepcopy(theta)
with torch.no_grad():
    for i, p in enumerate(expect.parameters()):
        mean_delta = sum(list(o.parameters())[i] - p for o in omegas) / world
        p.add_(0.5 * mean_delta)
reptile_update(theta, omegas[rank], 0.5)
err = max((a - b).abs().max().item() for a, b in zip(theta.parameters(), expect.parameters()))
flat = torch.cat([p.reshape(-1) for p in theta.parameters()])
ref = flat.clone()
dist.broadcast(ref, 0)
same = torch.equal(flat, ref)
dist.barrier()
if rank == 0:
    print(f"REPTILE_2GPU world={world} max_err={err:.3e} identical_across_ranks={same}")
assert err < 1e-6 and same
dist.destroy_process_group()
