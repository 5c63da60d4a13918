"""Host side of the training workloads (SURVEY.md sections 8a row a19, 8e): the reference's inner-loop step
(train_only_inner_loop.py:523-575 / train.py:436-517) and its Reptile outer update (train.py:524-534), one process per
GPU.  Everything numerical happens in the sm_100a kernels (forward, backward, Adam, delta / apply); this file only
sequences them and owns the two collectives the path has:

  * data-parallel training (BASELINE configs[2]): ONE all-reduce of the flat fp32 gradient (4.30 M floats) per step;
  * meta training (configs[3]): one style task per rank, no communication inside the inner loop, ONE all-reduce of the
    flat parameter delta per outer iteration (optim.reptile_update).
"""
from __future__ import annotations

import copy
import random
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .optim import FusedAdam, reptile_update


def _dist_on(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


@torch.no_grad()
def allreduce_gradients(params: Sequence[torch.nn.Parameter], group=None, flat: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Mean of the gradients over ranks with ONE collective: gradients are packed into a flat fp32 buffer, all-reduced
    (NCCL over NVLink on GPUs, gloo in the CPU tests), and every p.grad is re-pointed at its slice of the buffer so the
    fused Adam kernel reads the averaged values in place (no unpack copy).  Returns the flat buffer."""
    params = [p for p in params if p.requires_grad]
    for p in params:
        if p.grad is None:
            raise RuntimeError("allreduce_gradients: a trainable parameter has no gradient")
    total = sum(p.numel() for p in params)
    if flat is None or flat.numel() != total or flat.device != params[0].device:
        flat = torch.empty(total, dtype=torch.float32, device=params[0].device)
    torch.cat([p.grad.reshape(-1) for p in params], out=flat)
    if _dist_on(group):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / dist.get_world_size(group))
    off = 0
    for p in params:
        n = p.numel()
        p.grad = flat[off:off + n].view_as(p)
        off += n
    return flat


class InnerLoopTrainer:
    """The reference's training step on omega copies of the style transformer and the decoder, with the Swin encoder
    frozen (train.py:216-218,306-309,436-517).  `model` is a MasterStyleTransferModel, `loss_fn` a custom_loss."""

    def __init__(self, model, loss_fn, inner_lr: float = 1e-4, max_layers: int = 4, data_parallel: bool = False, group=None,
                 seed: int = 0):
        self.model, self.loss_fn = model, loss_fn
        for p in model.swin_encoder.parameters():
            p.requires_grad = False
        self.omega_st = copy.deepcopy(model.style_transformer).train()
        self.omega_dec = copy.deepcopy(model.decoder).train()
        self.params: List[torch.nn.Parameter] = list(self.omega_st.parameters()) + list(self.omega_dec.parameters())
        self.opt = FusedAdam(self.params, lr=inner_lr)
        self.max_layers, self.data_parallel, self.group = max_layers, data_parallel, group
        self._rng = random.Random(seed)  # shared seed: every rank samples the same layer count
        self._flat = None

    def load_from_theta(self) -> None:
        """omega <- theta (train.py:428-431)."""
        self.omega_st.load_state_dict(self.model.style_transformer.state_dict())
        self.omega_dec.load_state_dict(self.model.decoder.state_dict())

    def step(self, content: torch.Tensor, style: torch.Tensor, num_layers: Optional[int] = None):
        """One inner-loop update; returns the device tensor (total, content, style) without synchronising."""
        if num_layers is None:
            num_layers = self._rng.randint(1, self.max_layers)
        with torch.no_grad():
            fc = self.model.swin_encoder(content)
            fs = self.model.swin_encoder(style)
        out = self.omega_dec(self.omega_st(fc, fs, num_layers).permute(0, 3, 1, 2))
        total, closs, sloss = self.loss_fn(content, style, out, output_content_and_style_loss=True)
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        if self.data_parallel:
            self._flat = allreduce_gradients(self.params, self.group, self._flat)
        self.opt.step()
        return torch.stack([total.detach(), closs.detach(), sloss.detach()])

    def outer_update(self, outer_lr: float) -> None:
        """theta += outer_lr * mean_over_ranks(omega - theta) for the style transformer and the decoder (train.py:524-534)."""
        reptile_update(self.model.style_transformer, self.omega_st, outer_lr, self.group)
        reptile_update(self.model.decoder, self.omega_dec, outer_lr, self.group)


def meta_iteration(trainer: InnerLoopTrainer, style: torch.Tensor, content_batches: Iterable[torch.Tensor], outer_lr: float,
                   num_layers: Optional[int] = None):
    """One outer iteration of train.py:400-534 for this rank's style task: omega <- theta, one inner step per content
    batch, then the (all-reduced) Reptile update.  Returns the last inner loss tensor."""
    trainer.load_from_theta()
    last = None
    for content in content_batches:
        last = trainer.step(content, style, num_layers)
    trainer.outer_update(outer_lr)
    return last
