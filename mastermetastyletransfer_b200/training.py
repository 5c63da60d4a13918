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

from .optim import FusedAdam, _bump_versions, reptile_update


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
    lo, hi = flat.data_ptr(), flat.data_ptr() + 4 * total
    if any(lo <= p.grad.data_ptr() < hi for p in params):  # gradients accumulated into last step's views: pack out of place
        flat.copy_(torch.cat([p.grad.reshape(-1) for p in params]))
    else:
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
                 seed: int = 0, capturable: bool = False, fast_adaptation: bool = False):
        """fast_adaptation=True is the reference's few-shot stage (train_only_inner_loop.py:306-318): everything frozen except
        the style ENCODER of the style transformer; the adjoint kernels still run the whole backward (the encoder's gradients
        need the chain through the CNN decoder and the style decoder), only the encoder's 2.37 M parameters are updated."""
        self.model, self.loss_fn = model, loss_fn
        for p in model.swin_encoder.parameters():
            p.requires_grad = False
        self.fast_adaptation = fast_adaptation
        if fast_adaptation:
            for p in model.style_transformer.decoder.parameters():
                p.requires_grad = False
            for p in model.style_transformer.encoder.parameters():
                p.requires_grad = True
            for p in model.decoder.parameters():
                p.requires_grad = False
        if _dist_on(group):
            # every rank must start from the same theta (the all-reduced gradient / Reptile delta is applied to each rank's own
            # copy): one broadcast of the flat parameter vector from rank 0 instead of trusting seeds and checkpoints to agree
            with torch.no_grad():
                theta = list(model.style_transformer.parameters()) + list(model.decoder.parameters())
                flat = torch.cat([p.data.reshape(-1) for p in theta])
                dist.broadcast(flat, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
                off = 0
                for p in theta:
                    p.data.copy_(flat[off:off + p.numel()].view_as(p))
                    off += p.numel()
            _bump_versions(theta)
        self.omega_st = copy.deepcopy(model.style_transformer).train()
        self.omega_dec = copy.deepcopy(model.decoder).train()
        self.params: List[torch.nn.Parameter] = list(self.omega_st.parameters()) + list(self.omega_dec.parameters())
        self._theta: List[torch.nn.Parameter] = list(model.style_transformer.parameters()) + list(model.decoder.parameters())
        self.opt = FusedAdam(self.params, lr=inner_lr, capturable=capturable)
        self.max_layers, self.data_parallel, self.group = max_layers, data_parallel, group
        self._rng = random.Random(seed)  # shared seed: every rank samples the same layer count
        self._flat = None

    def load_from_theta(self) -> None:
        """omega <- theta (train.py:428-431: load_state_dict of the two trained modules; their buffers -- the relative-position
        index maps -- are constants).  One multi-tensor copy into the existing omega storage, so a captured training step
        (GraphedTrainStep) keeps reading the right memory."""
        with torch.no_grad():
            torch._foreach_copy_([p.data for p in self.params], [p.data for p in self._theta])
        _bump_versions(self.params)  # `.data` has its own version counter: the packed-weight caches key on the parameters'

    def step(self, content: torch.Tensor, style: torch.Tensor, num_layers: Optional[int] = None):
        """One inner-loop update; returns the device tensor (total, content, style) without synchronising."""
        if num_layers is None:
            num_layers = self._rng.randint(1, self.max_layers)
        with torch.no_grad():
            branch = self._fork_weight_packing()
            loss_branch = self._fork_content_style_taps(content, style)
            fc = self.model.swin_encoder(content)
            fs = self.model.swin_encoder(style)
            if branch is not None:
                torch.cuda.current_stream(content.device).wait_stream(branch)
        out = self.omega_dec(self.omega_st(fc, fs, num_layers).permute(0, 3, 1, 2))
        if loss_branch is not None:
            torch.cuda.current_stream(content.device).wait_stream(loss_branch)
        total, closs, sloss = self.loss_fn(content, style, out, output_content_and_style_loss=True)
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        if self.data_parallel:
            self._flat = allreduce_gradients(self.params, self.group, self._flat)
        self.opt.step()
        return torch.stack([total.detach(), closs.detach(), sloss.detach()])

    def _fork_weight_packing(self):
        """While a step is being captured into a CUDA graph: re-pack omega's updated weights (58 small launches: bf16 tile images
        of every Linear / conv, forward and transposed) on a branch stream, i.e. as graph nodes parallel to the frozen encoder's
        forward, which does not depend on them.  The modules' own forward then finds the packed weights in the cache.
        Eager steps keep everything on one stream (returns None)."""
        if not (self.params and self.params[0].is_cuda and torch.cuda.is_current_stream_capturing()):
            return None
        from . import train_engine as te
        from .style_transformer import packed_weights
        dev = self.params[0].device
        if getattr(self, "_pack_stream", None) is None:
            self._pack_stream = torch.cuda.Stream(device=dev)
        self._pack_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._pack_stream):
            packed_weights(self.omega_st, te.StyleTransformerTrainWeights)
            packed_weights(self.omega_dec, te.CnnDecoderTrainWeights)
        return self._pack_stream

    def _fork_content_style_taps(self, content, style):
        """While a step is being captured: the VGG passes of the content and style images (no gradient, independent of the
        model) as a graph branch parallel to the encoder / style transformer / decoder forward, whose batch-8 kernels leave
        most SMs idle.  Returns the branch stream (joined before the loss) or None (eager: one stream)."""
        if not (content.is_cuda and torch.cuda.is_current_stream_capturing()):
            return None
        from .autograd_fns import prefetch_content_style_taps
        dev = content.device
        if getattr(self, "_loss_stream", None) is None:
            self._loss_stream = torch.cuda.Stream(device=dev)
        self._loss_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._loss_stream):
            prefetch_content_style_taps(self.loss_fn, content, style)
        return self._loss_stream

    # ---- checkpoint / resume (the reference saves only the two trained modules, train.py:288-299) ----
    def state_dict(self) -> dict:
        """Everything a resumed run needs: theta (the two trained modules, under the names the reference saves them with),
        omega, the inner optimiser (torch.optim.Adam layout) and the layer-count sampler."""
        clone = lambda sd: {k: v.detach().clone() for k, v in sd.items()}
        return {"style_transformer": clone(self.model.style_transformer.state_dict()), "decoder": clone(self.model.decoder.state_dict()),
                "omega_style_transformer": clone(self.omega_st.state_dict()), "omega_decoder": clone(self.omega_dec.state_dict()),
                "optimizer": self.opt.state_dict(), "layer_sampler": self._rng.getstate()}

    def load_state_dict(self, sd: dict) -> None:
        """In place (parameter storage, optimiser moments and a captured GraphedTrainStep stay valid)."""
        self.model.style_transformer.load_state_dict(sd["style_transformer"])
        self.model.decoder.load_state_dict(sd["decoder"])
        self.omega_st.load_state_dict(sd["omega_style_transformer"])
        self.omega_dec.load_state_dict(sd["omega_decoder"])
        self.opt.load_state_dict(sd["optimizer"])
        state = sd["layer_sampler"]
        self._rng.setstate((state[0], tuple(state[1]), state[2]))  # a torch.save / json round trip turns the tuple into a list

    def outer_update(self, outer_lr: float) -> None:
        """theta += outer_lr * mean_over_ranks(omega - theta) for the style transformer and the decoder (train.py:524-534)."""
        reptile_update(self._theta, self.params, outer_lr, self.group)  # both modules: one delta buffer, ONE all-reduce


def meta_iteration(trainer: InnerLoopTrainer, style: torch.Tensor, content_batches: Iterable[torch.Tensor], outer_lr: float,
                   num_layers: Optional[int] = None, graphed: Optional["GraphedTrainStep"] = None):
    """One outer iteration of train.py:400-534 for this rank's style task: omega <- theta, one inner step per content
    batch, then the (all-reduced) Reptile update.  Returns the last inner loss tensor.  graphed: a GraphedTrainStep built
    on `trainer` -- every inner step is then one CUDA-graph replay (its layer count is the graph's)."""
    if graphed is not None and graphed.trainer is not trainer:
        raise ValueError("meta_iteration: `graphed` must wrap the same InnerLoopTrainer")
    trainer.load_from_theta()
    last = None
    for content in content_batches:
        last = graphed.step(content, style) if graphed is not None else trainer.step(content, style, num_layers)
    trainer.outer_update(outer_lr)
    return last


class GraphedTrainStep:
    """One inner-loop step (forward, loss, backward, optional gradient all-reduce, Adam) captured in a CUDA graph: the ~360
    kernel launches of a step cost ~15 ms of host time when issued one by one through ctypes, more than the GPU work itself;
    a replay costs one launch.  The layer count is a host-side choice (train.py:449), so there is one graph per value.
    `trainer` must have been built with capturable=True (device-side Adam step / learning rate)."""

    def __init__(self, trainer: InnerLoopTrainer, batch: int, size: int, num_layers: int = 1, warmup: int = 3):
        if not trainer.opt.capturable:
            raise ValueError("GraphedTrainStep needs InnerLoopTrainer(..., capturable=True)")
        self.trainer, self.num_layers = trainer, num_layers
        dev = trainer.params[0].device
        self.content = torch.zeros(batch, 3, size, size, device=dev)
        self.style = torch.zeros(batch, 3, size, size, device=dev)
        self._rand_fill()
        # the warm-up steps really train (on noise): snapshot parameters and optimiser state, restore after the capture
        opt = trainer.opt
        snap = [[p.detach().clone() for p in trainer.params], [[t.clone() for t in st["exp_avg"]] for st in opt.state],
                [[t.clone() for t in st["exp_avg_sq"]] for st in opt.state], [d.clone() for d in opt._dev_state], opt.step_count]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):  # sizes every workspace, packs weights, sets kernel attributes
                trainer.step(self.content, self.style, num_layers)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        opt.zero_grad(set_to_none=True)
        from . import ops
        n0 = ops.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = trainer.step(self.content, self.style, num_layers)
        self.launches = ops.launch_count - n0  # kernels of this library inside one replay
        torch.cuda.synchronize(dev)
        from .style_transformer import pin_state
        self._pinned = pin_state([trainer.model, trainer.omega_st, trainer.omega_dec, trainer.loss_fn])  # baked-in addresses stay valid
        with torch.no_grad():
            for p, v in zip(trainer.params, snap[0]):
                p.copy_(v)
            for st, ea, es in zip(opt.state, snap[1], snap[2]):
                for t, v in zip(st["exp_avg"], ea):
                    t.copy_(v)
                for t, v in zip(st["exp_avg_sq"], es):
                    t.copy_(v)
            for d, v in zip(opt._dev_state, snap[3]):
                d.copy_(v)
        opt.step_count = snap[4]

    def _rand_fill(self):
        g = torch.Generator(device=self.content.device).manual_seed(0)
        self.content.copy_(torch.rand(self.content.shape, generator=g, device=self.content.device))
        self.style.copy_(torch.rand(self.style.shape, generator=g, device=self.style.device))

    def step(self, content: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
        self.content.copy_(content, non_blocking=True)
        self.style.copy_(style, non_blocking=True)
        self.trainer.opt.sync_lr()
        self.graph.replay()
        # the replayed Adam kernel wrote omega through raw pointers: anything keyed on (data_ptr, _version) -- the packed
        # weights an eager forward of omega would reuse -- must see a new version (host-side only, no launch)
        _bump_versions(self.trainer.params)
        return self.losses
