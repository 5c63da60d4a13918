"""Training-step glue on the GPU: fused multi-tensor Adam and the Reptile-style
outer update of train.py:524-534, generalised to one style task per GPU with an NCCL all-reduce of the
parameter deltas.  Parameters stay ordinary nn.Parameters; the kernels get a device
table of their pointers."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import MstTensorTable


def _bump_versions(params) -> None:
    """The kernels write parameters through raw pointers; bump the tensors' version counters so that everything keyed
    on (data_ptr, _version) -- the packed-weight caches of the forward path -- sees the update."""
    params = list(params)
    try:
        torch._C._autograd._unsafe_set_version_counter(tuple(params), tuple(p._version + 1 for p in params))
    except Exception:  # older/newer torch without the helper: an in-place no-op bumps the counter too
        with torch.no_grad():
            for p in params:
                p.add_(0.0)


class _Table:
    """Device-side description of a list of fp32 tensors (rebuilt only when a data pointer changes)."""

    def __init__(self, lists: List[List[torch.Tensor]]):
        n = len(lists[0])
        dev = lists[0][0].device
        chunk = _lib.lib().mst_opt_chunk_elems()
        starts, numel, offs, c, off = [], [], [], 0, 0
        for t in lists[0]:
            starts.append(c)
            numel.append(t.numel())
            offs.append(off)
            c += -(-t.numel() // chunk)
            off += (t.numel() + 3) // 4 * 4
        self.total, self.n_chunks, self.n = off, c, n
        self.key = tuple(t.data_ptr() for l in lists for t in l)
        self._host = []  # pinned staging: the uploads are async copies, legal inside a CUDA-graph capture and replayable

        def upload(values, dtype):
            h = torch.tensor(values, dtype=dtype).pin_memory()
            self._host.append(h)
            return h.to(dev, non_blocking=True)

        self._keep = [upload(starts, torch.int32), upload(numel, torch.int64), upload(offs, torch.int64)]
        for l in lists:
            for t in l:
                if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                    raise TypeError("optimiser kernels need contiguous fp32 CUDA tensors")
            self._keep.append(upload([t.data_ptr() for t in l], torch.int64))
        tb = MstTensorTable()
        tb.chunk_start, tb.numel, tb.flat_offset = (k.data_ptr() for k in self._keep[:3])
        ptrs = [k.data_ptr() for k in self._keep[3:]] + [None] * 4
        tb.a, tb.b, tb.c, tb.d = ptrs[:4]
        tb.n_tensors, tb.n_chunks = n, c
        self.tb = tb


class FusedAdam:
    """torch.optim.Adam semantics (lr, betas, eps, weight_decay; no amsgrad) with one kernel launch per parameter group.
    Built the way the reference builds its optimisers: from a parameter iterable (train.py:398) or from a list of
    {'params': ...} groups (train_only_inner_loop.py:468-478); `param_groups[i]['lr']` can be rescheduled between steps
    (train_only_inner_loop.py:321-340)."""

    def __init__(self, params: Iterable, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, capturable: bool = False):
        """capturable=True keeps {lr, step} on the device so that step() can be captured in a CUDA graph and replayed
        (training.GraphedTrainStep); change the learning rate with set_lr() between replays."""
        params = list(params)
        groups = params if params and isinstance(params[0], dict) else [{"params": params}]
        self.param_groups = []
        for g in groups:
            ps = [p for p in g["params"] if p.requires_grad]
            self.param_groups.append({"params": ps, "lr": g.get("lr", lr), "betas": g.get("betas", betas), "eps": g.get("eps", eps),
                                      "weight_decay": g.get("weight_decay", weight_decay)})
        self.params = [p for g in self.param_groups for p in g["params"]]
        self.state = [{"exp_avg": [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in g["params"]],
                       "exp_avg_sq": [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in g["params"]],
                       "table": None, "tables": {}} for g in self.param_groups]
        # Device tables built while a CUDA graph was being captured: their pinned host staging is read by the captured H2D
        # copy nodes on EVERY replay, so they must outlive every later table change (an eager step with other gradient
        # pointers, a second GraphedTrainStep on the same trainer).  Never dropped.
        self._captured_tables = []
        self.step_count = 0
        self.capturable = capturable
        self._dev_state = None
        if capturable:
            dev = self.params[0].device
            self._dev_state = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in self.param_groups]  # {float lr, int step}
            for g, st in zip(self.param_groups, self._dev_state):
                st.view(torch.float32)[0] = float(g["lr"])
            self._dev_lr = [float(g["lr"]) for g in self.param_groups]  # host mirror of the device-side learning rates

    def set_lr(self, lr: float, group: Optional[int] = None) -> None:
        """Learning-rate change that a captured graph sees (device-side state); eager mode just updates param_groups."""
        for i, g in enumerate(self.param_groups):
            if group is None or group == i:
                g["lr"] = lr
                if self._dev_state is not None:
                    self._dev_state[i].view(torch.float32)[0] = float(lr)
                    self._dev_lr[i] = float(lr)

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value):
        self.set_lr(value)

    def sync_lr(self) -> None:
        """capturable mode: push `param_groups[i]['lr']` edits (the reference's rescheduling idiom,
        train_only_inner_loop.py:321-340) to the device-side state that the kernels and captured graphs read.  Called by every
        eager step() and by GraphedTrainStep.step() before a replay."""
        if self._dev_state is None or torch.cuda.is_current_stream_capturing():
            return
        for i, g in enumerate(self.param_groups):
            if float(g["lr"]) != self._dev_lr[i]:
                self._dev_state[i].view(torch.float32)[0] = float(g["lr"])
                self._dev_lr[i] = float(g["lr"])

    def zero_grad(self, set_to_none: bool = False):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    # ---- checkpoint / resume: torch.optim.Adam's state_dict layout, so a run can move between the two ----
    def _steps_taken(self) -> int:
        if self._dev_state is not None:  # graph replays advance the device-side counter only
            return max(self.step_count, max(int(st[1].item()) for st in self._dev_state))
        return self.step_count

    def state_dict(self) -> dict:
        """{'state': {index: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]} exactly as torch.optim.Adam packs it
        (parameters numbered in group order, moments cloned)."""
        step = self._steps_taken()
        state, groups, i = {}, [], 0
        for g, st in zip(self.param_groups, self.state):
            idx = []
            for m, v in zip(st["exp_avg"], st["exp_avg_sq"]):
                if step > 0:  # torch creates a parameter's state at its first step
                    state[i] = {"step": torch.tensor(float(step)), "exp_avg": m.detach().clone(), "exp_avg_sq": v.detach().clone()}
                idx.append(i)
                i += 1
            groups.append({"lr": g["lr"], "betas": tuple(g["betas"]), "eps": g["eps"], "weight_decay": g["weight_decay"],
                           "amsgrad": False, "maximize": False, "foreach": None, "capturable": self.capturable,
                           "differentiable": False, "fused": None, "decoupled_weight_decay": False, "params": idx})
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, sd: dict) -> None:
        """Accepts FusedAdam.state_dict() and torch.optim.Adam.state_dict() of an optimiser built over the same parameter
        groups.  The kernel keeps ONE step count for all tensors, so every stored 'step' must be the same."""
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"]) for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has different parameter groups")
        for a in groups:
            if a.get("amsgrad") or a.get("maximize") or a.get("decoupled_weight_decay"):
                raise ValueError("FusedAdam implements plain Adam only (no amsgrad / maximize / decoupled weight decay)")
        steps = {int(float(e["step"])) for e in sd["state"].values()}
        if len(steps) > 1:
            raise ValueError("FusedAdam keeps one step count for all parameters; the loaded state has several")
        n = sum(len(g["params"]) for g in self.param_groups)
        if sd["state"] and len(sd["state"]) != n:
            raise ValueError("loaded state dict covers only some of the parameters")
        for g, a, st in zip(self.param_groups, groups, self.state):
            g["lr"], g["betas"], g["eps"], g["weight_decay"] = a["lr"], tuple(a["betas"]), a["eps"], a["weight_decay"]
            for j, i in enumerate(a["params"]):
                e = sd["state"].get(i)
                if e is None:
                    st["exp_avg"][j].zero_()
                    st["exp_avg_sq"][j].zero_()
                    continue
                if e["exp_avg"].shape != st["exp_avg"][j].shape:
                    raise ValueError(f"moment {i} has shape {tuple(e['exp_avg'].shape)}, expected {tuple(st['exp_avg'][j].shape)}")
                st["exp_avg"][j].copy_(e["exp_avg"])  # in place: the device tables (and captured graphs) keep their pointers
                st["exp_avg_sq"][j].copy_(e["exp_avg_sq"])
        self.step_count = steps.pop() if steps else 0
        if self._dev_state is not None:
            for i, (g, ds) in enumerate(zip(self.param_groups, self._dev_state)):
                ds.view(torch.float32)[0] = float(g["lr"])
                ds[1] = self.step_count
                self._dev_lr[i] = float(g["lr"])

    @torch.no_grad()
    def step(self):
        self.step_count += 1
        self.sync_lr()
        for gi, (g, st) in enumerate(zip(self.param_groups, self.state)):
            if not g["params"]:
                continue
            grads = []
            for p in g["params"]:
                if p.grad is None:
                    raise RuntimeError("FusedAdam.step: every parameter needs a gradient (the kernel updates all tensors in one launch)")
                grads.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
            lists = [[p.data for p in g["params"]], grads, st["exp_avg"], st["exp_avg_sq"]]
            key = tuple(t.data_ptr() for l in lists for t in l)
            if st["table"] is None or st["table"].key != key:
                tbl = st["tables"].get(key)
                if tbl is None:
                    tbl = _Table(lists)
                    if torch.cuda.is_current_stream_capturing():
                        self._captured_tables.append(tbl)  # immortal: a graph replays uploads from its pinned staging
                    if len(st["tables"]) >= 8:  # eager training allocates fresh gradients every step: bound the cache
                        st["tables"].clear()
                    st["tables"][key] = tbl
                st["table"] = tbl
            n = sum(p.numel() for p in g["params"])
            tb = st["table"].tb
            if self.capturable:
                ds = self._dev_state[gi]  # (not list.index: comparing group dicts would compare parameter tensors)
                ops._launch("mst_adam_step", lambda: _lib.lib().mst_adam_step_dev(C.byref(tb), float(g["betas"][0]), float(g["betas"][1]),
                                                                                 float(g["eps"]), float(g["weight_decay"]), ds.data_ptr(), 1,
                                                                                 ops._stream()), nbytes=28.0 * n)
            else:
                ops._launch("mst_adam_step", lambda: _lib.lib().mst_adam_step(C.byref(tb), float(g["lr"]), float(g["betas"][0]),
                                                                             float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                                                                             int(self.step_count), ops._stream()), nbytes=28.0 * n)
        _bump_versions(self.params)


@torch.no_grad()
def reptile_update(theta, omega, outer_lr: float, group=None) -> None:
    """theta += outer_lr * mean_over_ranks(omega - theta)  (train.py:524-534 when there is a single rank).

    With torch.distributed initialised, each rank holds an omega trained on its own style task; the flat delta
    buffer is all-reduced (NCCL on GPUs) and every rank applies the same averaged update."""
    as_list = lambda m: [p for _, p in m.named_parameters()] if isinstance(m, torch.nn.Module) else list(m)
    tp, op = as_list(theta), as_list(omega)  # modules, or parameter lists (several modules in ONE collective)
    if len(tp) != len(op) or any(a.shape != b.shape for a, b in zip(tp, op)):
        raise ValueError("theta and omega must have identical parameter lists")
    table = _Table([[p.data for p in tp], [p.data for p in op]])
    flat = torch.empty(table.total, dtype=torch.float32, device=tp[0].device)
    n = sum(p.numel() for p in tp)
    ops._launch("mst_reptile_delta", lambda: _lib.lib().mst_reptile_delta(C.byref(table.tb), flat.data_ptr(), ops._stream()), nbytes=12.0 * n)
    world = 1
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    ops._launch("mst_reptile_apply", lambda: _lib.lib().mst_reptile_apply(C.byref(table.tb), flat.data_ptr(), float(outer_lr) / world,
                                                                         ops._stream()), nbytes=12.0 * n)
    _bump_versions(tp)
