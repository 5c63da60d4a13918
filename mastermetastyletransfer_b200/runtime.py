"""CUDA-graph runtime around the drop-in model: one captured graph per (batch, size, layers) so a
stylization step is a single graph launch (no per-kernel Python/ctypes overhead), plus the pinned-host
entry point that bench.py times end to end."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class GraphedStylizer:
    def __init__(self, model: torch.nn.Module, batch: int, size: int, layers: int = 1, device: Optional[torch.device] = None):
        self.model, self.batch, self.size, self.layers = model, batch, size, layers
        self.device = device or next(model.parameters()).device
        self.content = torch.zeros(batch, 3, size, size, device=self.device)
        self.style = torch.zeros(batch, 3, size, size, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.copy_done = torch.cuda.Event()
        with torch.no_grad():
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                for _ in range(2):  # packs weights, sizes the workspace
                    self.model(self.content, self.style, layers)
                n0 = ops.launch_count
                self.model(self.content, self.style, layers)
                self.launches_per_step = ops.launch_count - n0
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.output = self.model(self.content, self.style, layers)
        torch.cuda.synchronize(self.device)

    def load(self, content: torch.Tensor, style: torch.Tensor) -> None:
        self.content.copy_(content, non_blocking=True)
        self.style.copy_(style, non_blocking=True)

    def replay(self) -> torch.Tensor:
        """Enqueue one stylization of the resident inputs on the current stream; returns the static output."""
        self.graph.replay()
        return self.output

    def stylize_many(self, batches):
        """Pipelined host API: `batches` is a sequence of (content_pinned, style_pinned, out_pinned).  Host->device
        copies of batch i+1 and the device->host copy of batch i-1 run on two copy streams while the graph of
        batch i executes (double-buffered device staging, one small device-to-device copy each way).
        Returns after every output has landed in its pinned buffer."""
        dev = self.device
        if not hasattr(self, "_h2d"):
            self._h2d, self._d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            self._stage_in = [(torch.empty_like(self.content), torch.empty_like(self.style)) for _ in range(2)]
            self._stage_out = [torch.empty_like(self.output) for _ in range(2)]
        main = torch.cuda.current_stream(dev)
        in_ready = [torch.cuda.Event() for _ in range(2)]     # staging slot filled by the H2D stream
        in_free = [torch.cuda.Event() for _ in range(2)]      # staging slot consumed by the compute stream
        out_ready = [torch.cuda.Event() for _ in range(2)]    # output slot filled by the compute stream
        out_free = [torch.cuda.Event() for _ in range(2)]     # output slot drained by the D2H stream
        self._h2d.wait_stream(main)
        self._d2h.wait_stream(main)
        n = len(batches)

        def issue_h2d(i):
            slot = i & 1
            with torch.cuda.stream(self._h2d):
                if i >= 2:
                    self._h2d.wait_event(in_free[slot])
                self._stage_in[slot][0].copy_(batches[i][0], non_blocking=True)
                self._stage_in[slot][1].copy_(batches[i][1], non_blocking=True)
                in_ready[slot].record(self._h2d)

        if n:
            issue_h2d(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                issue_h2d(i + 1)
            main.wait_event(in_ready[slot])
            self.content.copy_(self._stage_in[slot][0], non_blocking=True)
            self.style.copy_(self._stage_in[slot][1], non_blocking=True)
            in_free[slot].record(main)
            self.graph.replay()
            if i >= 2:
                main.wait_event(out_free[slot])
            self._stage_out[slot].copy_(self.output, non_blocking=True)
            out_ready[slot].record(main)
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(out_ready[slot])
                batches[i][2].copy_(self._stage_out[slot], non_blocking=True)
                out_free[slot].record(self._d2h)
        main.wait_stream(self._d2h)
        main.synchronize()

    def stylize_host(self, content_pinned: torch.Tensor, style_pinned: torch.Tensor, out_pinned: torch.Tensor) -> torch.Tensor:
        """Pinned host images in, pinned host images out (H2D + graph + D2H on the current stream, then sync)."""
        self.load(content_pinned, style_pinned)
        self.graph.replay()
        out_pinned.copy_(self.output, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_pinned
