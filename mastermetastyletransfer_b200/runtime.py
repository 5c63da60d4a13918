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
        from .style_transformer import pin_state
        self._pinned = pin_state([self.model])  # packed weights / workspace buffers whose addresses the graph baked in

    def _run_u8(self, c8, s8, o8, normalize: bool) -> None:
        """uint8 images -> model -> uint8 images.  ToTensor + Normalize run inside the patch-embedding kernel's loads and clip * 255
        in the last convolution's epilogue when the model offers forward_u8 for this size (bit-identical to converting first /
        afterwards), else as their own launches."""
        if hasattr(self.model, "forward_u8") and not self.model.training and ops.patch_embed_u8_supported(self.size):
            self.model.forward_u8(c8, s8, self.layers, normalize=(ops.IMAGENET_MEAN, ops.IMAGENET_STD) if normalize else None, out_u8=o8)
            return
        else:
            mean = ops.IMAGENET_MEAN if normalize else None
            ops.images_u8_to_nchw(c8, self.content, mean)
            ops.images_u8_to_nchw(s8, self.style, mean)
            out = self.model(self.content, self.style, self.layers)
        ops.images_nchw_to_u8(out, o8)

    def _u8(self, normalize: bool = True):
        """The uint8 boundary of test_model.py around the same model, as a second graph: uint8 [B,S,S,3] images (decoded and
        resized, :39-44) -> ToTensor + ImageNet Normalize (:48,:111-125) -> model -> clip(x*255) -> uint8 [B,S,S,3] (:207).
        A quarter of the host<->device bytes of the fp32 entry points."""
        if getattr(self, "_u8_norm", None) != normalize:
            B, S, dev = self.batch, self.size, self.device
            self.content_u8 = torch.zeros(B, S, S, 3, dtype=torch.uint8, device=dev)
            self.style_u8 = torch.zeros(B, S, S, 3, dtype=torch.uint8, device=dev)
            self.output_u8 = torch.empty(B, S, S, 3, dtype=torch.uint8, device=dev)

            def run():
                self._run_u8(self.content_u8, self.style_u8, self.output_u8, normalize)

            with torch.no_grad():
                self.stream.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(self.stream):
                    run()
                self.stream.synchronize()
                self.graph_u8 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_u8, stream=self.stream):
                    run()
            torch.cuda.synchronize(dev)
            self._u8_norm = normalize
        return (self.content_u8, self.style_u8), self.output_u8, self.graph_u8

    def _u8_slots(self, normalize: bool = True):
        """Two copies of the uint8 graph, each with its own input / output buffers: the pipelined entry point copies host images
        straight into slot i & 1 and reads the result straight out of it -- no device-to-device staging copies between the
        copy streams and the graph (three per step with a single graph)."""
        if getattr(self, "_u8_slots_norm", None) != normalize:
            B, S, dev = self.batch, self.size, self.device
            slots = []
            with torch.no_grad():
                for _ in range(2):
                    c8 = torch.zeros(B, S, S, 3, dtype=torch.uint8, device=dev)
                    s8 = torch.zeros(B, S, S, 3, dtype=torch.uint8, device=dev)
                    o8 = torch.empty(B, S, S, 3, dtype=torch.uint8, device=dev)

                    def run(c8=c8, s8=s8, o8=o8):
                        self._run_u8(c8, s8, o8, normalize)

                    self.stream.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(self.stream):
                        run()
                    self.stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        run()
                    slots.append((c8, s8, o8, g))
            torch.cuda.synchronize(dev)
            self._u8_slot_graphs = slots
            self._u8_slots_norm = normalize
        return self._u8_slot_graphs

    def load(self, content: torch.Tensor, style: torch.Tensor) -> None:
        self.content.copy_(content, non_blocking=True)
        self.style.copy_(style, non_blocking=True)

    def replay(self) -> torch.Tensor:
        """Enqueue one stylization of the resident inputs on the current stream; returns the static output."""
        self.graph.replay()
        return self.output

    def stylize_many(self, batches, u8: bool = False, normalize: bool = True):
        """Pipelined host API: `batches` is a sequence of (content_pinned, style_pinned, out_pinned).  Host->device
        copies of batch i+1 and the device->host copy of batch i-1 run on two copy streams while the graph of
        batch i executes (fp32: double-buffered device staging, one small device-to-device copy each way; uint8: two graphs with
        their own buffers, no staging copies -- _stylize_many_u8).  Returns after every output has landed in its pinned buffer.
        u8=True: the buffers are uint8 [B,S,S,3] images (see _u8); otherwise normalised fp32 [B,3,S,S] tensors."""
        dev = self.device
        if not hasattr(self, "_h2d"):
            self._h2d, self._d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        if u8:
            return self._stylize_many_u8(batches, normalize)
        d_content, d_style, d_output, graph, key = self.content, self.style, self.output, self.graph, "_stage_f32"
        if not hasattr(self, key):
            setattr(self, key, ([(torch.empty_like(d_content), torch.empty_like(d_style)) for _ in range(2)],
                                [torch.empty_like(d_output) for _ in range(2)]))
        stage_in, stage_out = getattr(self, key)
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
                stage_in[slot][0].copy_(batches[i][0], non_blocking=True)
                stage_in[slot][1].copy_(batches[i][1], non_blocking=True)
                in_ready[slot].record(self._h2d)

        if n:
            issue_h2d(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                issue_h2d(i + 1)
            main.wait_event(in_ready[slot])
            d_content.copy_(stage_in[slot][0], non_blocking=True)
            d_style.copy_(stage_in[slot][1], non_blocking=True)
            in_free[slot].record(main)
            graph.replay()
            if i >= 2:
                main.wait_event(out_free[slot])
            stage_out[slot].copy_(d_output, non_blocking=True)
            out_ready[slot].record(main)
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(out_ready[slot])
                batches[i][2].copy_(stage_out[slot], non_blocking=True)
                out_free[slot].record(self._d2h)
        main.wait_stream(self._d2h)
        main.synchronize()

    def _stylize_many_u8(self, batches, normalize: bool) -> None:
        """stylize_many for uint8 images: slot-specific graphs (see _u8_slots), host <-> device copies on the two copy streams."""
        dev = self.device
        slots = self._u8_slots(normalize)
        main = torch.cuda.current_stream(dev)
        in_ready = [torch.cuda.Event() for _ in range(2)]   # slot inputs filled by the H2D stream
        in_free = [torch.cuda.Event() for _ in range(2)]    # slot inputs consumed by the graph
        out_ready = [torch.cuda.Event() for _ in range(2)]  # slot output written by the graph
        out_free = [torch.cuda.Event() for _ in range(2)]   # slot output drained by the D2H stream
        self._h2d.wait_stream(main)
        self._d2h.wait_stream(main)
        n = len(batches)

        def issue_h2d(i):
            slot = i & 1
            with torch.cuda.stream(self._h2d):
                if i >= 2:
                    self._h2d.wait_event(in_free[slot])
                slots[slot][0].copy_(batches[i][0], non_blocking=True)
                slots[slot][1].copy_(batches[i][1], non_blocking=True)
                in_ready[slot].record(self._h2d)

        if n:
            issue_h2d(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                issue_h2d(i + 1)
            main.wait_event(in_ready[slot])
            if i >= 2:
                main.wait_event(out_free[slot])
            slots[slot][3].replay()
            in_free[slot].record(main)
            out_ready[slot].record(main)
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(out_ready[slot])
                batches[i][2].copy_(slots[slot][2], non_blocking=True)
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

    def stylize_host_u8(self, content_u8_pinned: torch.Tensor, style_u8_pinned: torch.Tensor, out_u8_pinned: torch.Tensor,
                        normalize: bool = True) -> torch.Tensor:
        """uint8 [B,S,S,3] pinned host images in, stylised uint8 [B,S,S,3] pinned host images out (one blocking call)."""
        (d_content, d_style), d_output, graph = self._u8(normalize)
        d_content.copy_(content_u8_pinned, non_blocking=True)
        d_style.copy_(style_u8_pinned, non_blocking=True)
        graph.replay()
        out_u8_pinned.copy_(d_output, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_u8_pinned
