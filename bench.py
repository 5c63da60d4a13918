#!/usr/bin/env python
"""Headline benchmark: stylised images/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size 256|512] [--layers k]

N=1 workload = BASELINE configs[1]: zero-shot stylization, batch 32 at 256x256, synthetic tensors,
forward only, one transformer layer.  For N>1 (torchrun, one rank per GPU) every rank stylises its own
batch of 32 (images are independent: no data-path collective, weak scaling).

`value`  : whole-job images/s with inputs resident in HBM, one CUDA-graph launch per step, timed with
           CUDA events between barrier+synchronize pairs, max over ranks.
`e2e`    : same metric through the public host API at the reference's own image boundary (test_model.py:39-48,207):
           pinned uint8 HWC images in -> H2D into the graph's input buffers -> graph (ToTensor/Normalize inside the patch-embedding
           kernel's loads, clip*255 in the last convolution's epilogue) -> D2H of the uint8 stylised images, every step
           (GraphedStylizer.stylize_many(u8=True): two graphs, one per staging slot, copies overlapped with the previous / next
           graph).  `e2e_f32` is the same through fp32 NCHW pinned tensors (4x the bytes).
`roofline`: the tensor-core kernel family with the largest share of the step (chosen from the measured per-family times) --
           algorithmic FLOPs of its launches / their summed CUDA-event durations, against the measured sustained bf16 peak.
`cpu_baseline`: the CPU oracle port of the reference timed on this box's host cores (bounded sample).
`gpu_eager_baseline`: the reference's ops (the oracle restatement: plain torch, fp32, and with TF32 matmuls) on the same B200,
           eager -- SURVEY 8d's "beat this" number next to the CPU one.
`config5`: BASELINE configs[4] (512x512, batch 16/GPU) device-resident and end-to-end images/s in the same line.
`loss_forward`: the VGG-19 loss forward at batch 32 and 8: HBM GB/s of the reduction kernels against the measured copy rate.
`summary`: the headline numbers again, LAST in the line (a truncated tail of stdout still shows them).
`--impl reference`: the reference's CPU implementation of the path (its oracle port: the reference is
           Python and /root/reference does not exist on the GPU box), all host threads, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import torch  # noqa: E402

GF_PER_IMAGE = {256: {"swin": 7.552, "st": 8.556, "cnn": 7.965}, 512: {"swin": 29.609, "st": 34.226, "cnn": 31.860}}


def flops_per_image(size: int, layers: int) -> float:
    g = GF_PER_IMAGE[size]
    return (2 * g["swin"] + layers * g["st"] + g["cnn"]) * 1e9  # BASELINE.md section 2 (reference-algorithm FLOPs)


def ncu_traffic(kernel: str, size: int, batch: int):
    """DRAM bytes per launch of `kernel` (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the family's launches of
    one forward) from the committed `ncu --set full` capture of this workload (tools/ncu_traffic.py), or None."""
    p = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(f"b{batch}_{size}", {}).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d["bf16_tflops_sustained"], d["hbm_gbs"], "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        busy = [s for s in sm if s > 0.5 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_oracle_rate(size: int, layers: int, batch: int, iters: int, threads: int, min_seconds: float = 0.0, max_iters: int = 200):
    """images/s of the CPU oracle port (full forward, fp32) on a bounded sample: `iters` timed forwards, continued until
    `min_seconds` of timed work have accumulated (at most `max_iters`)."""
    from mastermetastyletransfer_b200 import synthetic
    from mastermetastyletransfer_b200.full_model import MasterStyleTransferModel
    from oracle import master_oracle as O
    torch.set_num_threads(threads)
    m = MasterStyleTransferModel()
    synthetic.fill_state_dict_(m, 0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    content, style = synthetic.synthetic_images(batch, size, seed=0)
    with torch.no_grad():
        O.full_forward(sd, content, style, layers)  # warm-up
        times = []
        while len(times) < iters or (sum(times) < min_seconds and len(times) < max_iters):
            t0 = time.perf_counter()
            O.full_forward(sd, content, style, layers)
            times.append(time.perf_counter() - t0)
    return batch / min(times), times


TRAIN_GF_PER_SAMPLE = {1: 253.34, 4: 330.35}  # fwd+bwd, 256^2, reference-algorithm FLOPs


def bench_training(dev, rank: int, world: int, steps: int, warmup: int, batch: int = 8, size: int = 256, layers: int = 1):
    """BASELINE configs[2] / [3]: the reference's inner-loop step (frozen encoder, style transformer + decoder trained with the
    VGG-19 loss, Adam) at batch 8/GPU, data-parallel across ranks with one gradient all-reduce; and one meta iteration
    (omega <- theta, one inner step on this rank's style task, all-reduced Reptile update).  Device-timed, max over ranks."""
    import torch.distributed as dist
    from mastermetastyletransfer_b200 import ops, synthetic
    from mastermetastyletransfer_b200.full_model import MasterStyleTransferModel
    from mastermetastyletransfer_b200.loss import custom_loss
    from mastermetastyletransfer_b200.training import GraphedTrainStep, InnerLoopTrainer, meta_iteration

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    model = MasterStyleTransferModel()
    synthetic.fill_state_dict_(model, 0)
    model = model.to(dev)
    loss_fn = custom_loss("/nonexistent")
    synthetic.fill_state_dict_(loss_fn, 1)
    loss_fn = loss_fn.to(dev)
    for m in (model.style_transformer.encoder, model.style_transformer.decoder):  # fixed-work steps: stochastic depth off
        m.stochastic_depth.p = 0.0
    model.style_transformer.encoder.encoder_stochastic_depth_prob = 0.0
    content, style = synthetic.synthetic_images(batch, size, seed=rank)
    style = style[:1].repeat(batch, 1, 1, 1)  # one style image repeated (train_only_inner_loop.py:491-496)
    content, style = content.to(dev), style.to(dev)
    from mastermetastyletransfer_b200.parallel import max_over_ranks
    out = {}
    identical = True
    for name, dp in (("train_step", True), ("train_step_eager", True), ("meta_step", False), ("meta_step_eager", False)):
        is_graphed = name in ("train_step", "meta_step")
        trainer = InnerLoopTrainer(model, loss_fn, inner_lr=1e-4, data_parallel=dp and world > 1, capturable=is_graphed)
        if name == "train_step":
            graphed = GraphedTrainStep(trainer, batch, size, layers)
            run = lambda: graphed.step(content, style)
        elif name == "train_step_eager":
            run = lambda: trainer.step(content, style, layers)
        elif name == "meta_step":  # omega <- theta, one graphed inner step, (all-reduced) Reptile update
            graphed = GraphedTrainStep(trainer, batch, size, layers)
            run = lambda: meta_iteration(trainer, style, [content], 1e-4, layers, graphed=graphed)
        else:
            run = lambda: meta_iteration(trainer, style, [content], 1e-4, layers)
        for _ in range(warmup):
            last = run()
        n0 = ops.launch_count
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            last = run()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1) / steps, dev)
        gf = TRAIN_GF_PER_SAMPLE.get(layers)
        if world > 1 and is_graphed:
            # multi-GPU self-check (tests/test_gpu_multi.py needs >= 2 GPUs and is skipped on a 1-GPU test box): after the timed
            # steps every rank must hold bit-identical parameters -- omega after data-parallel steps (same averaged gradient, same
            # Adam), theta after meta iterations (same averaged Reptile delta).  Checksums: per-tensor sum and sum of squares in
            # fp64, MAX and MIN over ranks must coincide.
            ps = trainer.params if dp else trainer._theta
            chk = torch.stack([torch.stack([p.detach().double().sum(), (p.detach().double() ** 2).sum()]) for p in ps]).flatten()
            hi, lo = chk.clone(), chk.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            identical = identical and bool(torch.equal(hi, lo))
        out[name] = {"ms_per_step": ms, "samples_per_s": world * batch / (ms / 1e3), "batch_per_gpu": batch, "size": size,
                     "layers": layers, "n_gpus": world, "launches_per_step": (ops.launch_count - n0) // steps + (graphed.launches if is_graphed else 0),
                     "tflops": (gf * 1e9 * batch / (ms * 1e9)) if gf and size == 256 else None,
                     "loss": [round(v, 5) for v in last.tolist()],
                     "collective": ("all-reduce of the 4.30 M fp32 gradient per step" if name.startswith("train_step") else
                                    "all-reduce of the 4.30 M fp32 (omega - theta) delta per outer iteration") if world > 1 else "none (1 rank)",
                     "timed": ("one CUDA-graph replay per inner step (training.GraphedTrainStep)" if is_graphed else
                               "eager launches through the C ABI (no CUDA graph)") + ", CUDA events, max over ranks"}
    out["replicas_identical"] = identical if world > 1 else None  # None: one rank, nothing to compare
    return out


def gpu_eager_baseline(model, content, style, layers, dev, iters: int = 5):
    """The reference's operator sequence (the oracle restatement: plain torch ops, fp32 parameters) run EAGERLY on this GPU:
    fp32 (cuBLAS/cuDNN fp32 math) and with TF32 tensor-core matmuls/convs allowed -- what a user of the reference gets on a B200
    without this library.  Not on the product path: bench.py's baseline leg only."""
    from oracle import master_oracle as O
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    out = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            for _ in range(2):
                O.full_forward(sd, content, style, layers)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                O.full_forward(sd, content, style, layers)
            b.record()
            torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b) / iters
        out[name] = {"value": content.shape[0] / (ms / 1e3), "unit": "images/s", "ms_per_step": ms}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out["what"] = f"oracle/master_oracle.full_forward (torch eager ops of the reference) on cuda, batch {content.shape[0]}, {iters} timed forwards"
    return out


def loss_reduction_rates(dev, B: int, size: int, hbm_gbs: float):
    """Device-side HBM rate of the loss reductions on the four tap shapes of a batch-B loss call (3B images in the statistics
    pass, content + output images in the content term): ten back-to-back launches of one op are captured in a CUDA graph and the
    replay is timed, so the number is the kernels' own time -- an eager launch through ctypes costs ~10 us of host time per call,
    more than the smaller taps take.  Taps that fit the 126 MB L2 are marked: their rate is not an HBM rate."""
    from mastermetastyletransfer_b200 import ops
    rec, tot_b, tot_t = {}, {"tap_stats": 0.0, "content_term": 0.0}, {"tap_stats": 0.0, "content_term": 0.0}
    for i, (hw, c) in enumerate(((size // 2, 128), (size // 4, 256), (size // 8, 512), (size // 16, 512))):
        T = hw * hw
        x = torch.randn(3 * B, T, c, device=dev).bfloat16()
        mean, var = torch.empty(3 * B, c, device=dev), torch.empty(3 * B, c, device=dev)
        part = torch.empty(592, device=dev)
        xv = x.view(3 * B, T * c)
        fns = {"tap_stats": (lambda: ops.tap_stats(x, mean, var, 3 * B, T, c), 2.0 * 3 * B * T * c),
               "content_term": (lambda: ops.content_term(xv[:B], xv[2 * B:], mean[:B], var[:B], mean[2 * B:], var[2 * B:], B, T, c, False, part),
                                2.0 * 2 * B * T * c)}
        for name, (fn, nbytes) in fns.items():
            fn()
            torch.cuda.synchronize(dev)
            st = torch.cuda.Stream(device=dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(st):
                fn()
                st.synchronize()
                with torch.cuda.graph(graph, stream=st):
                    for _ in range(10):
                        fn()
            torch.cuda.synchronize(dev)
            graph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                graph.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            us = e0.elapsed_time(e1) * 1e3 / 30
            rec[f"{name}_tap{i}"] = {"us": round(us, 2), "gbs": round(nbytes / us / 1e3, 1), "frac_of_hbm": round(nbytes / us / 1e3 / hbm_gbs, 3),
                                     "mbytes": round(nbytes / 1e6, 1), "fits_l2": nbytes < 100e6}
            tot_b[name] += nbytes
            tot_t[name] += us
        del x
    for name in tot_b:
        rec[name + "_all_taps"] = {"us": round(tot_t[name], 1), "gbs": round(tot_b[name] / tot_t[name] / 1e3, 1),
                                   "frac_of_hbm": round(tot_b[name] / tot_t[name] / 1e3 / hbm_gbs, 3)}
    return rec


def bench_loss_forward(dev, hbm_gbs: float, size: int = 256):
    """VGG-19 content/style loss forward (custom_loss.forward, rows a15-a18) at batch 32 and 8: whole-call milliseconds and the
    HBM rate of the reduction kernels (tap statistics incl. their finalisation, content term) summed over the four taps, from
    CUDA events around every launch; algorithmic bytes = each bf16 tap tensor read once per kernel."""
    from mastermetastyletransfer_b200 import ops, synthetic
    from mastermetastyletransfer_b200.loss import custom_loss
    loss_fn = custom_loss("/nonexistent")
    synthetic.fill_state_dict_(loss_fn, 1)
    loss_fn = loss_fn.to(dev).eval()
    out = {}
    for B in (32, 8):
        content, style = synthetic.synthetic_images(B, size, seed=3)
        outimg, _ = synthetic.synthetic_images(B, size, seed=4)
        c, s, o = content.to(dev), style.to(dev), outimg.to(dev)
        with torch.no_grad():
            for _ in range(2):
                loss_fn(c, s, o, output_content_and_style_loss=True)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                loss_fn(c, s, o, output_content_and_style_loss=True)
            b.record()
            torch.cuda.synchronize(dev)
            with ops.timing() as rec:
                loss_fn(c, s, o, output_content_and_style_loss=True)
            torch.cuda.synchronize(dev)
        fam = {}
        for name, flops, nbytes, e0, e1, _d in rec:
            f = fam.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0.0, "flops": 0.0})
            f["launches"] += 1
            f["ms"] += e0.elapsed_time(e1)
            f["bytes"] += nbytes
            f["flops"] += flops
        red = {k: {"launches": v["launches"], "us": round(1e3 * v["ms"], 1), "gbs": round(v["bytes"] / (v["ms"] * 1e6), 1),
                   "frac_of_hbm": round(v["bytes"] / (v["ms"] * 1e6) / hbm_gbs, 3)}
               for k, v in fam.items() if k in ("tap_stats_kernel", "content_term_kernel", "loss_finalize_kernel") and v["bytes"]}
        conv_ms = sum(v["ms"] for k, v in fam.items() if v["flops"])
        conv_fl = sum(v["flops"] for k, v in fam.items() if v["flops"])
        out[f"batch{B}"] = {"ms_per_call": a.elapsed_time(b) / 5, "reductions_eager_events": red,
                            "reductions": loss_reduction_rates(dev, B, size, hbm_gbs),
                            "vgg_convs": {"ms": round(conv_ms, 3), "tflops": round(conv_fl / (conv_ms * 1e9), 1) if conv_ms else None},
                            "roofline_hbm": {"peak_gbs": hbm_gbs, "bytes": "each bf16 tap tensor read once per kernel launch"}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=256, choices=[256, 512])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default 32 @256, 16 @512)")
    ap.add_argument("--layers", type=int, default=1)
    ap.add_argument("--cpu-baseline", type=int, default=1)
    ap.add_argument("--train-steps", type=int, default=5, help="timed steps of the secondary training-step measurement (0 = skip)")
    ap.add_argument("--config5", type=int, default=1, help="also measure BASELINE configs[4] (512x512, batch 16) in the same run")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: keep a private handle to it and point fd 1 at stderr, so that nothing a native
    # library prints (NCCL's version banner goes to stdout whatever NCCL_DEBUG_FILE says on some boxes) can end up in front of it
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    emit = lambda line: (json_out.write(line + "\n"), json_out.flush())
    args.warmup = max(args.warmup, 3)
    batch = args.batch or (32 if args.size == 256 else 16)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = f"zero-shot stylization, batch {batch}/GPU at {args.size}x{args.size}, forward only, {args.layers} transformer layer(s), window 8/shift 4"
    config = {"workload": workload, "batch_per_gpu": batch, "size": args.size, "layers": args.layers,
              "weights": "seeded random init (reference checkpoints unavailable offline)"}
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        t0 = time.perf_counter()
        cpu_batch = 4 if args.size == 256 else 1
        steps = min(args.steps, 10)
        _, times = cpu_oracle_rate(args.size, args.layers, cpu_batch, steps, threads)  # 1 warm-up + `steps` timed forwards
        ms = 1e3 * sum(times) / len(times)
        value = cpu_batch / (ms / 1e3)
        sample = f"{steps} steps of batch {cpu_batch} at {args.size}x{args.size} (CPU oracle port of the reference, fp32, {threads} threads)"
        config["cpu_sample_batch"] = cpu_batch  # the CPU arm times a bounded sample of the workload: batch 4 (1 at 512x512) per step
        config["workload"] += f" -- CPU arm: bounded sample, batch {cpu_batch} per step"
        emit(json.dumps({
            "impl": "reference", "metric": "stylized_images_per_sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}))
        return

    import torch.distributed as dist
    from mastermetastyletransfer_b200 import ops, synthetic
    from mastermetastyletransfer_b200.full_model import MasterStyleTransferModel
    from mastermetastyletransfer_b200.runtime import GraphedStylizer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the ONE JSON line (NCCL prints its banner to stdout)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from mastermetastyletransfer_b200.parallel import max_over_ranks

    model = MasterStyleTransferModel()
    synthetic.fill_state_dict_(model, 0)
    model = model.eval().to(dev)

    def timed(fn, steps):
        """device-side milliseconds for `fn()` (which enqueues `steps` steps), barrier + synchronize on both sides, max over ranks"""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b), dev)

    def measure_stylization(size, nbatch, steps, warmup, sample_clocks):
        """device-resident and end-to-end images/s of one (size, batch) workload; returns (record, runner, clocks)"""
        content, style = synthetic.synthetic_images(nbatch, size, seed=rank)
        runner = GraphedStylizer(model, nbatch, size, args.layers, dev)
        runner.load(content.to(dev), style.to(dev))
        torch.cuda.synchronize(dev)
        for _ in range(warmup):
            runner.replay()
        sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None
        ms = timed(lambda: [runner.replay() for _ in range(steps)], steps) / steps
        clocks = sampler.stop() if sampler else None
        rec = {"value": world * nbatch / (ms / 1e3), "ms_per_step": ms}
        # ---- end to end, uint8 image boundary (test_model.py:39-48,207): K distinct pinned host batches in, K pinned
        #      host results out, H2D of the next batch and D2H of the previous result overlap the running graph
        nbuf = min(steps, 4)
        gu = torch.Generator().manual_seed(100 + rank)
        host8 = [(torch.randint(0, 256, (nbatch, size, size, 3), generator=gu, dtype=torch.uint8).pin_memory(),
                  torch.randint(0, 256, (nbatch, size, size, 3), generator=gu, dtype=torch.uint8).pin_memory(),
                  torch.empty(nbatch, size, size, 3, dtype=torch.uint8).pin_memory()) for _ in range(nbuf)]
        batches8 = [host8[i % nbuf] for i in range(steps)]
        runner.stylize_many(batches8[:warmup], u8=True)
        ms8 = timed(lambda: runner.stylize_many(batches8, u8=True), steps) / steps
        h2d8, d2h8 = 2 * host8[0][0].numel(), host8[0][2].numel()
        rec["e2e"] = {"value": world * nbatch / (ms8 / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d8, "d2h_bytes_per_step": d2h8,
                      "ms_per_step": ms8, "h2d_gbs_per_rank": h2d8 / (ms8 * 1e6), "d2h_gbs_per_rank": d2h8 / (ms8 * 1e6),
                      "api": "GraphedStylizer.stylize_many(u8=True): pinned uint8 [B,S,S,3] images in / out (test_model.py's boundary), "
                             "ToTensor+Normalize and clip*255 on the device, copies overlapped with the neighbouring graphs"}
        c8, s8, o8 = host8[0]
        ms8s = timed(lambda: [runner.stylize_host_u8(c8, s8, o8) for _ in range(steps)], steps) / steps
        rec["e2e"]["single_blocking_call_ms"] = ms8s
        # ---- the same through fp32 NCHW pinned tensors (4x the bytes)
        host = [(content.pin_memory(), style.pin_memory(), torch.empty(nbatch, 3, size, size).pin_memory()) for _ in range(nbuf)]
        batches = [host[i % nbuf] for i in range(steps)]
        runner.stylize_many(batches[:warmup])
        ms32 = timed(lambda: runner.stylize_many(batches), steps) / steps
        c_pin, s_pin, o_pin = host[0]
        ms32s = timed(lambda: [runner.stylize_host(c_pin, s_pin, o_pin) for _ in range(steps)], steps) / steps
        rec["e2e_f32"] = {"value": world * nbatch / (ms32 / 1e3), "unit": "images/s", "ms_per_step": ms32,
                          "h2d_bytes_per_step": 8 * c_pin.numel(), "d2h_bytes_per_step": 4 * o_pin.numel(),
                          "h2d_gbs_per_rank": 8 * c_pin.numel() / (ms32 * 1e6), "single_blocking_call_ms": ms32s,
                          "api": "GraphedStylizer.stylize_many: pinned fp32 [B,3,S,S] tensors in / out"}
        rec["launches_per_step"] = runner.launches_per_step
        return rec, runner, clocks

    main_rec, runner, clocks = measure_stylization(args.size, batch, args.steps, args.warmup, True)
    ms_step, value = main_rec["ms_per_step"], main_rec["value"]

    launches_per_step = main_rec["launches_per_step"]
    # ---------------- per-kernel-family timing (eager pass with CUDA events around every launch) ----------------
    burst, sustained, hbm, peak_src = peaks()
    roofline, families = None, None
    if rank == 0:
        with ops.timing() as rec:
            with torch.no_grad():
                model(runner.content, runner.style, args.layers)
        torch.cuda.synchronize(dev)
        fam = {}
        if os.environ.get("MST_BENCH_DETAIL"):
            with open(os.environ["MST_BENCH_DETAIL"], "w") as fh:
                for name, flops, nbytes, a, b, desc in rec:
                    ms = a.elapsed_time(b)
                    fh.write(f"{name:24s} {ms*1e3:9.1f} us  {flops/(ms*1e9) if flops else 0:8.1f} TF/s  {nbytes/(ms*1e6) if nbytes else 0:8.1f} GB/s  {desc}\n")
        for name, flops, nbytes, a, b, _desc in rec:
            f = fam.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            f["launches"] += 1
            f["ms"] += a.elapsed_time(b)
            f["flops"] += flops
            f["bytes"] += nbytes
        tot_ms = sum(f["ms"] for f in fam.values())
        families = {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / tot_ms, 4),
                        "tflops": round(v["flops"] / (v["ms"] * 1e9), 2) if v["flops"] else None,
                        "gbs": round(v["bytes"] / (v["ms"] * 1e6), 1) if v["bytes"] else None} for k, v in fam.items()}
        dom = max((k for k in fam if fam[k]["flops"]), key=lambda k: fam[k]["ms"])  # dominant tensor-core kernel family
        g = fam[dom]
        achieved = g["flops"] / (g["ms"] * 1e9)
        roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                    "frac": achieved / sustained, "traffic": ncu_traffic(dom, args.size, batch), "peak_source": f"bf16_tflops_sustained, {peak_src}",
                    "launches_per_step": g["launches"], "share_of_step": g["ms"] / tot_ms,
                    "step": {"achieved": flops_per_image(args.size, args.layers) * batch / (ms_step * 1e9), "unit": "TFLOP/s",
                             "frac": flops_per_image(args.size, args.layers) * batch / (ms_step * 1e9) / sustained}}

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and args.cpu_baseline:
        cpu_batch = 4 if args.size == 256 else 1
        rate, times = cpu_oracle_rate(args.size, args.layers, cpu_batch, 3, threads, min_seconds=10.0)
        cpu = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port", "median": cpu_batch / statistics.median(times),
               "sample": f"best of {len(times)} full forwards of batch {cpu_batch} at {args.size}x{args.size} ({sum(times):.1f} s of CPU work), "
                         "fp32 CPU oracle port of the reference"}

    # ---------------- the reference's ops on this GPU, eager (SURVEY 8d "GPU reference baseline") ----------------
    gpu_eager = None
    if rank == 0 and world == 1 and args.cpu_baseline:
        gpu_eager = gpu_eager_baseline(model, runner.content, runner.style, args.layers, dev)

    # ---------------- BASELINE configs[4]: 512x512, batch 16/GPU, in the same line (at every N) ----------------
    config5 = None
    if args.size == 256 and args.config5:
        del runner
        torch.cuda.empty_cache()
        rec5, runner5, _ = measure_stylization(512, 16, max(3, args.steps // 2), 3, False)
        rec5.pop("e2e_f32", None)
        config5 = {"workload": "zero-shot stylization, batch 16/GPU at 512x512, forward only, %d transformer layer(s)" % args.layers,
                   "unit": "images/s", **rec5,
                   "tflops": flops_per_image(512, args.layers) * 16 / (rec5["ms_per_step"] * 1e9),
                   "frac_of_sustained_bf16": flops_per_image(512, args.layers) * 16 / (rec5["ms_per_step"] * 1e9) / sustained}
        del runner5
        runner = None
        torch.cuda.empty_cache()

    # ---------------- VGG-19 loss forward: HBM rate of the reduction kernels ----------------
    loss_forward = bench_loss_forward(dev, hbm) if (rank == 0 and args.size == 256) else None

    # ---------------- secondary: training-step workloads (BASELINE configs[2], [3]) ----------------
    training = None
    if args.train_steps > 0 and args.size == 256:
        runner = None
        torch.cuda.empty_cache()
        training = bench_training(dev, rank, world, args.train_steps, 3)

    if rank == 0:
        act_mb = batch * (args.size // 8) ** 2 * 256 * 4 / 1e6
        config["l2"] = f"no explicit flush: a step streams >1 GB of activations (feature map alone {act_mb:.0f} MB fp32 x dozens of tensors) through a 126 MB L2"
        config["timed"] = "one CUDA-graph replay per step"
        tr = (training or {}).get("train_step") or {}
        summary = {"images_per_s": round(value, 1), "ms_per_step": round(ms_step, 4), "e2e_u8_images_per_s": round(main_rec["e2e"]["value"], 1),
                   "e2e_f32_images_per_s": round(main_rec["e2e_f32"]["value"], 1), "h2d_gbs_per_rank_u8": round(main_rec["e2e"]["h2d_gbs_per_rank"], 2),
                   "step_frac_of_bf16_peak": round(roofline["step"]["frac"], 4) if roofline else None,
                   "config5_images_per_s": round(config5["value"], 1) if config5 else None,
                   "config5_e2e_images_per_s": round(config5["e2e"]["value"], 1) if config5 else None,
                   "train_step_ms": round(tr["ms_per_step"], 4) if tr else None,
                   "replicas_identical": (training or {}).get("replicas_identical"),
                   "gpu_eager_fp32_images_per_s": round(gpu_eager["fp32"]["value"], 1) if gpu_eager else None,
                   "n_gpus": world}
        emit(json.dumps({
            "metric": "stylized_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": main_rec["e2e"], "e2e_f32": main_rec["e2e_f32"],
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "roofline": roofline, "cpu_baseline": cpu, "kernel_families": families, "training": training,
            "loss_forward": loss_forward, "gpu_eager_baseline": gpu_eager, "config5": config5, "summary": summary}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
