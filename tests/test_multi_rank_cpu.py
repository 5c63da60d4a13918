"""CPU suite: the N>1 plumbing under gloo, world_size 2 (image sharding, max-over-ranks timing)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mastermetastyletransfer_b200 import parallel


def test_shard_range_covers_everything_once():
    for total in (1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = parallel.shard_range(33, rank, world)
        images = torch.arange(33.0)[lo:hi]
        mx = parallel.max_over_ranks(10.0 + rank)
        total = parallel.sum_over_ranks(float(images.sum()))
        count = parallel.sum_over_ranks(float(hi - lo))
        dist.barrier()
        q.put((rank, mx, total, count))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mx, total, count in out:
        assert mx == 11.0            # max over ranks, as bench.py reduces its device timings
        assert total == sum(range(33))  # every image processed exactly once
        assert count == 33


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mastermetastyletransfer_b200.training import allreduce_gradients
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2), requires_grad=False)]
        params[0].grad = torch.full((3, 5), float(rank + 1))
        params[1].grad = torch.arange(7.0) * (rank + 1)
        flat = allreduce_gradients(params)
        ok_views = params[0].grad.data_ptr() == flat.data_ptr() and params[1].grad.data_ptr() == flat[15:].data_ptr()
        flat2 = allreduce_gradients(params, flat=flat)  # reuses the buffer on later steps
        q.put((rank, params[0].grad.clone(), params[1].grad.clone(), flat.numel(), ok_views, flat2.data_ptr() == flat.data_ptr()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce():
    """The DP collective of the training step: one flat all-reduce, mean over ranks, p.grad re-pointed into the buffer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g0, g1, n, ok_views, reused in out:
        assert n == 22 and ok_views and reused  # only trainable parameters are packed
        assert torch.allclose(g0, torch.full((3, 5), 1.5))  # mean of 1 and 2, then mean of equal values again
        assert torch.allclose(g1, torch.arange(7.0) * 1.5)


def test_tensor_table_layout_is_consistent():
    """Host-side chunk table of the multi-tensor optimiser kernels (no GPU needed for the arithmetic)."""
    from mastermetastyletransfer_b200 import _lib
    chunk = _lib.lib().mst_opt_chunk_elems()
    sizes = [1, chunk - 1, chunk, chunk + 1, 5 * chunk + 7]
    starts, c = [], 0
    for n in sizes:
        starts.append(c)
        c += -(-n // chunk)
    assert starts == [0, 1, 2, 3, 5] and c == 11
