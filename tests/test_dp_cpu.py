"""CPU, gloo, world_size 2: the flat-bucket gradient averaging equals the single-process mean of the
per-shard gradients (the property generator/train.py:74-79 provides), with one collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gtos_b200.dp import FlatGradBucket, shard_range
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    data = torch.randn(8, 6)
    lo, hi = shard_range(8, rank, world)
    bucket = FlatGradBucket(model.parameters())
    bucket.zero()
    model(data[lo:hi]).pow(2).mean().backward()
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
    bucket.all_reduce_mean()
    q.put((rank, bucket.flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_single_process_mean():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    data = torch.randn(8, 6)
    grads = []
    for lo, hi in ((0, 4), (4, 8)):
        model.zero_grad()
        model(data[lo:hi]).pow(2).mean().backward()
        grads.append(torch.cat([p.grad.flatten() for p in model.parameters()]))
    ref = (grads[0] + grads[1]) / 2
    assert torch.allclose(got[0], ref, atol=1e-6) and torch.allclose(got[1], ref, atol=1e-6)


def _worker_overlap(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gtos_b200.dp import OverlappedGradBuckets, shard_range
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 1))
    unused = torch.nn.Parameter(torch.ones(3))                 # never receives a gradient: finish() must flush its group
    data = torch.randn(8, 6)
    lo, hi = shard_range(8, rank, world)
    # groups in backward order: last layer first
    groups = [list(model[4].parameters()), list(model[2].parameters()) + [unused], list(model[0].parameters())]
    b = OverlappedGradBuckets(groups)
    for _ in range(2):                                         # second round checks that zero() re-arms the hooks
        b.zero()
        model(data[lo:hi]).pow(2).mean().backward()
        b.finish()
    b.unpack()
    assert model[0].weight.grad.data_ptr() >= b.flat.data_ptr()
    q.put((rank, b.flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_buckets_match_single_process_mean():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 1))
    data = torch.randn(8, 6)
    grads = []
    for lo, hi in ((0, 4), (4, 8)):
        model.zero_grad()
        model(data[lo:hi]).pow(2).mean().backward()
        order = list(model[4].parameters()) + list(model[2].parameters()) + [None] + list(model[0].parameters())
        grads.append(torch.cat([p.grad.flatten() if p is not None else torch.zeros(3) for p in order]))
    ref = (grads[0] + grads[1]) / 2
    assert torch.allclose(got[0], ref, atol=1e-6) and torch.allclose(got[1], ref, atol=1e-6)


def test_shard_range_requires_equal_shards():
    import pytest
    from gtos_b200.dp import shard_range
    assert shard_range(128, 3, 8) == (48, 64)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
