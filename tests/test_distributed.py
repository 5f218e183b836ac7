"""world_size-2 gloo test (CPU) of the data-parallel host logic: sharding covers every row once and
the flat-bucket gradient all-reduce reproduces the single-process gradient of the global mean loss."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepdow_b200.distributed import GradAllReducer, global_mean_loss, shard_batch, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 4097):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                rows += list(range(lo, hi))
            assert rows == list(range(n))
            sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)
    with pytest.raises(ValueError):
        shard_batch([torch.zeros(4, 2), torch.zeros(3)], 0, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    X = torch.randn(11, 6)
    y = torch.randn(11, 3)
    (Xl, yl) = shard_batch([X, y], rank, world)
    per_sample = ((net(Xl) - yl) ** 2).sum(1)
    global_mean_loss(per_sample, X.shape[0]).backward()
    GradAllReducer(net.parameters(), average=False)()
    got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    if rank == 0:
        net.zero_grad()
        ((net(X) - y) ** 2).sum(1).mean().backward()
        want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        ret["err"] = float((got - want).abs().max())
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["err"] < 1e-6
