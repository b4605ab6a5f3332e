"""Data-parallel plumbing on CPU: world_size 2, gloo backend.  Checks that bucketed,
overlapped gradient averaging reproduces the full-batch gradient and that utterance
sharding covers the batch exactly once."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 8), torch.nn.Tanh(),
                               torch.nn.Linear(8, 1))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = pkg("dp")
    model = _model()
    if rank == 1:                      # deliberately different start: broadcast must fix it
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    dp.broadcast_parameters(model, src=0)
    sync = dp.GradAllReduce(model, bucket_mb=0.001)     # tiny buckets -> several all-reduces
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 16, generator=g), torch.randn(8, 1, generator=g)
    lo, hi = dp.shard_utterances(8, rank, world)
    for step in range(2):              # second step checks reset()
        sync.reset()
        loss = ((model(x[lo:hi]) - y[lo:hi]) ** 2).mean()
        loss.backward()
        sync.finish()
    out[rank] = [p.grad.clone() for p in model.parameters()] + [torch.tensor([lo, hi])]
    dist.destroy_process_group()


def test_gradient_allreduce_matches_full_batch():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    model = _model()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 16, generator=g), torch.randn(8, 1, generator=g)
    ((model(x) - y) ** 2).mean().backward()
    ref = [p.grad for p in model.parameters()]
    for rank in range(world):
        for got, want in zip(out[rank][:-1], ref):
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-7)
    assert out[0][-1].tolist() == [0, 4] and out[1][-1].tolist() == [4, 8]


def test_single_process_is_a_no_op():
    dp = pkg("dp")
    model = _model()
    sync = dp.GradAllReduce(model)
    x = torch.randn(4, 16)
    model(x).sum().backward()
    sync.finish()
    assert all(p.grad is not None and p.grad.abs().sum() > 0 for p in model.parameters())
    assert sync.grad_bytes() == sum(p.numel() * 4 for p in model.parameters())
    assert dp.shard_utterances(10, 0, 1) == (0, 10)
