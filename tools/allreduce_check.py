#!/usr/bin/env python3
"""Correctness + bandwidth of the peer-memory gradient all-reduce (csrc/allreduce.cu) against NCCL, N ranks of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_check.py [floats]"""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib
dp = importlib.import_module("end-to-end_asr_pytorch_b200.dp")

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 52287873            # the recipe's CIF_Model: 209 MB of fp32 gradients
say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)
assert dp.PeerAllReduce.available(dev)
for ctas, mc in ((16, True), (32, True), (64, True), (16, False), (32, False), (64, False)):
    try:
        ar = dp.PeerAllReduce(n, dev, ctas=ctas, multicast=mc)
    except RuntimeError as e:
        say("skipped (%s)" % e)
        continue
    say("world %d, %d floats (%.1f MB), %d CTAs, flavour: %s" % (world, n, 4 * n / 1e6, ctas, ar.flavour()))
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    for it in range(3):                                             # three rounds back to back: the flag words are reused
        x = torch.randn(n, device=dev, generator=g)
        ar.flat.copy_(x)
        ref = x.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref /= world
        ar.launch()
        ar.wait()
        torch.cuda.synchronize()
        err = (ar.flat - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert err <= 1e-6 * max(scale, 1.0), ("mismatch", it, err, scale)
        # every rank must hold the same bits
        chk = ar.flat.view(torch.int32).sum(dtype=torch.int64)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert int(lo) == int(hi), "ranks disagree"
    say("  correct (max error vs NCCL %.2e of %.2e), identical on all ranks" % (err, scale))
    def timed(fn, iters=20):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    def ours():
        ar.launch(); ar.wait()
    buf = torch.randn(n, device=dev)
    nccl_one = lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    per = 25 * 1024 * 1024 // 4
    chunks = list(buf.split(per))
    def nccl_buckets():
        hs = [dist.all_reduce(c, op=dist.ReduceOp.AVG, async_op=True) for c in chunks]
        for h in hs:
            h.wait()
    t_ours, t_one, t_b = timed(ours), timed(nccl_one), timed(nccl_buckets)
    alg = lambda ms: 4 * n / ms / 1e6
    say("  ours %.3f ms (%.0f GB/s algorithmic) | NCCL one call %.3f ms (%.0f GB/s) | NCCL 25 MB buckets %.3f ms (%.0f GB/s)" % (
        t_ours, alg(t_ours), t_one, alg(t_one), t_b, alg(t_b)))
    del ar
dist.destroy_process_group()
