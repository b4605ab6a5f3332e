"""The peer-memory gradient all-reduce (csrc/allreduce.cu, dp.PeerAllReduce) on the GPU: plumbing + kernel in a one-rank
group on any box, and against NCCL on two ranks where the box has two GPUs (tools/allreduce_check.py under torchrun)."""
import os
import subprocess
import sys
import tempfile

import pytest
import torch

from helpers import pkg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_one_rank_group_round_trip():
    """World size 1: the symmetric allocation, the handle tables and the kernel's barrier / slice arithmetic all run; the mean
    over one rank is the input itself.  Ranges (offset, length) leave the rest of the bucket alone."""
    import torch.distributed as dist
    dp = pkg("dp")
    created = False
    if not dist.is_initialized():
        store = dist.FileStore(tempfile.mktemp(prefix="asr_ar_"), 1)
        dist.init_process_group("nccl", store=store, rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        if not dp.PeerAllReduce.available(torch.device("cuda", 0)):
            pytest.skip("symmetric memory is not available in this torch build")
        n = 1_000_003
        ar = dp.PeerAllReduce(n, torch.device("cuda", 0), ctas=8)
        x = torch.randn(n, device="cuda")
        for _ in range(3):                       # flag words are reused from launch to launch
            ar.flat.copy_(x)
            ar.launch()
            ar.wait()
            torch.cuda.synchronize()
            assert torch.equal(ar.flat, x)
        ar.flat.copy_(x)
        ar.launch(offset=4096, numel=8192)
        ar.wait()
        torch.cuda.synchronize()
        assert torch.equal(ar.flat, x)
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_match_nccl():
    """Two ranks: bit-identical to NCCL's sum / 2 on every rank, three rounds back to back, both flavours (multicast where
    the box has it)."""
    env = dict(os.environ)
    env.pop("RANK", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(ROOT, "tools", "allreduce_check.py"), "3000003"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("correct") >= 2, out.stdout[-3000:]
