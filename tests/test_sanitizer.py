"""compute-sanitizer over one small call of every kernel family (memcheck: out-of-bounds / misaligned accesses;
racecheck on the shared-memory pipelines of the CTC lattice).  Slow (the tool serialises and instruments every launch), so it
has its own marker:  python -m pytest tests -m gpu_sanitize"""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu_sanitize
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SANITIZER = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"

# the smoke entry point (CIF forward / backward, CTC through the reference's entry point, attention forward) plus the
# kernels it does not reach: attention backward, the persistent forward kernel, the fp32 / bf16 GEMMs (unsplit, split-K inside
# a cluster - distributed shared memory -, unaligned outputs through shared memory), the column sums over row-chunk clusters,
# dropout + residual + LayerNorm forward / backward, the one-rank all-reduce
SCRIPT = r"""
import importlib, sys, tempfile
sys.path.insert(0, %r)
import torch
import __graft_entry__ as ge
ge.smoke()
pkg = importlib.import_module(ge.PKG)
ops, lib = importlib.import_module(ge.PKG + ".ops"), importlib.import_module(ge.PKG + "._lib")
g = torch.Generator().manual_seed(3)
for variant in (21, 40):
    lib.set_option("mha_variant", variant)
    q, k, v = (torch.randn(3, 300, 2, 64, generator=g).cuda().to(torch.bfloat16).requires_grad_(True) for _ in range(3))
    out = ops.mha_core(q, k, v, kv_len=torch.tensor([300, 211, 129], dtype=torch.int32).cuda(), causal=(variant == 40))
    out.float().square().sum().backward()
lib.set_option("mha_variant", 0)
x = torch.randn(200, 96, device="cuda", requires_grad=True)
w = torch.randn(77, 96, device="cuda", requires_grad=True)
ops.linear_f32_autograd(x, w, None).square().sum().backward()
a = torch.randn(300, 1000, device="cuda")
b = torch.randn(132, 1000, device="cuda")
bias = torch.randn(132, device="cuda")
for splits in (0, 3, 8):
    lib.set_option("gemm_split_k", splits)
    ops.gemm_f32(a, b, bias=bias)
    ops.gemm_f32(a.t().contiguous(), b.t().contiguous(), a_mn_major=True, b_mn_major=True)
    ops.gemm_bf16(a.bfloat16(), b.bfloat16(), bias=bias, relu=True)
    ops.gemm_bf16(a.bfloat16(), b.bfloat16(), out_dtype=torch.float32)
lib.set_option("gemm_split_k", 0)
ops.gemm_f32(a[:, :512].contiguous(), torch.randn(4233, 512, device="cuda"))              # rows of 4233 floats: staged epilogue
ops.gemm_bf16(a[:, :512].contiguous().bfloat16(), torch.randn(4233, 512, device="cuda").bfloat16())
ops.colsum(torch.randn(9000, 70, device="cuda").bfloat16()); ops.colsum(torch.randn(9000, 33, device="cuda"))
for dt in (torch.float32, torch.bfloat16):
    y = torch.randn(5, 77, 512, device="cuda").to(dt).requires_grad_(True)
    r = torch.randn(5, 77, 512, device="cuda", requires_grad=True)
    gam, bet = torch.ones(512, device="cuda", requires_grad=True), torch.zeros(512, device="cuda", requires_grad=True)
    m = (torch.rand(5, 77, 1, device="cuda") < 0.8)
    ops.residual_layer_norm(y, r, gam, bet, 1e-5, dropout_p=0.1, seed=5, row_scale=m).square().sum().backward()
    ops.residual_layer_norm(y.detach().requires_grad_(True), None, gam, bet).sum().backward()
import torch.distributed as dist
dp = importlib.import_module(ge.PKG + ".dp")
dist.init_process_group("nccl", store=dist.FileStore(tempfile.mktemp(prefix="asr_san_"), 1), rank=0, world_size=1,
                        device_id=torch.device("cuda", 0))
if dp.PeerAllReduce.available(torch.device("cuda", 0)):
    ar = dp.PeerAllReduce(40004, torch.device("cuda", 0), ctas=4)
    ar.flat.normal_()
    ar.launch(); ar.wait()
torch.cuda.synchronize()
dist.destroy_process_group()
print("sanitized run complete")
""" % ROOT


@pytest.mark.parametrize("tool", ["memcheck"])
def test_kernels_are_clean_under_compute_sanitizer(tool):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not os.path.exists(SANITIZER):
        pytest.skip("compute-sanitizer not installed")
    # (--report-api-errors no: under the tool torch's symmetric-memory set-up probes the multicast driver calls, gets
    # CUDA_ERROR_INVALID_VALUE and carries on without multicast - a host API return code, not a kernel fault)
    out = subprocess.run([SANITIZER, "--tool", tool, "--report-api-errors", "no", "--error-exitcode", "7", sys.executable, "-c", SCRIPT],
                         capture_output=True, text=True, timeout=1800, cwd=ROOT)
    tail = (out.stdout + out.stderr)[-4000:]
    assert "sanitized run complete" in out.stdout, tail
    assert out.returncode == 0 and "ERROR SUMMARY: 0 errors" in (out.stdout + out.stderr), tail
