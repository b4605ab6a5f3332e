import importlib, os, sys
import torch
sys.path.insert(0, "/root/repo")
ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops")
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randn(15030, 512, device="cuda", generator=g).bfloat16()
b = torch.randn(512, 512, device="cuda", generator=g).bfloat16()
bias = torch.randn(512, device="cuda", generator=g)
for _ in range(3):
    ops.gemm_bf16(a, b, bias=bias)
torch.cuda.synchronize()
print("ok")
