#!/usr/bin/env python3
"""Runs LayerNorm(dropout(y) + residual) * mask forward and backward twice at the bench shape (for ncu captures), then a
fp32 and a bf16 split-K GEMM at a model shape: python tools/ln_once.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops")
M, D = 64 * 1600, 512
g = torch.Generator(device="cuda").manual_seed(10)
y = torch.randn(M, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
res = torch.randn(M, D, device="cuda", generator=g).requires_grad_(True)
gam = torch.ones(D, device="cuda", requires_grad=True)
bet = torch.zeros(D, device="cuda", requires_grad=True)
mask = (torch.rand(M, device="cuda", generator=g) < 0.8).float()
go = torch.randn(M, D, device="cuda", generator=g)
for _ in range(2):
    out = ops.residual_layer_norm(y, res, gam, bet, 1e-5, dropout_p=0.1, seed=11, row_scale=mask)
    out.backward(go)
a = torch.randn(1344, 2048, device="cuda", generator=g)
b = torch.randn(512, 2048, device="cuda", generator=g)
for _ in range(2):
    ops.gemm_f32(a, b)
    ops.gemm_bf16(a.bfloat16(), b.bfloat16())
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
