#!/usr/bin/env python3
"""Fused linear layers vs torch (cuBLAS + separate elementwise kernels): python tools/gemm_bench.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib
ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops")
F = torch.nn.functional


def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

g = torch.Generator(device="cuda").manual_seed(1)
for M in (10688, 64 * 1600):
    d, di = 512, 2048
    x = torch.randn(M, d, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(di, d, device="cuda", generator=g) * d ** -0.5).bfloat16(); b1 = torch.randn(di, device="cuda", generator=g)
    w2 = (torch.randn(d, di, device="cuda", generator=g) * di ** -0.5).bfloat16(); b2 = torch.randn(d, device="cuda", generator=g)
    gam, bet = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    h = ops.linear_act(x, w1, b1, relu=True)
    b1h, b2h = b1.bfloat16(), b2.bfloat16()
    gamh, beth = gam.bfloat16(), bet.bfloat16()
    rows = [("w_1 + bias + relu  [M,512]x[2048,512]", lambda: ops.linear_act(x, w1, b1, relu=True), lambda: torch.relu(F.linear(x, w1, b1h)), 2.0 * M * d * di),
            ("w_2 + bias + residual + LayerNorm [M,2048]x[512,2048]", lambda: ops.linear_residual_layernorm(h, w2, b2, x, gam, bet),
             lambda: F.layer_norm(F.linear(h, w2, b2h) + x, (d,), gamh, beth), 2.0 * M * d * di),
            ("fc + bias + residual + LayerNorm [M,512]x[512,512]", lambda: ops.linear_residual_layernorm(x, w2[:, :512].contiguous(), b2, x, gam, bet),
             lambda: F.layer_norm(F.linear(x, w2[:, :512].contiguous(), b2h) + x, (d,), gamh, beth), 2.0 * M * d * d)]
    lib = importlib.import_module("end-to-end_asr_pytorch_b200._lib")
    for var in (1, 2, 3):
        lib.set_option("gemm_variant", var)
        a = timed(rows[0][1])
        print("M=%6d w_1 variant %d: %.3f ms %6.0f TFLOP/s" % (M, var, a, rows[0][3] / a / 1e9), flush=True)
    lib.set_option("gemm_variant", 0)
    for name, ours, ref, flop in rows:
        a, b = timed(ours), timed(ref)
        print("M=%6d %-56s ours %.3f ms %6.0f TFLOP/s | torch %.3f ms %6.0f TFLOP/s" % (M, name, a, flop / a / 1e9, b, flop / b / 1e9), flush=True)

# fp32 GEMM on the tensor cores (3xTF32) vs torch's fp32 (cuBLAS SIMT sgemm) and torch with TF32 allowed
for M, N, K in ((1344, 2048, 512), (1344, 4233, 512), (10688, 2048, 512), (102400, 2048, 512), (102400, 512, 2048)):
    x = torch.randn(M, K, device="cuda", generator=g); w = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
    b = torch.randn(N, device="cuda", generator=g)
    flop = 2.0 * M * N * K
    a = timed(lambda: ops.linear_f32(x, w, b))
    torch.backends.cuda.matmul.allow_tf32 = False
    t32 = timed(lambda: F.linear(x, w, b))
    torch.backends.cuda.matmul.allow_tf32 = True
    ttf = timed(lambda: F.linear(x, w, b))
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = F.linear(x.double(), w.double(), b.double())
    e_ours = ((ops.linear_f32(x, w, b).double() - ref).abs().max() / ref.abs().max()).item()
    e_t32 = ((F.linear(x, w, b).double() - ref).abs().max() / ref.abs().max()).item()
    print("fp32 linear M=%6d N=%4d K=%4d: 3xTF32 %.3f ms %5.0f TFLOP/s err %.1e | torch fp32 %.3f ms %5.0f TFLOP/s err %.1e | torch tf32 %.3f ms %5.0f TFLOP/s"
          % (M, N, K, a, flop / a / 1e9, e_ours, t32, flop / t32 / 1e9, e_t32, ttf, flop / ttf / 1e9), flush=True)
