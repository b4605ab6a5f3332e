#!/usr/bin/env python3
"""fp32 GEMM epilogue: thread-per-row 16-byte stores (direct) vs rows through shared memory (gemm_stage_out = 2) at the
vocabulary projection's shapes: python tools/gemm_f32_epilogue_probe.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops")
lib = importlib.import_module("end-to-end_asr_pytorch_b200._lib")


def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator(device="cuda").manual_seed(1)
M, V, H = 102400, 4233, 512
h = torch.randn(M, H, device="cuda", generator=g)
w = torch.randn(V, H, device="cuda", generator=g) * H ** -0.5
bias = torch.randn(V, device="cuda", generator=g)
logits = torch.empty(M, 4236, device="cuda")
gl = torch.randn(M, 4236, device="cuda", generator=g)
gh = torch.empty(M, H, device="cuda")
for stage in (0, 2):
    lib.set_option("gemm_stage_out", stage)
    f = timed(lambda: ops.gemm_f32(h, w, bias=bias, out=logits))
    dx = timed(lambda: ops.gemm_f32(gl[:, :V], w, b_mn_major=True, out=gh))
    dw = timed(lambda: ops.gemm_f32(gl[:, :V], h, a_mn_major=True, b_mn_major=True))
    flop = 2.0 * M * V * H
    print("stage_out %d: fwd %.3f ms %.0f TFLOP/s | dX %.3f ms %.0f | dW %.3f ms %.0f" % (stage, f, flop / f / 1e9, dx, flop / dx / 1e9, dw, flop / dw / 1e9), flush=True)
lib.set_option("gemm_stage_out", 0)
