#!/usr/bin/env python3
"""Attention core timing: python tools/mha_bench.py [B L H]  (forward variants and backward warp-group counts)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asr_b200
lib = asr_b200._lib; L = lib.lib(); ptr, sp, check = lib.ptr, lib.stream_ptr, lib.check
shapes = [(16, 2048, 8), (16, 512, 8), (4, 4096, 8)] if len(sys.argv) < 4 else [tuple(int(x) for x in sys.argv[1:4])]
for B, Ls, H in shapes:
    g = torch.Generator(device="cuda").manual_seed(5)
    q, k, v, do = (torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4))
    out = torch.empty_like(q); lse = torch.empty(B, H, Ls, device="cuda")
    gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    wsb = L.asr_mha_bwd_workspace_bytes(B, H, Ls, Ls, 64); ws = torch.empty(wsb // 4 + 1, device="cuda")
    def fwd(): check(L.asr_mha_fwd_bf16(ptr(q), ptr(k), ptr(v), None, None, 0, B, H, Ls, Ls, 64, 0.125, ptr(out), ptr(lse), sp()), "fwd")
    def bwd(): check(L.asr_mha_bwd_bf16(ptr(q), ptr(k), ptr(v), ptr(out), ptr(do), ptr(lse), None, None, 0, B, H, Ls, Ls, 64, 0.125, ptr(gq), ptr(gk), ptr(gv), ptr(ws), wsb, sp()), "bwd")
    def timed(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    flop = 4.0 * B * H * Ls * Ls * 64
    for var in (3, 4, 8, 10, 21):
        lib.set_option("mha_variant", var)
        ms = timed(fwd)
        print("B=%d L=%d H=%d fwd variant %d: %.3f ms  %.0f TFLOP/s" % (B, Ls, H, var, ms, flop / ms / 1e9), flush=True)
    lib.set_option("mha_variant", 0)
    for grp in (2, 4):
        lib.set_option("mha_bwd_groups", grp)
        ms = timed(bwd)
        print("B=%d L=%d H=%d bwd %2d softmax warps: %.3f ms  %.0f TFLOP/s" % (B, Ls, H, 4 * grp, ms, 2.5 * flop / ms / 1e9), flush=True)
    lib.set_option("mha_bwd_groups", 0)
    def fwd_d(): check(L.asr_mha_fwd_dropout_bf16(ptr(q), ptr(k), ptr(v), None, None, 0, B, H, Ls, Ls, 64, 0.125, 0.1, 77, ptr(out), ptr(lse), sp()), "fwd")
    def bwd_d(): check(L.asr_mha_bwd_dropout_bf16(ptr(q), ptr(k), ptr(v), ptr(out), ptr(do), ptr(lse), None, None, 0, B, H, Ls, Ls, 64, 0.125, 0.1, 77, ptr(gq), ptr(gk), ptr(gv), ptr(ws), wsb, sp()), "bwd")
    ms = timed(fwd_d); print("B=%d L=%d H=%d fwd with dropout 0.1: %.3f ms  %.0f TFLOP/s" % (B, Ls, H, ms, flop / ms / 1e9), flush=True)
    ms = timed(bwd_d); print("B=%d L=%d H=%d bwd with dropout 0.1: %.3f ms  %.0f TFLOP/s" % (B, Ls, H, ms, 2.5 * flop / ms / 1e9), flush=True)
