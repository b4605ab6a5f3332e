#!/usr/bin/env python3
"""Attention core forward / backward: TFLOP/s per kernel variant on the microbench shapes + max error vs an fp32 reference.
    python tools/mha_bench.py [variant ...]      (mha_variant values; default 21 22)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
lib = bench.pkg("_lib"); L = lib.lib(); p, sp = lib.ptr, lib.stream_ptr
variants = [int(x) for x in sys.argv[1:]] or [21, 22]
g = torch.Generator(device="cuda").manual_seed(5)
for (B, Ls, H) in ((16, 2048, 8), (8, 4096, 8), (64, 512, 8)):
    q, k, v, do = (torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4))
    out = torch.empty_like(q); lse = torch.empty(B, H, Ls, device="cuda")
    qs, ks, vs = (t[:1].float().permute(0, 2, 1, 3) for t in (q, k, v))
    ref = torch.softmax(qs @ ks.transpose(-1, -2) * 0.125, -1) @ vs
    ref = ref.permute(0, 2, 1, 3)
    for var in variants:
        lib.set_option("mha_variant", var)
        fwd = lambda: lib.check(L.asr_mha_fwd_bf16(p(q), p(k), p(v), None, None, 0, B, H, Ls, Ls, 64, 0.125, p(out), p(lse), sp()), "fwd")
        ms = min(bench.cuda_time(fwd, 10, warm=3) for _ in range(3))
        err = (out[:1].float() - ref).abs().max().item() / ref.abs().max().item()
        print("B=%d L=%d variant %d: fwd %.4f ms  %.1f TFLOP/s  max err %.2e of scale" % (B, Ls, var, ms, 4.0 * B * H * Ls * Ls * 64 / ms / 1e9, err), flush=True)
    lib.set_option("mha_variant", 0)
