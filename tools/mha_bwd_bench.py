#!/usr/bin/env python3
"""Attention backward (delta + main kernel + dQ convert): time per call for the softmax-backward layouts
(mha_bwd_groups: 0 default, 2 = 8 warps, 5 = 16 warps + dedicated dQ warps).
    python tools/mha_bwd_bench.py [groups ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
lib = bench.pkg("_lib"); L = lib.lib(); p, sp = lib.ptr, lib.stream_ptr
groups = [int(x) for x in sys.argv[1:]] or [0, 5]
g = torch.Generator(device="cuda").manual_seed(5)
for (B, Ls, H, drop, causal) in ((16, 2048, 8, 0.0, 0), (8, 4096, 8, 0.0, 0), (16, 2048, 8, 0.1, 0), (16, 2048, 8, 0.0, 1), (90, 167, 8, 0.1, 0), (64, 512, 8, 0.0, 0)):
    q, k, v, do = (torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4))
    o = torch.empty_like(q); lse = torch.empty(B, H, Ls, device="cuda")
    lib.check(L.asr_mha_fwd_dropout_bf16(p(q), p(k), p(v), None, None, causal, B, H, Ls, Ls, 64, 0.125, drop, 1234, p(o), p(lse), sp()), "fwd")
    gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    wsb = L.asr_mha_bwd_workspace_bytes(B, H, Ls, Ls, 64)
    ws = torch.empty(wsb // 4 + 1, device="cuda")
    res, outs = [], {}
    for grp in groups:
        lib.set_option("mha_bwd_groups", grp)
        bwd = lambda: lib.check(L.asr_mha_bwd_dropout_bf16(p(q), p(k), p(v), p(o), p(do), p(lse), None, None, causal, B, H, Ls, Ls, 64,
                                                           0.125, drop, 1234, p(gq), p(gk), p(gv), p(ws), wsb, sp()), "bwd")
        ms = min(bench.cuda_time(bwd, 10, warm=3) for _ in range(3))
        outs[grp] = (gq.float().clone(), gk.float().clone(), gv.float().clone())
        pairs = B * (Ls * (Ls + 1) / 2 if causal else Ls * Ls)
        res.append("groups %d: %.4f ms %.0f TFLOP/s" % (grp, ms, 10.0 * H * 64 * pairs / ms / 1e9))
    lib.set_option("mha_bwd_groups", 0)
    diff = max((a - b).abs().max().item() / (b.abs().max().item() + 1e-9) for a, b in zip(outs[groups[0]], outs[groups[-1]]))
    print("B=%d L=%d drop=%.1f causal=%d: " % (B, Ls, drop, causal) + " | ".join(res) + " | max diff %.1e" % diff, flush=True)
