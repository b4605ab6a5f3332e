#!/usr/bin/env python3
"""Runs the attention forward/backward once at one shape (for ncu captures): python tools/mha_once.py B L H [causal]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asr_b200
lib = asr_b200._lib; L = lib.lib(); ptr, sp, check = lib.ptr, lib.stream_ptr, lib.check
B, Ls, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
causal = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if os.environ.get("ASR_MHA_VARIANT"): lib.set_option("mha_variant", int(os.environ["ASR_MHA_VARIANT"]))
g = torch.Generator(device="cuda").manual_seed(5)
q, k, v, do = (torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4))
out = torch.empty_like(q); lse = torch.empty(B, H, Ls, device="cuda")
gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
wsb = L.asr_mha_bwd_workspace_bytes(B, H, Ls, Ls, 64); ws = torch.empty(wsb // 4 + 1, device="cuda")
for _ in range(2):
    check(L.asr_mha_fwd_bf16(ptr(q), ptr(k), ptr(v), None, None, causal, B, H, Ls, Ls, 64, 0.125, ptr(out), ptr(lse), sp()), "fwd")
    check(L.asr_mha_bwd_bf16(ptr(q), ptr(k), ptr(v), ptr(out), ptr(do), ptr(lse), None, None, causal, B, H, Ls, Ls, 64, 0.125, ptr(gq), ptr(gk), ptr(gv), ptr(ws), wsb, sp()), "bwd")
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
