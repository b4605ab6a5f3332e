#!/usr/bin/env python3
"""Runs the CTC kernels once at one shape (for ncu captures).
    [ASR_CTC_CHUNKS=n] python tools/ctc_once.py B T S [grad] [lattice variant]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200
from helpers import make_ctc_inputs
lib = asr_b200._lib; L = lib.lib(); ptr, sp, check = lib.ptr, lib.stream_ptr, lib.check
B, T, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
grad = len(sys.argv) > 4 and sys.argv[4] == "1"
if len(sys.argv) > 5: lib.set_option("ctc_lattice_variant", int(sys.argv[5]))
if os.environ.get("ASR_CTC_CHUNKS"): lib.set_option("ctc_chunks", int(os.environ["ASR_CTC_CHUNKS"]))
V = 4233
logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=1236)
tgt_len = targets.ne(0).sum(1).to(torch.int32)
nll = torch.empty(B, device="cuda"); g = torch.empty_like(logits) if grad else None
wsb = L.asr_ctc_workspace_bytes(B, T, V, S); ws = torch.empty(wsb // 4 + 1, device="cuda")
for _ in range(2):
    check(L.asr_ctc_fwd_bwd_f32(ptr(logits), ptr(targets), ptr(in_len), ptr(tgt_len), B, T, V, S, V - 1, ptr(nll), ptr(g), ptr(ws), wsb, sp()), "ctc")
torch.cuda.synchronize()
print("nll[0:4]", nll[:4].tolist())
