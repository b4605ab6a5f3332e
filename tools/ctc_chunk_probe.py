#!/usr/bin/env python3
"""Times the whole CTC call (K1+K2+K3) for different batch-slice counts of the multi-stream pipeline."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200
from helpers import make_ctc_inputs
lib = asr_b200._lib; L = lib.lib(); ptr, sp, check = lib.ptr, lib.stream_ptr, lib.check
V = 4233
for (B, T, S) in [(256, 1600, 80), (32, 1600, 80), (64, 400, 20)]:
    logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=1236)
    tgt_len = targets.ne(0).sum(1).to(torch.int32)
    nll = torch.empty(B, device="cuda"); g = torch.empty_like(logits)
    wsb = L.asr_ctc_workspace_bytes(B, T, V, S); ws = torch.empty(wsb // 4 + 1, device="cuda")
    def run():
        check(L.asr_ctc_fwd_bwd_f32(ptr(logits), ptr(targets), ptr(in_len), ptr(tgt_len), B, T, V, S, V - 1, ptr(nll), ptr(g), ptr(ws), wsb, sp()), "ctc")
    def timed(fn, n=5):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    for variant in (0, 1):
        lib.set_option("ctc_lattice_variant", variant)
        for chunks in (1, 2, 4, 8, 0):
            lib.set_option("ctc_chunks", chunks)
            print(dict(B=B, T=T, S=S, lattice_variant=variant, chunks=chunks, ctc_us=round(timed(run))), flush=True)
    lib.set_option("ctc_chunks", 0)
