#!/usr/bin/env python3
"""Times the CTC stages separately (K1, K1+K2 forward-only, K1+K2 with gradient) for both lattice variants."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200
from helpers import make_ctc_inputs
lib = asr_b200._lib; L = lib.lib(); ptr, sp, check = lib.ptr, lib.stream_ptr, lib.check
V = 4233
for (B, T, S) in [(32, 200, 10), (64, 400, 20), (128, 800, 40), (32, 1600, 80), (256, 1600, 80)]:
    logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=1236)
    tgt_len = targets.ne(0).sum(1).to(torch.int32)
    nll = torch.empty(B, device="cuda"); g = torch.empty_like(logits)
    wsb = L.asr_ctc_workspace_bytes(B, T, V, S); ws = torch.empty(wsb // 4 + 1, device="cuda")
    def run(stages, grad=True):
        check(L.asr_ctc_stages_f32(ptr(logits), ptr(targets), ptr(in_len), ptr(tgt_len), B, T, V, S, V - 1, ptr(nll), ptr(g) if grad else None, ptr(ws), wsb, stages, sp()), "ctc")
    def timed(fn, n=4):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    for variant in (2, 3):
        lib.set_option("ctc_lattice_variant", variant)
        k1 = timed(lambda: run(1))
        k12 = timed(lambda: run(3))
        k12f = timed(lambda: run(3, grad=False))
        k1f = timed(lambda: run(1, grad=False))
        whole = timed(lambda: check(L.asr_ctc_fwd_bwd_f32(ptr(logits), ptr(targets), ptr(in_len), ptr(tgt_len), B, T, V, S, V - 1, ptr(nll), ptr(g), ptr(ws), wsb, sp()), "ctc"))
        nbytes = 8 * V * int(in_len.sum().item())
        print(dict(B=B, T=T, S=S, lattice_variant=variant, K1_us=round(k1), K2_grad_us=round(k12 - k1), K2_fwd_only_us=round(k12f - k1f),
                   whole_call_us=round(whole), whole_GBps=round(nbytes / whole / 1e3)), flush=True)
    lib.set_option("ctc_lattice_variant", 0)
