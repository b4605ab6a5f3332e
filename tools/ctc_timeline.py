#!/usr/bin/env python3
"""Timeline of the batch-sliced CTC pipeline, rebuilt from Python with the staged entry point so each
kernel's start/end can be bracketed by events (diagnostics only; normalisation is per slice here)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200
from helpers import make_ctc_inputs
lib = asr_b200._lib; L = lib.lib(); ptr, check = lib.ptr, lib.check
V = 4233
B, T, S = 256, 1600, 80
nchunk = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lib.set_option("ctc_lattice_variant", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=1236)
tgt_len = targets.ne(0).sum(1).to(torch.int32)
nll = torch.empty(B, device="cuda"); g = torch.empty_like(logits)
per = B // nchunk
wsb = L.asr_ctc_workspace_bytes(per, T, V, S)
wss = [torch.empty(wsb // 4 + 1, device="cuda") for _ in range(nchunk)]
lo, hi = 0, -5
rows = torch.cuda.Stream(priority=0)
lats = [torch.cuda.Stream(priority=-1) for _ in range(nchunk)]
def stage(c, stages, stream):
    b0 = c * per
    check(L.asr_ctc_stages_f32(ptr(logits[b0:b0 + per]), ptr(targets[b0:b0 + per]), ptr(in_len[b0:b0 + per]), ptr(tgt_len[b0:b0 + per]),
                               per, T, V, S, V - 1, ptr(nll[b0:b0 + per]), ptr(g[b0:b0 + per]), ptr(wss[c]), wsb, stages, stream.cuda_stream), "ctc")
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    torch.cuda.synchronize()
    start = ev(); start.record(torch.cuda.current_stream())
    rows.wait_event(start)
    marks = []
    for c in range(nchunk):
        stage(c, 1, rows)
        e1 = ev(); e1.record(rows)
        lats[c].wait_event(e1)
        stage(c, 2, lats[c])
        e2 = ev(); e2.record(lats[c])
        stage(c, 4, lats[c])
        e3 = ev(); e3.record(lats[c])
        marks.append((e1, e2, e3))
    torch.cuda.synchronize()
    if rep == 2:
        for c, (e1, e2, e3) in enumerate(marks):
            print("chunk %d: K1 end %.0f us, K2 end %.0f us, K3 end %.0f us" % (c, start.elapsed_time(e1) * 1e3, start.elapsed_time(e2) * 1e3, start.elapsed_time(e3) * 1e3))
