#!/usr/bin/env python3
"""How much does the lattice kernel (K2) slow down next to (a) a device-to-device copy that saturates HBM,
(b) an ALU-only spin kernel, (c) the row kernel K1 of another slice?  Diagnostics for the sliced pipeline."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200
from helpers import make_ctc_inputs
lib = asr_b200._lib; L = lib.lib(); ptr, check = lib.ptr, lib.check
V = 4233
B, T, S = 64, 1600, 80
lib.set_option("ctc_lattice_variant", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
def mk(seed):
    logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=seed)
    tgt_len = targets.ne(0).sum(1).to(torch.int32)
    nll = torch.empty(B, device="cuda"); g = torch.empty_like(logits)
    wsb = L.asr_ctc_workspace_bytes(B, T, V, S); ws = torch.empty(wsb // 4 + 1, device="cuda")
    return dict(logits=logits, targets=targets, in_len=in_len, tgt_len=tgt_len, nll=nll, g=g, ws=ws, wsb=wsb)
A, Bb = mk(1), mk(2)
def stage(d, stages, stream):
    check(L.asr_ctc_stages_f32(ptr(d["logits"]), ptr(d["targets"]), ptr(d["in_len"]), ptr(d["tgt_len"]), B, T, V, S, V - 1,
                               ptr(d["nll"]), ptr(d["g"]), ptr(d["ws"]), d["wsb"], stages, stream.cuda_stream), "ctc")
hi = torch.cuda.Stream(priority=-1); lo = torch.cuda.Stream(priority=0)
src = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); dst = torch.empty_like(src)
x = torch.randn(148 * 2048 * 4, device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
def k2_time(background):
    out = []
    for rep in range(3):
        stage(A, 1, hi)      # refill the table (K2 overwrites it)
        torch.cuda.synchronize()
        s0 = ev(); s0.record(lo)
        background(lo)
        hi.wait_event(s0)
        e0, e1 = ev(), ev()
        e0.record(hi); stage(A, 2, hi); e1.record(hi)
        eb = ev(); eb.record(lo)
        torch.cuda.synchronize()
        out.append((round(e0.elapsed_time(e1) * 1e3), round(s0.elapsed_time(eb) * 1e3)))
    return out[-1]
def none(st): pass
def copies(st):
    with torch.cuda.stream(st):
        for _ in range(6): dst.copy_(src, non_blocking=True)
def alu(st):
    with torch.cuda.stream(st):
        y = x
        for _ in range(60): y = torch.sin(y) * 1.0001 + 0.1   # elementwise, L2-resident: mostly issue slots
def k1(st):
    for _ in range(2): stage(Bb, 1, st)
for name, bg in (("alone", none), ("d2d copies", copies), ("elementwise", alu), ("K1 other slice", k1)):
    k2, bgt = k2_time(bg)
    print("K2 next to %-16s: K2 %5d us   (background ran %5d us)" % (name, k2, bgt), flush=True)
