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
# CIF kernels as background (the bench's overlapped step puts them next to the last lattice)
Bc, Tc, Hc = 256, 1600, 512
from helpers import make_cif_inputs
hid, alp = make_cif_inputs(Bc, Tc, Hc, 80, seed=5)
Lc = asr_b200.ops.cif_label_len(alp)
cb = dict(out=torch.empty(Bc, Lc, Hc, device="cuda"), fire_t=torch.empty(Bc, Lc, dtype=torch.int32, device="cuda"),
          n_fired=torch.empty(Bc, dtype=torch.int32, device="cuda"), cur=torch.empty(Bc, Tc, device="cuda"),
          rem=torch.empty(Bc, Tc, device="cuda"), sched=torch.empty(Bc, Tc, dtype=torch.int32, device="cuda"),
          asum=torch.empty(Bc, device="cuda"), g_out=torch.randn(Bc, Lc, Hc, device="cuda"),
          g_hidden=torch.empty(Bc, Tc, Hc, device="cuda"), g_alpha=torch.empty(Bc, Tc, device="cuda"), ws=torch.empty(Bc * Tc, device="cuda"))
def cif_fwd_bg(variant):
    def run(st):
        lib.set_option("cif_fwd_variant", variant)
        for _ in range(4):
            check(L.asr_cif_fwd_f32(ptr(hid), ptr(alp), 0.95, Bc, Tc, Hc, Lc, ptr(cb["out"]), ptr(cb["fire_t"]), ptr(cb["n_fired"]),
                                    ptr(cb["cur"]), ptr(cb["rem"]), ptr(cb["sched"]), ptr(cb["asum"]), None, None, st.cuda_stream), "cif_fwd")
        lib.set_option("cif_fwd_variant", 0)
    return run
def cif_bwd_bg(st):
    for _ in range(3):
        check(L.asr_cif_bwd_f32(ptr(hid), ptr(cb["g_out"]), ptr(cb["n_fired"]), ptr(cb["cur"]), ptr(cb["rem"]), ptr(cb["sched"]), Bc, Tc, Hc, Lc,
                                ptr(cb["g_hidden"]), ptr(cb["g_alpha"]), ptr(cb["ws"]), cb["ws"].numel() * 4, st.cuda_stream), "cif_bwd")
def k3(st):
    for _ in range(6): stage(Bb, 4, st)
cif_fwd_bg(3)(torch.cuda.current_stream()); torch.cuda.synchronize()
stage(Bb, 3, lo); torch.cuda.synchronize()
for name, bg in (("alone", none), ("d2d copies", copies), ("elementwise", alu), ("K1 other slice", k1), ("cif fwd v3 x4", cif_fwd_bg(3)),
                 ("cif fwd v2 x4", cif_fwd_bg(2)), ("cif bwd x3", cif_bwd_bg), ("K3 x6", k3)):
    k2, bgt = k2_time(bg)
    print("K2 next to %-16s: K2 %5d us   (background ran %5d us)" % (name, k2, bgt), flush=True)
