#!/usr/bin/env python3
"""Per-CTA timeline of the two-tile attention forward (build with ASR_NVCC_EXTRA="-DASR_MHA_TRACE -DASR_MHA_TRACE_LIGHT"):
how long a CTA lives, how much of that is the key-block loop, and how long an SM waits between two CTAs.
    python tools/mha_cta_timeline.py [variant] [L]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asr_b200
lib = asr_b200._lib; L = lib.lib(); ptr, sp, check = lib.ptr, lib.stream_ptr, lib.check
var = int(sys.argv[1]) if len(sys.argv) > 1 else 21
Ls = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
B, H = 16 * 2048 // Ls, 8
g = torch.Generator(device="cuda").manual_seed(5)
q, k, v = (torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
out = torch.empty_like(q); lse = torch.empty(B, H, Ls, device="cuda")
dll = ctypes.CDLL(L._name)
n = B * H * ((Ls + 255) // 256)
buf = (ctypes.c_longlong * (2048 * 5))()
lib.set_option("mha_variant", var)
for _ in range(3):
    check(L.asr_mha_fwd_bf16(ptr(q), ptr(k), ptr(v), None, None, 0, B, H, Ls, Ls, 64, 0.125, ptr(out), ptr(lse), sp()), "fwd")
torch.cuda.synchronize()
assert dll.asr_debug_mha_cta(buf) == 0
rows = [[buf[i * 5 + j] for j in range(5)] for i in range(min(n, 2048))]
by_sm = {}
for r in rows:
    by_sm.setdefault(r[0], []).append(r)
life, loop, pro, epi, gaps, first = [], [], [], [], [], []
for sm, rs in by_sm.items():
    rs.sort(key=lambda r: r[1])
    for i, r in enumerate(rs):
        life.append(r[4] - r[1]); pro.append(r[2] - r[1]); loop.append(r[3] - r[2]); epi.append(r[4] - r[3])
        if i > 0:
            gaps.append(r[1] - rs[i - 1][4])
med = lambda x: sorted(x)[len(x) // 2] if x else 0
print("variant %d L=%d: %d CTAs on %d SMs (%.2f per SM)" % (var, Ls, len(rows), len(by_sm), len(rows) / len(by_sm)))
print("median cycles: CTA life %d = entry->first block %d + key-block loop %d + last block->exit %d; gap between CTAs on an SM %d (min %d max %d)" % (
    med(life), med(pro), med(loop), med(epi), med(gaps), min(gaps) if gaps else 0, max(gaps) if gaps else 0))
span = [max(r[4] for r in rs) - min(r[1] for r in rs) for rs in by_sm.values()]
print("per-SM busy span: median %d max %d cycles; sum of CTA lives per SM median %d" % (med(span), max(span), med([sum(r[4] - r[1] for r in rs) for rs in by_sm.values()])))
