#!/usr/bin/env python3
"""clock64 stamps of one CTA of the single-read attention forward (build with ASR_NVCC_EXTRA=-DASR_MHA_TRACE).
    ASR_NVCC_EXTRA=-DASR_MHA_TRACE python tools/mha_trace.py [variant ...]
Points per (tile, key block): 0 loop top, 1 scores ready, 2 scores in registers, 3 row maximum known,
4 rescale decided, 5 P tile free, 6 exponentials done, 7 P handed to the tensor core;
MMA thread: 8 P seen, 9 P V issued + committed, 10 S may be overwritten, 11 Q K^T issued + committed."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asr_b200
lib = asr_b200._lib; L = lib.lib(); ptr, sp, check = lib.ptr, lib.stream_ptr, lib.check
B, Ls, H = 16, 2048, 8
g = torch.Generator(device="cuda").manual_seed(5)
q, k, v = (torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
out = torch.empty_like(q); lse = torch.empty(B, H, Ls, device="cuda")
dll = ctypes.CDLL(L._name)
buf = (ctypes.c_longlong * 640)()
for var in [int(x) for x in sys.argv[1:]] or [21]:
    lib.set_option("mha_variant", var)
    for _ in range(3):
        check(L.asr_mha_fwd_bf16(ptr(q), ptr(k), ptr(v), None, None, 0, B, H, Ls, Ls, 64, 0.125, ptr(out), ptr(lse), sp()), "fwd")
    torch.cuda.synchronize()
    assert dll.asr_debug_mha_trace(buf) == 0
    t0 = min(buf[0], buf[320])
    print("variant %d (cycles since the first stamp; columns = points 0..7)" % var)
    for j in range(16):
        for t in range(2):
            row = [buf[(t * 16 + j) * 20 + i] - t0 for i in range(20)]
            print("  tile %d block %2d: " % (t, j) + " ".join("%7d" % x for x in row[:8]) + "   | wait S %5d ld %5d max %5d pvwait %5d exp %5d | mma: PV go %6d issued %6d  S go %6d issued %6d" % (
                row[1] - row[0], row[2] - row[1], row[3] - row[2], row[5] - row[4], row[6] - row[5], row[8], row[9], row[10], row[11]) + " | P done per warp " + " ".join("%6d" % x for x in row[12:16]) + " | mma at PV %6d V ok %6d" % (row[16], row[17]))
