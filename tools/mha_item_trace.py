#!/usr/bin/env python3
"""Item-level clock64 stamps of CTA 0 of the persistent attention forward (build with ASR_NVCC_EXTRA=-DASR_MHA_TRACE).
Per (tile, item): 0 item start, 1 S(0) in registers, 2 token taken for block 0, 3 block 0 handed to the tensor core,
4 deferred epilogue of the previous item done, 6 / 7 scores of block 1 / 2 in registers, 5 last block handed over."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asr_b200
lib = asr_b200._lib; L = lib.lib(); ptr, sp, check = lib.ptr, lib.stream_ptr, lib.check
var = int(sys.argv[1]) if len(sys.argv) > 1 else 40
Ls = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
B, H = 16 * 2048 // Ls, 8
g = torch.Generator(device="cuda").manual_seed(5)
q, k, v = (torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
out = torch.empty_like(q); lse = torch.empty(B, H, Ls, device="cuda")
dll = ctypes.CDLL(L._name)
buf = (ctypes.c_longlong * 1024)()
lib.set_option("mha_variant", var)
for _ in range(3):
    check(L.asr_mha_fwd_bf16(ptr(q), ptr(k), ptr(v), None, None, 0, B, H, Ls, Ls, 64, 0.125, ptr(out), ptr(lse), sp()), "fwd")
torch.cuda.synchronize()
assert dll.asr_debug_mha_trace(buf) == 0
t0 = min(buf[0], buf[320])
print("variant %d L=%d: cycles since the first stamp" % (var, Ls))
for kk in range(8):
    for t in range(2):
        row = [buf[(t * 16 + kk) * 20 + i] - t0 for i in range(14)]
        print("  item %d tile %d: start %7d  S0 in regs %7d  token %7d  blk0 done %7d  epilogue done %7d  S1 %7d  S2 %7d  last blk done %7d" % (
            kk, t, row[0], row[1], row[2], row[3], row[4], row[6], row[7], row[5]) +
              "  | before s_full wait %7d after %7d | issuer: S0 of this item go %7d issued %7d, PV0 issued %7d, PV last issued %7d" % (row[8], row[9], row[10], row[11], row[12], row[13]))
