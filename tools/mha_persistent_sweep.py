import os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
lib = bench.pkg("_lib"); L = lib.lib(); p, sp = lib.ptr, lib.stream_ptr
g = torch.Generator(device="cuda").manual_seed(5)
for (B, Ls, H) in ((90, 167, 8), (128, 256, 8), (64, 512, 8), (43, 768, 8), (32, 1024, 8), (21, 1536, 8), (16, 2048, 8), (8, 4096, 8)):
    q, k, v = (torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
    out = torch.empty_like(q); lse = torch.empty(B, H, Ls, device="cuda")
    res = []
    for causal in (0, 1):
        for var in (21, 40):
            lib.set_option("mha_variant", var)
            fwd = lambda: lib.check(L.asr_mha_fwd_bf16(p(q), p(k), p(v), None, None, causal, B, H, Ls, Ls, 64, 0.125, p(out), p(lse), sp()), "fwd")
            ms = min(bench.cuda_time(fwd, 10, warm=3) for _ in range(3))
            res.append("%s v%d %.4f ms" % ("causal" if causal else "full", var, ms))
    lib.set_option("mha_variant", 0)
    print("B=%d L=%d: " % (B, Ls) + " | ".join(res), flush=True)
