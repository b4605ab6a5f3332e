import os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
lib = bench.pkg("_lib"); L = lib.lib(); p, sp = lib.ptr, lib.stream_ptr
g = torch.Generator(device="cuda").manual_seed(5)
for (B, Ls, H) in ((90, 167, 8), (64, 512, 8), (32, 1024, 8), (16, 2048, 8), (8, 4096, 8)):
    q, k, v = (torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
    out = torch.empty_like(q); lse = torch.empty(B, H, Ls, device="cuda")
    res = []
    outs = {}
    for var in (3, 40):
        lib.set_option("mha_variant", var)
        fwd = lambda: lib.check(L.asr_mha_fwd_dropout_bf16(p(q), p(k), p(v), None, None, 0, B, H, Ls, Ls, 64, 0.125, 0.1, 1234, p(out), p(lse), sp()), "fwd")
        ms = min(bench.cuda_time(fwd, 10, warm=3) for _ in range(3))
        outs[var] = out.float().clone()
        res.append("v%d %.4f ms %.0f TFLOP/s" % (var, ms, 4.0 * B * H * Ls * Ls * 64 / ms / 1e9))
    lib.set_option("mha_variant", 0)
    err = (outs[3] - outs[40]).abs().max().item() / outs[3].abs().max().item()
    print("dropout 0.1 B=%d L=%d: " % (B, Ls) + " | ".join(res) + " | max diff between the kernels %.2e of scale" % err, flush=True)
