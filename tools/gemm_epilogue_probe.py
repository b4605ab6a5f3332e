import importlib, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops"); lib = importlib.import_module("end-to-end_asr_pytorch_b200._lib")
def timed_graph(fn, calls=20, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(calls): fn()
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps): g.replay()
        e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / calls * 1e3
gen = torch.Generator(device="cuda").manual_seed(1)
for (M, N, K) in ((15030, 512, 512), (15030, 2048, 512)):
    a = torch.randn(M, K, device="cuda", generator=gen).bfloat16(); b = torch.randn(N, K, device="cuda", generator=gen).bfloat16()
    bias = torch.randn(N, device="cuda", generator=gen)
    for dbg in (0, 1):
        lib.set_option("gemm_debug", dbg)
        print(M, N, K, "debug", dbg, "with bias %.2f us" % timed_graph(lambda: ops.gemm_bf16(a, b, bias=bias)), "no bias %.2f us" % timed_graph(lambda: ops.gemm_bf16(a, b)), flush=True)
    lib.set_option("gemm_debug", 0)
