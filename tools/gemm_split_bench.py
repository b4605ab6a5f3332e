#!/usr/bin/env python3
"""Split-K flavours of the fp32 / bf16 GEMMs at the model's shapes: inside a cluster (gemm_split_mode 0), workspace +
reduce kernel (1), no split (gemm_split_k 1).  CUDA-graph replays of 20 calls, so launch gaps are the graph's:
    python tools/gemm_split_bench.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops")
lib = importlib.import_module("end-to-end_asr_pytorch_b200._lib")


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
shapes = [("fwd  x W^T", 1344, 512, 512, False, False), ("fwd  w_1", 1344, 2048, 512, False, False), ("fwd  w_2", 1344, 512, 2048, False, False),
          ("dX   g W", 1344, 512, 512, False, True), ("dX   w_1", 1344, 512, 2048, False, True), ("dX   w_2", 1344, 2048, 512, False, True),
          ("dW   g^T x", 512, 512, 1344, True, True), ("dW   w_1", 2048, 512, 1344, True, True), ("dW   w_2", 512, 2048, 1344, True, True),
          ("dec  x W^T", 896, 512, 512, False, False), ("dec  w_1", 896, 2048, 512, False, False), ("dec dW", 512, 512, 896, True, True),
          ("enc  x W^T", 15030, 512, 512, False, False), ("enc  w_1", 15030, 2048, 512, False, False), ("enc  w_2", 15030, 512, 2048, False, False),
          ("enc  dX", 15030, 512, 512, False, True), ("enc  dW", 512, 512, 15030, True, True), ("enc dW w_1", 2048, 512, 15030, True, True),
          ("fc 4233", 896, 4233, 512, False, False), ("fc dW", 4233, 512, 896, True, True), ("fc dX", 896, 512, 4233, False, True)]
sweep = "--sweep" in sys.argv
for dt in (torch.float32, torch.bfloat16):
    for name, M, N, K, amn, bmn in shapes:
        a = torch.randn((K, M) if amn else (M, K), device="cuda", generator=gen).to(dt)
        if amn and M % 8: a = torch.nn.functional.pad(a, (0, 8 - M % 8))[:, :M]
        b = torch.randn((K, N) if bmn else (N, K), device="cuda", generator=gen).to(dt)
        if not bmn and K % 8: b = torch.nn.functional.pad(b, (0, 8 - K % 8))[:, :K]
        if not amn and K % 8: a = torch.nn.functional.pad(a, (0, 8 - K % 8))[:, :K]
        bias = torch.randn(N, device="cuda", generator=gen)
        if dt == torch.float32:
            fn = lambda: ops.gemm_f32(a, b, a_mn_major=amn, b_mn_major=bmn, bias=bias)
        else:
            fn = lambda: ops.gemm_bf16(a, b, a_mn_major=amn, b_mn_major=bmn, bias=bias)
        res = []
        if sweep:
            for bn in (128, 256):
                if dt == torch.bfloat16:
                    lib.set_option("gemm_variant", 1 if bn == 128 else 2)
                else:
                    lib.set_option("gemm_f32_bn", bn)
                row = []
                for force in (1, 2, 3, 4, 5, 6, 8):
                    lib.set_option("gemm_split_k", force)
                    try:
                        row.append("%d:%5.1f" % (force, timed_graph(fn)))
                    except Exception as e:
                        row.append("%d: err" % force)
                res.append("bn%d " % bn + " ".join(row))
            lib.set_option("gemm_variant", 0); lib.set_option("gemm_f32_bn", 0); lib.set_option("gemm_split_k", 0)
            res.append("auto %5.1f" % timed_graph(fn))
        else:
            for label, mode, force, stage in (("cluster", 0, 0, 0), ("workspace", 1, 0, 0), ("no split", 0, 1, 0), ("no split, staged", 0, 1, 2),
                                              ("cluster, staged", 0, 0, 2)):
                lib.set_option("gemm_split_mode", mode); lib.set_option("gemm_split_k", force); lib.set_option("gemm_stage_out", stage)
                res.append("%s %6.2f" % (label, timed_graph(fn)))
            lib.set_option("gemm_split_mode", 0); lib.set_option("gemm_split_k", 0); lib.set_option("gemm_stage_out", 0)
            if dt == torch.bfloat16:
                for label, pers, var in (("one tile per CTA bn128", 1, 1), ("persistent bn128", 2, 1), ("one tile bn256", 1, 2), ("persistent bn256", 2, 2), ("auto", 0, 0)):
                    lib.set_option("gemm_persistent", pers); lib.set_option("gemm_variant", var); lib.set_option("gemm_split_k", 1 if pers else 0)
                    res.append("%s %6.2f" % (label, timed_graph(fn)))
                lib.set_option("gemm_persistent", 0); lib.set_option("gemm_variant", 0); lib.set_option("gemm_split_k", 0)
        print("%-8s %-11s M=%5d N=%5d K=%5d  us/call: %s" % (str(dt).split(".")[1], name, M, N, K, " | ".join(res)), flush=True)
