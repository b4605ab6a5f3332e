#!/usr/bin/env python3
"""Quick on-GPU timing probe of the hot-path kernels through the C ABI with
preallocated buffers (CUDA events on the launching stream, L2 flushed between
iterations).  Writes gpurun_out/probe_<tag>.json.  Not a benchmark of record -
bench.py is; this is the builder's iteration tool.

    python tools/gpu_probe.py <tag> [cif] [ctc] [once]
`once` runs every kernel exactly once (for an ncu launch list)."""
import ctypes
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200  # noqa: E402
from helpers import make_cif_inputs, make_ctc_inputs  # noqa: E402

ops = asr_b200.ops
lib = asr_b200._lib
L = lib.lib()
ptr, sp, check = lib.ptr, lib.stream_ptr, lib.check

ONCE = "once" in sys.argv
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=5, warmup=2):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return 0.0, 0.0
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


class CifBuffers:
    def __init__(self, B, T, H, n, seed=1238):
        self.B, self.T, self.H = B, T, H
        self.hidden, self.alphas = make_cif_inputs(B, T, H, n, seed=seed)
        self.L = ops.cif_label_len(self.alphas)
        d = "cuda"
        self.out = torch.empty(B, self.L, H, device=d)
        self.fire_t = torch.empty(B, self.L, dtype=torch.int32, device=d)
        self.n_fired = torch.empty(B, dtype=torch.int32, device=d)
        self.cur = torch.empty(B, T, device=d)
        self.rem = torch.empty(B, T, device=d)
        self.sched = torch.empty(B, T, dtype=torch.int32, device=d)
        self.asum = torch.empty(B, device=d)
        self.g_out = torch.randn(B, self.L, H, device=d)
        self.g_hidden = torch.empty(B, T, H, device=d)
        self.g_alpha = torch.empty(B, T, device=d)
        self.ws = torch.empty(B * T, device=d)
        self.fwd_bytes = 4 * (B * T * H + B * T) + 4 * B * self.L * H
        self.bwd_bytes = 4 * (2 * B * T * H + B * self.L * H + 4 * B * T)

    def fwd(self):
        check(L.asr_cif_fwd_f32(ptr(self.hidden), ptr(self.alphas), 0.95, self.B, self.T, self.H, self.L,
                                ptr(self.out), ptr(self.fire_t), ptr(self.n_fired), ptr(self.cur), ptr(self.rem),
                                ptr(self.sched), ptr(self.asum), None, None, sp()), "cif_fwd")

    def bwd(self):
        check(L.asr_cif_bwd_f32(ptr(self.hidden), ptr(self.g_out), ptr(self.n_fired), ptr(self.cur), ptr(self.rem),
                                ptr(self.sched), self.B, self.T, self.H, self.L, ptr(self.g_hidden),
                                ptr(self.g_alpha), ptr(self.ws), self.ws.numel() * 4, sp()), "cif_bwd")


def probe_cif(res, B=64, T=3000, H=512, n=300):
    c = CifBuffers(B, T, H, n)
    combos = [(2, 0, 0, 0), (3, 32, 4, 4)] if ONCE else [(2, 64, 3, 0), (2, 64, 2, 0), (2, 64, 4, 0), (2, 64, 6, 0), (2, 128, 2, 0), (2, 128, 3, 0), (2, 32, 4, 0), (2, 32, 6, 0), (3, 128, 4, 1), (3, 0, 0, 0), (0, 0, 0, 0)]
    for variant, width, stages, nw in combos:
        lib.set_option("cif_fwd_rows", nw)
        lib.set_option("cif_fwd_variant", variant)
        lib.set_option("cif_fwd_width", width)
        lib.set_option("cif_fwd_stages", stages)
        med, best = timeit(c.fwd)
        res.append({"kernel": "cif_fwd", "variant": variant, "width": width, "stages": stages, "nw": nw, "B": B, "T": T, "H": H,
                    "L": c.L, "us": med * 1e6, "best_us": best * 1e6, "GBps": c.fwd_bytes / max(med, 1e-9) / 1e9})
        print(res[-1], flush=True)
    for k in ("cif_fwd_variant", "cif_fwd_width", "cif_fwd_stages", "cif_fwd_rows"):
        lib.set_option(k, 0)
    c.fwd()
    med, best = timeit(c.bwd)
    res.append({"kernel": "cif_bwd", "B": B, "T": T, "H": H, "L": c.L, "us": med * 1e6, "best_us": best * 1e6,
                "GBps": c.bwd_bytes / max(med, 1e-9) / 1e9})
    print(res[-1], flush=True)


def probe_ctc(res, shapes):
    V = 4233
    for (B, T, S) in shapes:
        logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=1236)
        tgt_len = targets.ne(0).sum(1).to(torch.int32)
        valid = int(in_len.sum().item())
        alg = 8 * V * valid
        nll = torch.empty(B, device="cuda")
        g = torch.empty_like(logits)
        wsb = L.asr_ctc_workspace_bytes(B, T, V, S)
        ws = torch.empty(wsb // 4 + 1, device="cuda")

        def run(grad):
            check(L.asr_ctc_fwd_bwd_f32(ptr(logits), ptr(targets), ptr(in_len), ptr(tgt_len), B, T, V, S, V - 1,
                                        ptr(nll), ptr(g) if grad else None, ptr(ws), wsb, sp()), "ctc")
        lib.set_option("ctc_lattice_variant", 1)
        med, best = timeit(lambda: run(True), iters=4)
        res.append({"kernel": "ctc_fwd_bwd_one_warp_lattice", "B": B, "T": T, "S": S, "us": med * 1e6,
                    "GBps_alg": alg / max(med, 1e-9) / 1e9})
        print(res[-1], flush=True)
        lib.set_option("ctc_lattice_variant", 0)
        med, best = timeit(lambda: run(True), iters=4)
        res.append({"kernel": "ctc_fwd_bwd", "B": B, "T": T, "S": S, "V": V, "valid_frames": valid, "us": med * 1e6,
                    "best_us": best * 1e6, "GBps_alg": alg / max(med, 1e-9) / 1e9})
        print(res[-1], flush=True)
        med, best = timeit(lambda: run(False), iters=4)
        res.append({"kernel": "ctc_fwd_only", "B": B, "T": T, "S": S, "V": V, "us": med * 1e6,
                    "GBps_alg": 4 * V * valid / max(med, 1e-9) / 1e9})
        print(res[-1], flush=True)
        del logits, g
        torch.cuda.empty_cache()


def probe_mha(res):
    shapes = [(64, 167, 8), (64, 512, 8), (32, 1024, 8), (16, 2048, 8), (8, 4096, 8)]
    if ONCE:
        shapes = [(16, 2048, 8)]
    for (B, Ls, H) in shapes:
        for causal in (False, True):
            g = torch.Generator(device="cuda").manual_seed(5)
            q = torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16)
            k = torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16)
            v = torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16)
            do = torch.randn(B, Ls, H, 64, device="cuda", generator=g).to(torch.bfloat16)
            out = torch.empty_like(q)
            lse = torch.empty(B, H, Ls, device="cuda")
            gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            wsb = L.asr_mha_bwd_workspace_bytes(B, H, Ls, Ls, 64)
            ws = torch.empty(wsb // 4 + 1, device="cuda")
            scale = 0.125

            def fwd():
                check(L.asr_mha_fwd_bf16(ptr(q), ptr(k), ptr(v), None, None, int(causal), B, H, Ls, Ls, 64, scale,
                                         ptr(out), ptr(lse), sp()), "mha_fwd")

            def bwd():
                check(L.asr_mha_bwd_bf16(ptr(q), ptr(k), ptr(v), ptr(out), ptr(do), ptr(lse), None, None, int(causal),
                                         B, H, Ls, Ls, 64, scale, ptr(gq), ptr(gk), ptr(gv), ptr(ws), wsb, sp()), "mha_bwd")
            div = 2.0 if causal else 1.0
            fl = 4.0 * B * H * Ls * Ls * 64 / div
            for variant in (1, 2):
                lib.set_option("mha_variant", variant)
                med, best = timeit(fwd)
                res.append({"kernel": "mha_fwd", "variant": variant, "B": B, "L": Ls, "H": H, "causal": causal,
                            "us": med * 1e6, "TFLOPs": fl / max(med, 1e-9) / 1e12})
                print(res[-1], flush=True)
            lib.set_option("mha_variant", 0)
            med, best = timeit(bwd)
            fl = 10.0 * B * H * Ls * Ls * 64 / div
            res.append({"kernel": "mha_bwd", "B": B, "L": Ls, "H": H, "causal": causal, "us": med * 1e6,
                        "TFLOPs": fl / max(med, 1e-9) / 1e12})
            print(res[-1], flush=True)
            if not ONCE and not causal:
                import torch.nn.functional as F
                qh, kh, vh = (t.permute(0, 2, 1, 3).contiguous() for t in (q, k, v))
                med, best = timeit(lambda: F.scaled_dot_product_attention(qh, kh, vh))
                res.append({"kernel": "torch_sdpa_fwd", "B": B, "L": Ls, "us": med * 1e6,
                            "TFLOPs": 4.0 * B * H * Ls * Ls * 64 / med / 1e12})
                print(res[-1], flush=True)


def probe_cif_shapes(res):
    probe_cif(res)
    probe_cif(res, B=256, T=1600, H=512, n=80)
    probe_cif(res, B=128, T=1600, H=512, n=80)
    probe_cif(res, B=16, T=1000, H=256, n=60)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    res = []
    t0 = time.time()
    print(torch.cuda.get_device_name(0), os.cpu_count(), flush=True)
    if not ONCE:
        a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
        b = torch.empty_like(a)
        med, best = timeit(lambda: b.copy_(a))
        res.append({"kernel": "torch_copy_1GiB", "us": med * 1e6, "GBps": 2 * (1 << 30) / med / 1e9})
        print(res[-1], flush=True)
        del a, b
    sel = {"cif", "ctc", "mha"} & set(sys.argv)
    if "mha" in sel or not sel:
        probe_mha(res)
    if "cif" in sel or not sel:
        probe_cif_shapes(res)
    if "ctc" in sel or not sel:
        shapes = [(32, 1600, 80), (256, 1600, 80)] if ONCE else \
            [(32, 200, 10), (64, 400, 20), (128, 800, 40), (32, 1600, 80), (256, 1600, 80)]
        probe_ctc(res, shapes)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_%s.json" % tag), "w") as f:
        json.dump(res, f, indent=1)
    print("probe done in %.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
