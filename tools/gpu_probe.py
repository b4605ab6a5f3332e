#!/usr/bin/env python3
"""Quick on-GPU timing probe of the hot-path kernels (CUDA events, L2 flushed
between iterations).  Writes gpurun_out/probe_<tag>.json.  Not a benchmark of
record - bench.py is; this is the builder's iteration tool."""
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

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=5, warmup=2):
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


def probe_cif(res, B=64, T=3000, H=512, n=300):
    hidden, alphas = make_cif_inputs(B, T, H, n, seed=1238)
    L = ops.cif_label_len(alphas)
    fwd_bytes = 4 * (B * T * H + B * T) + 4 * B * L * H
    bwd_bytes = 4 * (2 * B * T * H + B * L * H + 4 * B * T)
    for variant, width, stages in [(1, 128, 0), (1, 64, 0), (1, 32, 0), (2, 128, 6), (2, 64, 6), (2, 32, 6),
                                   (2, 64, 3), (2, 64, 12), (2, 128, 3), (2, 128, 10), (2, 32, 12)]:
        lib.set_option("cif_fwd_variant", variant)
        lib.set_option("cif_fwd_width", width)
        lib.set_option("cif_fwd_stages", stages)
        med, best = timeit(lambda: ops.cif(hidden, alphas, 0.95, L=L, check_overflow=False))
        res.append({"kernel": "cif_fwd", "variant": variant, "width": width, "stages": stages, "B": B, "T": T, "H": H,
                    "L": L, "ms": med * 1e3, "best_ms": best * 1e3, "GBps": fwd_bytes / med / 1e9})
        print(res[-1], flush=True)
    for k in ("cif_fwd_variant", "cif_fwd_width", "cif_fwd_stages"):
        lib.set_option(k, 0)
    h = hidden.clone().requires_grad_(True)
    a = alphas.clone().requires_grad_(True)
    out = ops.cif(h, a, 0.95, L=L, check_overflow=False)
    g_out = torch.randn_like(out)
    med, best = timeit(lambda: torch.autograd.grad(out, [h, a], g_out, retain_graph=True))
    res.append({"kernel": "cif_bwd", "B": B, "T": T, "H": H, "L": L, "ms": med * 1e3, "best_ms": best * 1e3,
                "GBps": bwd_bytes / med / 1e9})
    print(res[-1], flush=True)


def probe_ctc(res, shapes):
    for (B, T, S) in shapes:
        V = 4233
        logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=1236)
        valid = int(in_len.sum().item())
        alg = 8 * V * valid
        lg = logits.requires_grad_(True)

        def step():
            loss = ops.ctc_loss(lg, in_len, targets)
            loss.backward()
            lg.grad = None
        med, best = timeit(step, iters=4, warmup=2)
        res.append({"kernel": "ctc_fwd_bwd", "B": B, "T": T, "S": S, "V": V, "valid_frames": valid, "ms": med * 1e3,
                    "best_ms": best * 1e3, "GBps_alg": alg / med / 1e9})
        print(res[-1], flush=True)
        with torch.no_grad():
            med, best = timeit(lambda: ops.ctc_loss(logits.detach(), in_len, targets), iters=4, warmup=2)
        res.append({"kernel": "ctc_fwd_only", "B": B, "T": T, "S": S, "V": V, "ms": med * 1e3,
                    "GBps_alg": 4 * V * valid / med / 1e9})
        print(res[-1], flush=True)
        # torch's own path (what the reference calls) for comparison
        import torch.nn.functional as F
        lg2 = logits.detach().clone().requires_grad_(True)
        tl = targets.ne(0).int().sum(1)

        def ref_step():
            lp = F.log_softmax(lg2, dim=-1).transpose(0, 1)
            loss = F.ctc_loss(lp, targets, in_len, tl, blank=V - 1)
            loss.backward()
            lg2.grad = None
        med, best = timeit(ref_step, iters=3, warmup=1)
        res.append({"kernel": "torch_ctc_fwd_bwd", "B": B, "T": T, "S": S, "ms": med * 1e3, "GBps_alg": alg / med / 1e9})
        print(res[-1], flush=True)
        del logits, lg, lg2
        torch.cuda.empty_cache()


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    res = []
    t0 = time.time()
    print(torch.cuda.get_device_name(0), os.cpu_count(), flush=True)
    # memcpy ceiling on this box
    a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    b = torch.empty_like(a)
    med, best = timeit(lambda: b.copy_(a))
    res.append({"kernel": "torch_copy_1GiB", "ms": med * 1e3, "GBps": 2 * (1 << 30) / med / 1e9})
    print(res[-1], flush=True)
    del a, b
    probe_cif(res)
    probe_ctc(res, [(32, 200, 10), (64, 400, 20), (128, 800, 40), (32, 1600, 80), (256, 1600, 80)])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_%s.json" % tag), "w") as f:
        json.dump(res, f, indent=1)
    print("probe done in %.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
