#!/usr/bin/env python3
"""bench.py - throughput of the CIF + CTC training hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" is one pass of the hot path over one batch of synthetic
utterances, the work the reference does per training step around its encoder:

    CTC loss + gradient on the logits [B,T,V]          (a4: K1 row pass, K2 lattice, K3 sparse update)
    CIF forward on the encoder frames [B,T,H] + quantity term  (a1-a3)
    CIF backward                                         (a2')

Workload `cif_ctc_joint` = the largest shape of BASELINE config 2 (CTC sweep:
B=256, T=1600, S=80, V=4233) with the CIF layer of config 4 run on the same batch
(H=512).  Every rank processes its own batch (data parallel by utterance, weak
scaling); the hot-path kernels need no collective, only the scalar losses are
all-reduced for logging.

The JSON line carries:
  value        utterances/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the public Python API (ops.cif / ops.ctc_loss +
               autograd) with the inputs copied from pinned host memory and the losses
               read back every step
  roofline     the dominant kernel (ctc_rows) against the measured HBM copy peak,
               timed live with CUDA events inside the timed region
  cpu_baseline the reference's torch-CPU path (oracle/torch_port.py) on a bounded
               sample of the same workload, on this box's host cores
`--impl reference` times only that CPU path (rank 0), in the same JSON shape.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CIF/CTC train utts/sec"
UNIT = "utts/s"

WORKLOADS = {
    # name: per-GPU batch
    "cif_ctc_joint": dict(B=256, T=1600, S=80, V=4233, H=512),
    "cif_ctc_small": dict(B=32, T=200, S=10, V=4233, H=512),
    # BASELINE config 5: the whole CIF_Model (reference recipe defaults: LFR 4/3, 3 conv layers, 6+6
    # layers, d_model 512, 8 heads, d_inner 2048, V=4233) trained data-parallel with an NCCL gradient
    # all-reduce; 500 raw frames -> 167 LFR frames x 320 -> 21 encoder frames, 14 labels
    "train": dict(B=64, T=167, S=14, V=4233, H=512, D=320),
    "train_long": dict(B=32, T=534, S=45, V=4233, H=512, D=320),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cif_ctc_joint", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=4, help="utterances per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--serial", action="store_true", help="time the hot path one kernel at a time on one stream")
    ap.add_argument("--ctc-chunks", type=int, default=0, help="batch slices of the CTC pipeline (0 = library default)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="library tuning knob (asr_set_option), repeatable")
    ap.add_argument("--no-train-step", action="store_true", help="skip the full-model data-parallel step (config 5)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
# synthetic inputs, generated on the device that uses them
# ---------------------------------------------------------------------------------------
def make_inputs(w, device, seed):
    B, T, S, V, H = w["B"], w["T"], w["S"], w["V"], w["H"]
    g = torch.Generator(device=device).manual_seed(seed)
    logits = torch.randn(B, T, V, device=device, generator=g)
    hidden = torch.randn(B, T, H, device=device, generator=g)
    targets = torch.randint(1, V - 1, (B, S), device=device, generator=g)
    rep = torch.rand(B, S, device=device, generator=g) < 0.1
    for s in range(1, S):
        targets[:, s] = torch.where(rep[:, s], targets[:, s - 1], targets[:, s])
    in_len = torch.randint(int(0.6 * T), T + 1, (B,), device=device, generator=g).to(torch.int32)
    tgt_len = torch.randint(max(1, S // 2), S + 1, (B,), device=device, generator=g)
    targets = targets * (torch.arange(S, device=device)[None, :] < tgt_len[:, None]).long()
    # assigner output: sigmoid weights, zero on padded frames (attentionAssigner.py:34-40)
    alphas = torch.sigmoid(torch.randn(B, T, device=device, generator=g))
    alphas = alphas * (torch.arange(T, device=device)[None, :] < in_len[:, None]).float()
    noise = torch.rand(B, device=device, generator=g)
    return dict(logits=logits, hidden=hidden, targets=targets, in_len=in_len, tgt_len=tgt_len.to(torch.int32),
                alphas=alphas, noise=noise)


def scale_alphas(alphas, targets, noise):
    """cif_model.py:43-48 (torch glue, outside the kernels)."""
    _num = alphas.sum(-1)
    num = (targets > 0).float().sum(-1)
    return _num, num, alphas * ((num + noise - 0.5) / _num)[:, None]


# ---------------------------------------------------------------------------------------
# clocks: sampled with NVML while the timed region runs
# ---------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self.max_mhz = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append(mhz)
                for bit, name in self.REASONS.items():
                    if bits & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def start(self):
        if self.ok:
            self.th.start()

    def stop(self):
        self._stop.set()
        if self.ok:
            self.th.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
class HotPath:
    """Preallocated buffers + direct C-ABI calls (what the autograd wrappers do, minus the allocator)."""

    def __init__(self, w, inp, pkg):
        self.w, self.inp = w, inp
        self.lib = pkg._lib
        self.L = self.lib.lib()
        B, T, S, V, H = w["B"], w["T"], w["S"], w["V"], w["H"]
        dev = inp["logits"].device
        self.nll = torch.empty(B, device=dev)
        self.g_logits = torch.empty_like(inp["logits"])
        self.ws_bytes = self.L.asr_ctc_workspace_bytes(B, T, V, S)
        self.ws = torch.empty(self.ws_bytes // 4 + 1, device=dev)
        _num, num, self.alphas = scale_alphas(inp["alphas"], inp["targets"], inp["noise"])
        self.num = num.contiguous()
        self.alphas = self.alphas.contiguous()
        self.Lout = int(torch.round(self.alphas.sum(-1)).int().max().item())
        self.out = torch.empty(B, self.Lout, H, device=dev)
        self.fire_t = torch.empty(B, self.Lout, dtype=torch.int32, device=dev)
        self.n_fired = torch.empty(B, dtype=torch.int32, device=dev)
        self.cur = torch.empty(B, T, device=dev)
        self.rem = torch.empty(B, T, device=dev)
        self.sched = torch.empty(B, T, dtype=torch.int32, device=dev)
        self.asum = torch.empty(B, device=dev)
        self.qua = torch.empty(B, device=dev)
        self.g_out = torch.randn(B, self.Lout, H, device=dev)
        self.g_hidden = torch.empty_like(inp["hidden"])
        self.g_alpha = torch.empty(B, T, device=dev)
        self.cif_ws = torch.empty(B * T, device=dev)
        self.valid_frames = int(inp["in_len"].sum().item())
        self.n_kernels_per_step = 3 + 1 + 2
        self.cif_variant_overlapped = 3 if w["T"] >= 64 and w["H"] % 4 == 0 else 0

    def ctc(self, stages):
        w, i, p = self.w, self.inp, self.lib.ptr
        self.lib.check(self.L.asr_ctc_stages_f32(
            p(i["logits"]), p(i["targets"]), p(i["in_len"]), p(i["tgt_len"]), w["B"], w["T"], w["V"], w["S"],
            w["V"] - 1, p(self.nll), p(self.g_logits), p(self.ws), self.ws_bytes, stages, self.lib.stream_ptr()),
            "asr_ctc_stages_f32")

    def cif_fwd(self):
        w, i, p = self.w, self.inp, self.lib.ptr
        self.lib.check(self.L.asr_cif_fwd_f32(
            p(i["hidden"]), p(self.alphas), 0.95, w["B"], w["T"], w["H"], self.Lout, p(self.out), p(self.fire_t),
            p(self.n_fired), p(self.cur), p(self.rem), p(self.sched), p(self.asum), p(self.num), p(self.qua),
            self.lib.stream_ptr()), "asr_cif_fwd_f32")

    def cif_bwd(self):
        w, i, p = self.w, self.inp, self.lib.ptr
        self.lib.check(self.L.asr_cif_bwd_f32(
            p(i["hidden"]), p(self.g_out), p(self.n_fired), p(self.cur), p(self.rem), p(self.sched), w["B"], w["T"],
            w["H"], self.Lout, p(self.g_hidden), p(self.g_alpha), p(self.cif_ws), self.cif_ws.numel() * 4,
            self.lib.stream_ptr()), "asr_cif_bwd_f32")

    def step_overlapped(self, ev=None):
        """One hot-path pass the way the library is meant to be driven: one stream, the CTC call in
        its two phases with the CIF forward/backward pair queued in between, where it runs next to
        the last slice's latency-bound lattice.  The two halves share no data.
        ev = (before, after): CUDA events around the row kernels (all slices; they are the only
        work begin puts on this stream), i.e. the dominant kernel timed inside the timed region."""
        import ctypes
        w, i, p = self.w, self.inp, self.lib.ptr
        args = (p(i["logits"]), p(i["targets"]), p(i["in_len"]), p(i["tgt_len"]), w["B"], w["T"], w["V"], w["S"],
                w["V"] - 1, p(self.nll), p(self.g_logits), p(self.ws), self.ws_bytes, self.lib.stream_ptr())
        ticket = ctypes.c_int(0)
        if ev is not None:
            ev[0].record()
        self.lib.check(self.L.asr_ctc_begin_f32(*args, ctypes.byref(ticket)), "asr_ctc_begin_f32")
        if ev is not None:
            ev[1].record()
        # next to the lattices the warp-specialised CIF forward disturbs them least (measured: 2.79 ms
        # per step against 2.83 ms with the library's stand-alone choice, the one-warp TMA pipeline)
        self.lib.set_option("cif_fwd_variant", self.cif_variant_overlapped)
        self.cif_fwd()
        self.lib.set_option("cif_fwd_variant", 0)
        self.cif_bwd()
        self.lib.check(self.L.asr_ctc_finish_f32(*args, ticket.value), "asr_ctc_finish_f32")

    def step(self, ev=None):
        """One serial hot-path pass; ev = list of 6 CUDA events recorded between the stages."""
        def mark(k):
            if ev is not None:
                ev[k].record()
        mark(0)
        self.ctc(1)
        mark(1)
        self.ctc(2)
        mark(2)
        self.ctc(4)
        mark(3)
        self.cif_fwd()
        mark(4)
        self.cif_bwd()
        mark(5)

    def bytes_model(self):
        """ALGORITHMIC bytes per launch, stated in DESIGN.md."""
        w = self.w
        B, T, V, H, L = w["B"], w["T"], w["V"], w["H"], self.Lout
        return {
            "ctc_rows": 8 * V * self.valid_frames,                       # read logits once + write grad once, valid frames
            "ctc_total": 8 * V * self.valid_frames,
            "cif_fwd": 4 * (B * T * H + B * T) + 4 * B * L * H,
            "cif_bwd": 4 * (2 * B * T * H + B * L * H + 4 * B * T),
        }


def e2e_step(pkg, w, host, dev_buf, g_out):
    """Public-API step with host inputs: pinned H2D copies, ops.cif / ops.ctc_loss, autograd, D2H of the losses."""
    ops = pkg.ops
    for k in ("logits", "hidden", "alphas", "targets", "in_len", "noise"):
        dev_buf[k].copy_(host[k], non_blocking=True)
    logits = dev_buf["logits"].requires_grad_(True)
    hidden = dev_buf["hidden"].requires_grad_(True)
    alphas_raw = dev_buf["alphas"].requires_grad_(True)
    _num, num, alphas = scale_alphas(alphas_raw, dev_buf["targets"], dev_buf["noise"])
    fired = ops.cif(hidden, alphas, 0.95)
    qua = torch.pow(_num - num, 2).mean()
    ctc = ops.ctc_loss(logits, dev_buf["in_len"], dev_buf["targets"])
    total = ctc + 0.001 * qua + (fired * g_out[:, :fired.size(1)]).sum()
    total.backward()
    losses = torch.stack([ctc.detach(), qua.detach()]).cpu()          # D2H, synchronises
    for k in ("logits", "hidden", "alphas"):
        dev_buf[k].grad = None
        dev_buf[k].requires_grad_(False)
    return losses


def assigner_microbench(pkg, w, inp, device, iters=5):
    """CIF weight producer (assigner tail + scaling) on the bench shape: x = the encoder
    output of the step [B,T,H], ragged lengths.  HBM-bound: forward reads the valid rows of x once,
    backward reads them once more and writes g_x once."""
    lib = pkg._lib
    L = lib.lib()
    p, sp = lib.ptr, lib.stream_ptr
    B, T, D = w["B"], w["T"], w["H"]
    x = inp["hidden"]
    g = torch.Generator(device=device).manual_seed(6)
    wt = torch.randn(D, device=device, generator=g) * 0.05
    bias = torch.zeros(1, device=device)
    lens = inp["in_len"]
    noise = inp["tgt_len"].float() + 0.25
    alpha, a_raw = torch.empty(B, T, device=device), torch.empty(B, T, device=device)
    num = torch.empty(B, device=device)
    g_alpha, g_num = torch.randn(B, T, device=device, generator=g), torch.randn(B, device=device, generator=g)
    g_x, g_w, g_b = torch.empty_like(x), torch.empty(D, device=device), torch.empty(1, device=device)
    wsb = L.asr_cif_alpha_bwd_workspace_bytes(B, T, D)
    ws = torch.empty(wsb // 4 + 1, device=device)

    def fwd():
        lib.check(L.asr_cif_alpha_fwd_f32(p(x), p(wt), p(bias), p(lens), p(noise), B, T, D, p(alpha), p(a_raw), p(num), sp()), "alpha_fwd")

    def bwd():
        lib.check(L.asr_cif_alpha_bwd_f32(p(x), p(wt), p(lens), p(noise), p(a_raw), p(num), p(g_alpha), p(g_num), B, T, D,
                                          p(g_x), p(g_w), p(g_b), p(ws), wsb, sp()), "alpha_bwd")
    valid = int(lens.sum().item())
    res = {}
    for name, fn, nbytes in (("cif_alpha_fwd", fwd, 4 * valid * D + 12 * B * T),
                             ("cif_alpha_bwd", bwd, 4 * valid * D + 4 * B * T * D + 16 * B * T)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / (ms * 1e-3) / 1e9}
    return res


def spec_aug_microbench(pkg, device, iters=5):
    """SpecAugment on the device on a batch of LFR-stacked fbank of the bench size
    (B=256 x T=1600 x 320 bins, ragged, two bands + two spans per utterance).  HBM-bound: the batch is read
    once for the two means; only the masked cells are written."""
    import importlib
    ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops")
    B, T, V, R = 256, 1600, 320, 2
    g = torch.Generator(device=device).manual_seed(8)
    x = torch.randn(B, T, V, device=device, generator=g)
    lens = torch.randint(T // 2, T + 1, (B,), device=device, generator=g)
    fw = torch.randint(0, 27, (R, B), device=device, generator=g)
    f0 = (torch.rand(R, B, device=device, generator=g) * (V - fw)).long()
    tw = torch.randint(0, 40, (R, B), device=device, generator=g)
    t0 = (torch.rand(R, B, device=device, generator=g) * (lens[None] - tw)).long()
    masked = int((T * fw.sum() + V * tw.sum()).item())
    nbytes = 4 * B * T * V + 4 * masked

    def run():
        ops.spec_aug_apply(x, lens, f0, fw, t0, tw)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / (ms * 1e-3) / 1e9,
            "shape": "B=%d T=%d V=%d, %d bands + %d spans per utterance" % (B, T, V, R, R)}


def linear_microbench(pkg, device, iters=10):
    """Fused tcgen05 linear layers on the feed-forward block of the encoder at the bench batch
    (M = 64 x 1600 frames, d_model 512, d_inner 2048), next to torch's cuBLAS + eager epilogue on the same tensors."""
    import importlib
    ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops")
    F = torch.nn.functional
    M, d, di = 64 * 1600, 512, 2048
    g = torch.Generator(device=device).manual_seed(9)
    x = torch.randn(M, d, device=device, generator=g).bfloat16()
    w1 = (torch.randn(di, d, device=device, generator=g) * d ** -0.5).bfloat16()
    w2 = (torch.randn(d, di, device=device, generator=g) * di ** -0.5).bfloat16()
    b1, b2 = torch.randn(di, device=device, generator=g), torch.randn(d, device=device, generator=g)
    gam, bet = torch.ones(d, device=device), torch.zeros(d, device=device)
    h = ops.linear_act(x, w1, b1, relu=True)
    b1h, b2h, gamh, beth = b1.bfloat16(), b2.bfloat16(), gam.bfloat16(), bet.bfloat16()
    rows = [("linear_bias_relu", lambda: ops.linear_act(x, w1, b1, relu=True), lambda: torch.relu(F.linear(x, w1, b1h))),
            ("linear_residual_layernorm", lambda: ops.linear_residual_layernorm(h, w2, b2, x, gam, bet),
             lambda: F.layer_norm(F.linear(h, w2, b2h) + x, (d,), gamh, beth))]
    res = {}
    for name, ours, ref in rows:
        t = []
        for fn in (ours, ref):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1) / iters)
        flop = 2.0 * M * d * di
        res[name] = {"ms": t[0], "TFLOPs": flop / t[0] / 1e9, "torch_ms": t[1], "torch_TFLOPs": flop / t[1] / 1e9,
                     "shape": "M=%d K=%d N=%d bf16" % ((M, d, di) if name == "linear_bias_relu" else (M, di, d))}
    # fp32 in / fp32 out on the tensor cores (three TF32 products per tile) next to torch's fp32 GEMM (cuBLAS SIMT sgemm)
    xf, wf = x.float(), w1.float()
    t = []
    for fn in (lambda: ops.linear_f32(xf, wf, b1), lambda: F.linear(xf, wf, b1)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1) / iters)
    flop = 2.0 * M * d * di
    res["linear_f32_3xtf32"] = {"ms": t[0], "TFLOPs": flop / t[0] / 1e9, "torch_ms": t[1], "torch_TFLOPs": flop / t[1] / 1e9,
                                "shape": "M=%d K=%d N=%d f32" % (M, d, di)}
    return res


def attention_microbench(pkg, device, iters=5):
    """tcgen05 attention core, forward and backward, on the SURVEY 8(d) microbench shape
    (B*heads = 128, L = 2048, d = 64, no mask); CUDA events, inputs >> L2 per call not needed
    (compute bound).  Reported next to the hot-path kernels as the tensor-core row."""
    lib = pkg._lib
    L = lib.lib()
    p, sp = lib.ptr, lib.stream_ptr
    B, Ls, H = 16, 2048, 8
    g = torch.Generator(device=device).manual_seed(5)
    q, k, v, do = (torch.randn(B, Ls, H, 64, device=device, generator=g).to(torch.bfloat16) for _ in range(4))
    out = torch.empty_like(q)
    lse = torch.empty(B, H, Ls, device=device)
    gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    wsb = L.asr_mha_bwd_workspace_bytes(B, H, Ls, Ls, 64)
    ws = torch.empty(wsb // 4 + 1, device=device)

    def fwd():
        lib.check(L.asr_mha_fwd_bf16(p(q), p(k), p(v), None, None, 0, B, H, Ls, Ls, 64, 0.125, p(out), p(lse), sp()), "mha_fwd")

    def bwd():
        lib.check(L.asr_mha_bwd_bf16(p(q), p(k), p(v), p(out), p(do), p(lse), None, None, 0, B, H, Ls, Ls, 64, 0.125,
                                     p(gq), p(gk), p(gv), p(ws), wsb, sp()), "mha_bwd")
    res = {}
    for name, fn, flops in (("mha_fwd", fwd, 4.0 * B * H * Ls * Ls * 64), ("mha_bwd", bwd, 10.0 * B * H * Ls * Ls * 64)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[name] = {"ms": ms, "TFLOPs": flops / (ms * 1e-3) / 1e12, "shape": "B=%d L=%d heads=%d d=64 bf16" % (B, Ls, H)}
    return res


# ---------------------------------------------------------------------------------------
# workload "train": full CIF_Model step, data parallel with gradient all-reduce (config 5)
# ---------------------------------------------------------------------------------------
def _model_args(w):
    return argparse.Namespace(d_input=80, LFR_m=4, n_conv_layers=3, d_model=w["H"], n_layers_enc=6, n_head=8,
                              d_inner=2048, dropout=0.1, d_assigner_hidden=512, w_context=3, n_assigner_layers=3,
                              sos_id=2, vocab_size=w["V"], n_layers_dec=6, spec_aug_cfg=None)


def _train_inputs(w, device, seed):
    B, T, S, V, D = w["B"], w["T"], w["S"], w["V"], w["D"]
    g = torch.Generator(device=device).manual_seed(seed)
    feats = torch.randn(B, T, D, device=device, generator=g)
    lens = torch.randint(int(0.7 * T), T + 1, (B,), device=device, generator=g)
    lens[0] = T
    feats = feats * (torch.arange(T, device=device)[None, :, None] < lens[:, None, None]).float()
    targets = torch.randint(4, V - 1, (B, S), device=device, generator=g)
    tl = torch.randint(max(1, (2 * S) // 3), S + 1, (B,), device=device, generator=g)
    targets = targets * (torch.arange(S, device=device)[None, :] < tl[:, None]).long()
    return feats, lens, targets


def run_train(args, w, rank, world, device, steps=None, emit=True, tf32=False):
    """tf32=True lets torch run the model shell's fp32 Linear / Conv GEMMs on the tensor cores (TF32 inputs, fp32
    accumulate) instead of cuBLAS's SIMT sgemm - reported next to the reference-faithful fp32 step, never instead."""
    import importlib
    import torch.distributed as dist
    import asr_b200 as pkg
    tf32_before = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    if tf32:      # the plain run keeps torch's defaults (fp32 matmul; cuDNN may use TF32 for the convolutions, as in the reference's own torch)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    cm = importlib.import_module("end-to-end_asr_pytorch_b200.transformer.cif_model")
    lossm = importlib.import_module("end-to-end_asr_pytorch_b200.transformer.loss")
    dp = importlib.import_module("end-to-end_asr_pytorch_b200.dp")
    torch.manual_seed(1234)
    model = cm.CIF_Model.create_model(_model_args(w)).to(device).train()
    dp.broadcast_parameters(model, 0)
    sync = dp.GradAllReduce(model, bucket_mb=25)
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, betas=(0.9, 0.98), eps=1e-9)
    feats, lens, targets = _train_inputs(w, device, 1240 + rank)
    torch.manual_seed(100 + rank)          # per-rank noise / dropout streams
    n_params = sum(p.numel() for p in model.parameters())

    def step(f, l, t):
        sync.reset()
        ctc_logits, len_ctc, _num, num, logits = model(f, l, t)
        qua, ctc, ce = lossm.cal_ctc_qua_ce_loss(ctc_logits, len_ctc, _num, num, logits, t, smoothing=0.1)
        loss = 0.001 * qua + ctc + ce
        loss.backward()
        sync.finish()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    K, W = (steps or args.steps), max(args.warmup, 3)
    for _ in range(W):
        step(feats, lens, targets)
    barrier()
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    l0 = pkg._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        loss = step(feats, lens, targets)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = pkg._lib.launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    # end to end: features / targets from pinned host memory every step, loss read back
    host = [x.cpu().pin_memory() for x in (feats, lens, targets)]
    bufs = [torch.empty_like(x) for x in (feats, lens, targets)]
    Ke = max(3, min(K, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        for b_, h_ in zip(bufs, host):
            b_.copy_(h_, non_blocking=True)
        lv = step(*bufs).item()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sync.remove()
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": world * w["B"] / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (attention core bf16)", "data": "synthetic",
                "config": dict(workload=args.workload, per_gpu=w, model="CIF_Model reference recipe defaults",
                               params=n_params, grad_allreduce_bytes=sync.grad_bytes(),
                               parallelism="dp%d by utterance, NCCL gradient all-reduce (bucketed, overlapped)" % world,
                               note="training mode: attention-probability dropout 0.1 applied inside the tcgen05 kernels"),
                "clocks": clocks,
                "e2e": {"value": world * w["B"] * Ke / float(dt.item()), "unit": UNIT,
                        "h2d_bytes_per_step": sum(h.numel() * h.element_size() for h in host), "d2h_bytes_per_step": 4,
                        "steps": Ke, "api": "CIF_Model.forward + cal_ctc_qua_ce_loss + backward + all-reduce + Adam"},
                "gpu_launches": int(launches), "last_loss": lv, "roofline": None, "cpu_baseline": None}
        if emit:
            print(json.dumps(line), flush=True)
    del model, opt, sync
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32_before
    torch.cuda.empty_cache()
    return line


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def cpu_reference_step(w, sample, seed, threads=None):
    """The reference's torch-CPU path on `sample` utterances of the workload."""
    from oracle import torch_port
    if threads:
        torch.set_num_threads(threads)
    wc = dict(w, B=sample)
    inp = make_inputs(wc, "cpu", seed)
    g = torch.Generator().manual_seed(seed + 1)
    g_fired = torch.randn(sample, w["S"] + 2, w["H"], generator=g)

    def run():
        t0 = time.perf_counter()
        r = torch_port.joint_hot_path_step(inp["hidden"], inp["alphas"], inp["logits"], inp["in_len"], inp["targets"],
                                           inp["noise"], g_fired)
        return time.perf_counter() - t0, r
    return run


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run = cpu_reference_step(w, args.cpu_sample, 1236)
    for _ in range(max(args.warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = args.cpu_sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload=args.workload, **w, per_step_sample=args.cpu_sample,
                       note="torch-CPU port of the reference path (oracle/torch_port.py); each step = %d utterances "
                            "of the workload shape" % args.cpu_sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d utterances x T=%d, S=%d, V=%d, H=%d per step" % (
                             args.cpu_sample, w["T"], w["S"], w["V"], w["H"])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    if args.workload.startswith("train"):
        run_train(args, w, rank, world, device)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    import asr_b200 as pkg
    if args.ctc_chunks:
        pkg._lib.set_option("ctc_chunks", args.ctc_chunks)
    for kv in args.opt:
        key, val = kv.split("=")
        pkg._lib.set_option(key, int(val))
    launches0 = pkg._lib.launch_count()
    inp = make_inputs(w, device, 1236 + rank)
    hp = HotPath(w, inp, pkg)
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ---------------------------------------------------
    step_fn = hp.step if args.serial else hp.step_overlapped
    for _ in range(W):
        step_fn()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = pkg._lib.launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    live = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_start.record()
    for k in range(K):
        if args.serial:
            step_fn()
        else:
            step_fn(live[k])
    t_end.record()
    barrier()
    clocks = sampler.stop()
    timed_launches = pkg._lib.launch_count() - l0
    total_ms = t_start.elapsed_time(t_end)
    # per-kernel durations: the same K steps again, one kernel at a time on one stream with an event
    # between stages (inside the overlapped region a kernel's events would also time its neighbours)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(K)]
    hp.step()
    barrier()
    for k in range(K):
        hp.step(evs[k])
    barrier()
    stage_ms = [sum(evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(K)) / K for i in range(5)]
    # the dominant kernel inside the timed region: all row-kernel slices of a step, lattices running next to them
    rows_live_ms = stage_ms[0] if args.serial else sum(a.elapsed_time(b) for a, b in live) / K
    t = torch.tensor([total_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / K
    value = world * w["B"] / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host inputs ---------------------------
    e2e = None
    if not args.no_e2e:
        host = {k: inp[k].cpu().pin_memory() for k in ("logits", "hidden", "alphas", "targets", "in_len", "noise")}
        dev_buf = {k: torch.empty_like(inp[k]) for k in host}
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        Ke = max(3, min(K, 5))
        for _ in range(2):
            e2e_step(pkg, w, host, dev_buf, hp.g_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            losses = e2e_step(pkg, w, host, dev_buf, hp.g_out)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * w["B"] * Ke / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(losses.numel() * losses.element_size()) + 8, "steps": Ke,
               "api": "ops.cif + ops.ctc_loss + autograd.backward, pinned-host inputs, losses read back"}
        del host, dev_buf

    # ---- roofline of the dominant kernel, rank 0 --------------------------------------
    peaks, peak_src = load_peaks()
    bm = hp.bytes_model()
    L_out, valid_frames = hp.Lout, hp.valid_frames
    names = ["ctc_rows", "ctc_lattice", "ctc_apply", "cif_fwd", "cif_bwd"]
    kernels = []
    for i, n in enumerate(names):
        ent = {"kernel": n, "bound": "hbm", "ms": stage_ms[i], "share_of_step": stage_ms[i] / sum(stage_ms)}
        if n in bm:
            ent["algorithmic_bytes"] = bm[n]
            ent["GBps"] = bm[n] / (stage_ms[i] * 1e-3) / 1e9
            ent["frac_of_hbm_peak"] = ent["GBps"] / peaks["hbm_gbs"]
        kernels.append(ent)
    if rank == 0:
        asg = assigner_microbench(pkg, w, inp, device)
        for n in ("cif_alpha_fwd", "cif_alpha_bwd"):
            kernels.append({"kernel": n, "bound": "hbm", "ms": asg[n]["ms"], "algorithmic_bytes": asg[n]["algorithmic_bytes"],
                            "GBps": asg[n]["GBps"], "frac_of_hbm_peak": asg[n]["GBps"] / peaks["hbm_gbs"],
                            "in_timed_step": False, "note": "SURVEY 8(f2): assigner tail + alpha scaling, next-row kernel"})
        sa = spec_aug_microbench(pkg, device)
        kernels.append({"kernel": "spec_aug", "bound": "hbm", "ms": sa["ms"], "algorithmic_bytes": sa["algorithmic_bytes"],
                        "GBps": sa["GBps"], "frac_of_hbm_peak": sa["GBps"] / peaks["hbm_gbs"], "shape": sa["shape"],
                        "in_timed_step": False, "note": "SURVEY 8(f4): SpecAugment, three launches incl. the means"})
        lin = linear_microbench(pkg, device)
        for n in ("linear_bias_relu", "linear_residual_layernorm", "linear_f32_3xtf32"):
            kernels.append({"kernel": n, "bound": "tensor", "ms": lin[n]["ms"], "TFLOPs": lin[n]["TFLOPs"],
                            "frac_of_bf16_peak": lin[n]["TFLOPs"] / peaks["bf16_tflops"], "shape": lin[n]["shape"],
                            "torch_cublas_plus_eager_TFLOPs": lin[n]["torch_TFLOPs"], "in_timed_step": False,
                            "note": "SURVEY 8(f3): tcgen05 GEMM with the epilogue fused, next-row kernel"})
        att = attention_microbench(pkg, device)
        for n in ("mha_fwd", "mha_bwd"):
            kernels.append({"kernel": n, "bound": "tensor", "ms": att[n]["ms"], "TFLOPs": att[n]["TFLOPs"],
                            "frac_of_bf16_peak": att[n]["TFLOPs"] / peaks["bf16_tflops"], "shape": att[n]["shape"],
                            "in_timed_step": False})
    traffic = load_traffic()
    dom = kernels[0]
    live_gbps = bm["ctc_rows"] / (rows_live_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "asr::ctc_rows_kernel", "achieved": live_gbps, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": live_gbps / peaks["hbm_gbs"], "peak_source": peak_src,
                "traffic": traffic.get("ctc_rows_kernel_bytes_per_launch"),
                "algorithmic_bytes_per_launch": bm["ctc_rows"],
                "avg_launch_ms": rows_live_ms,
                "measured": "CUDA events on the launching stream around the row kernel inside the timed region "
                            "(one whole-batch pass = the step's slice launches back to back, lattices of earlier "
                            "slices running next to them); bytes and ncu traffic are for the whole batch",
                "alone": {"avg_launch_ms": stage_ms[0], "achieved": dom["GBps"], "frac": dom["GBps"] / peaks["hbm_gbs"],
                          "measured": "serial pass over the same K steps right after the timed region, one unsliced launch"},
                "serial_step_ms": sum(stage_ms),
                "joint_step_GBps": (bm["ctc_total"] + bm["cif_fwd"] + bm["cif_bwd"]) / (ms_per_step * 1e-3) / 1e9,
                "joint_step_frac": (bm["ctc_total"] + bm["cif_fwd"] + bm["cif_bwd"]) / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "ctc_whole_GBps": bm["ctc_total"] / (sum(stage_ms[:3]) * 1e-3) / 1e9,
                "ctc_whole_frac": bm["ctc_total"] / (sum(stage_ms[:3]) * 1e-3) / 1e9 / peaks["hbm_gbs"]}

    # ---- BASELINE config 5 next to the hot-path number: the whole CIF_Model trained data parallel
    #      (NCCL gradient all-reduce), a few steps, same launch ------------------------------
    train_step = None
    if not args.no_train_step:
        del hp
        torch.cuda.empty_cache()
        tl = run_train(args, dict(WORKLOADS["train"]), rank, world, device, steps=5, emit=False)
        if tl is not None:
            train_step = {k: tl[k] for k in ("value", "unit", "ms_per_step", "e2e", "gpu_launches")}
            train_step.update(workload="train", per_gpu=tl["config"]["per_gpu"], params=tl["config"]["params"],
                              grad_allreduce_bytes=tl["config"]["grad_allreduce_bytes"],
                              parallelism=tl["config"]["parallelism"],
                              dtype="f32 model shell as the reference (cuBLAS SIMT sgemm: ~45 % of the step), bf16 attention core")
        t2 = run_train(args, dict(WORKLOADS["train"]), rank, world, device, steps=5, emit=False, tf32=True)
        if t2 is not None and train_step is not None:
            train_step["with_tf32_matmul"] = {"value": t2["value"], "unit": t2["unit"], "ms_per_step": t2["ms_per_step"],
                                              "note": "same step with torch.backends.*.allow_tf32 = True for the model shell's "
                                                      "Linear / Conv GEMMs (reduced precision: informational, not the reported value)"}

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            run = cpu_reference_step(w, args.cpu_sample, 1236)
            run()
            dt_cpu, _ = run()
            cpu_baseline = {"value": args.cpu_sample / dt_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "%d utterances x T=%d, S=%d, V=%d, H=%d, one step after one warm-up (%.1f s)" % (
                                args.cpu_sample, w["T"], w["S"], w["V"], w["H"], dt_cpu)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload=args.workload, per_gpu=w, L=L_out, valid_frames=valid_frames,
                           parallelism="dp%d by utterance, no data-path collective" % world,
                           schedule="serial, one stream" if args.serial else
                           "one stream; CTC begin (rows + lattices on library streams) / CIF pair (forward: cif_fwd_variant=3) / CTC finish (apply)",
                           l2="inputs (%.1f GB logits + %.1f GB hidden per GPU) exceed the 126 MB L2; no flush needed" % (
                               inp["logits"].numel() * 4 / 1e9, inp["hidden"].numel() * 4 / 1e9)),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(timed_launches),
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "train_step": train_step,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
