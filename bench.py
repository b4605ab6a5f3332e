#!/usr/bin/env python3
"""bench.py - throughput of the CIF + CTC training hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" of the default workload `cif_ctc_joint` is one pass of the hot path over one batch of
synthetic utterances - the work the reference does per training step around its encoder - plus, on N > 1 GPUs, the
data-parallel gradient exchange of that training step:

    CTC loss + gradient on the logits [B,T,V]                  (a4: K1 row pass, K2 lattice, K3 sparse update)
    CIF forward on the encoder frames [B,T,H] + quantity term  (a1-a3)
    CIF backward                                               (a2')
    all-reduce (mean) of the model's fp32 gradients - the 52.3 M parameters of the reference recipe's CIF_Model = 209 MB -
    by this package's kernel over NVLink peer memory (csrc/allreduce.cu; NCCL in 25 MB buckets is the fallback and
    `--allreduce nccl`), launched behind the CTC row pass (from where gradients exist), waited for at the end of the step

Shape = the largest one of BASELINE config 2 (CTC sweep: B=256, T=1600, S=80, V=4233) with the CIF layer of config 4
run on the same batch (H=512).  Every rank processes its own batch (data parallel by utterance, weak scaling).

The JSON line carries:
  value          utterances/s with inputs resident in HBM, CUDA-event timed, max over ranks, all-reduce included
  e2e            same metric through the public Python API (ops.cif / ops.ctc_loss + autograd) with every step's inputs
                 copied from pinned host memory (prefetched on a copy stream while the previous step computes) and the
                 losses read back every step
  roofline       the dominant kernel (ctc_rows) against the measured HBM copy peak, timed live inside the timed region
  mha_roofline   the attention core forward / backward against the measured bf16 peak
  cpu_baseline   the reference's own CPU path on a bounded sample of the same workload, on this box's host cores
  self_check     outputs of the timed (overlapped) schedule compared bit for bit with the serial schedule after the loop
  ctc_sweep      BASELINE config 2: whole-call CTC GB/s at the sweep shapes
  train_step     BASELINE config 5: the whole CIF_Model trained data parallel (CUDA-graph step, gradient all-reduce)
  transformer_step  BASELINE config 3: SpeechTransformer 6+6 bf16 training step
`--impl reference` times the reference's CPU path (rank 0), `--impl reference-gpu` the reference's eager-GPU path
(its Python CIF loop, ATen ctc_loss) on cuda:0, both in the same JSON shape.  The reference modules are the unmodified
files staged under baseline/_ref/src (oracle/build_ref.py); without them the op-for-op port oracle/torch_port.py runs.
"""
import argparse
import ctypes
import importlib
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
PKG = "end-to-end_asr_pytorch_b200"

WORKLOADS = {
    # name: per-GPU batch
    "cif_ctc_joint": dict(B=256, T=1600, S=80, V=4233, H=512),
    "cif_ctc_small": dict(B=32, T=200, S=10, V=4233, H=512),
    # BASELINE config 5: the whole CIF_Model (reference recipe defaults: LFR 4/3, 3 conv layers, 6+6 layers, d_model 512,
    # 8 heads, d_inner 2048, V=4233) trained data-parallel with an NCCL gradient all-reduce; 500 raw frames -> 167 LFR
    # frames x 320 -> 21 encoder frames, 14 labels
    "train": dict(B=64, T=167, S=14, V=4233, H=512, D=320),
    "train_long": dict(B=32, T=534, S=45, V=4233, H=512, D=320),
    # BASELINE config 3: Transformer(Encoder(320, 6, 8, 512, 2048), Decoder(.., 4233, 6, 8, 512, 2048)), batch_frames 15000
    # (transformer.sh:18) = 90 utterances x 167 LFR frames, 14 labels (+ <eos>), bf16
    "transformer_bf16": dict(B=90, T=167, S=14, V=4233, H=512, D=320),
}
CTC_SWEEP = [(32, 200, 10), (64, 400, 20), (128, 800, 40), (32, 1600, 80)]      # BASELINE config 2 (the 256 x 1600 x 80 corner is the default workload)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="cif_ctc_joint", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=8, help="utterances per reference-arm step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--serial", action="store_true", help="time the hot path one kernel at a time on one stream")
    ap.add_argument("--ctc-chunks", type=int, default=0, help="batch slices of the CTC pipeline (0 = library default)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="library tuning knob (asr_set_option), repeatable")
    ap.add_argument("--no-extras", action="store_true", help="skip the microbenches, the sweep and the full-model steps")
    ap.add_argument("--no-graph", action="store_true", help="full-model steps: eager launches instead of a CUDA graph")
    ap.add_argument("--no-allreduce", action="store_true", help="hot-path step without the gradient all-reduce (N > 1)")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "peer", "nccl"],
                    help="gradient all-reduce of the hot-path step at N > 1: this package's peer-memory kernel or NCCL buckets")
    ap.add_argument("--allreduce-at", default="rows", choices=["rows", "start", "end"],
                    help="where the hot-path step launches its gradient all-reduce: behind the CTC row pass (default), at the "
                         "start of the step, or after the last kernel (no overlap) - for measuring the overlap")
    return ap.parse_args()


def pkg(sub=None):
    return importlib.import_module(PKG if sub is None else PKG + "." + sub)


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


def cuda_time(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


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
# gradient buckets of the model the hot path belongs to (the all-reduce of the training step)
# ---------------------------------------------------------------------------------------
def _model_args(w):
    return argparse.Namespace(d_input=80, LFR_m=4, n_conv_layers=3, d_model=w["H"], n_layers_enc=6, n_head=8,
                              d_inner=2048, dropout=0.1, d_assigner_hidden=512, w_context=3, n_assigner_layers=3,
                              sos_id=2, eos_id=3, vocab_size=w["V"], n_layers_dec=6, spec_aug_cfg=None)


def cif_model_param_count(w):
    cm = pkg("transformer.cif_model")
    with torch.device("meta"):
        model = cm.CIF_Model.create_model(_model_args(w))
    return sum(p.numel() for p in model.parameters())


class GradBuckets:
    """The fp32 gradients of the reference recipe's CIF_Model and their all-reduce (mean).  The hot-path step stands for
    the model's backward pass here, so the buffer holds synthetic gradients; its size and the collective are those of the
    real training step (bench.py --workload train runs it).
    backend "peer": one symmetric-memory buffer, averaged by this package's kernel over NVLink peer memory
    (dp.PeerAllReduce -> asr_allreduce_mean_f32); "nccl": 25 MB buckets, one NCCL all-reduce each (round 2's first
    version, and the fallback when symmetric memory is not available)."""

    def __init__(self, n_params, device, world, backend="auto", bucket_mb=25.0):
        self.world = world
        self.bytes = 4 * n_params
        self.n_params = n_params
        dp = pkg("dp")
        self.peer, fallback = None, ""
        if backend == "auto":
            backend = "peer" if (world > 1 and dp.PeerAllReduce.available(device)) else "nccl"
            if backend == "peer":
                # a box without peer access / symmetric memory: every rank fails alike and takes NCCL
                try:
                    self.peer = dp.PeerAllReduce(n_params, device)
                except Exception as e:      # noqa: BLE001
                    backend, fallback = "nccl", " (symmetric memory unavailable: %s)" % str(e)[:80]
        self.backend = backend
        g = torch.Generator(device=device).manual_seed(99)
        self.handles = []
        if backend == "peer":
            if self.peer is None:
                self.peer = dp.PeerAllReduce(n_params, device)
            self.peer.flat.copy_(torch.randn(n_params, device=device, generator=g) * 1e-3)
            self.flat = [self.peer.flat]
            self.how = ("this package's all-reduce kernel over NVLink peer memory (asr_allreduce_mean_f32, %s, %d CTAs; "
                        "torch symmetric memory only allocates and exchanges the handles), one launch over the %d fp32 "
                        "gradients of the recipe's CIF_Model" % (self.peer.flavour(), self.peer.ctas, n_params))
        else:
            per = int(bucket_mb * 1024 * 1024) // 4
            sizes = [per] * (n_params // per) + ([n_params % per] if n_params % per else [])
            self.flat = [torch.randn(n, device=device, generator=g) * 1e-3 for n in sizes]
            self.how = ("NCCL all-reduce (mean) of %d fp32 gradient buckets = the %d parameters of the recipe's CIF_Model%s"
                        % (len(self.flat), n_params, fallback))

    def launch(self):
        import torch.distributed as dist
        if self.world <= 1:
            return
        if self.peer is not None:
            self.peer.launch()
        else:
            self.handles = [dist.all_reduce(f, op=dist.ReduceOp.AVG, async_op=True) for f in self.flat]

    def wait(self):
        if self.peer is not None:
            if self.world > 1:
                self.peer.wait()  # stream-level: the compute stream waits for the side stream, the host does not
            return
        for h in self.handles:
            h.wait()          # stream-level: the compute stream waits for NCCL, the host does not
        self.handles = []


# ---------------------------------------------------------------------------------------
# our arm: the hot path through the C ABI
# ---------------------------------------------------------------------------------------
class HotPath:
    """Preallocated buffers + direct C-ABI calls (what the autograd wrappers do, minus the allocator)."""

    def __init__(self, w, inp, buckets=None, allreduce_at="rows"):
        self.w, self.inp = w, inp
        self.allreduce_at = allreduce_at
        self.lib = pkg("_lib")
        self.L = self.lib.lib()
        self.buckets = buckets
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
        # next to the lattices the warp-specialised CIF forward disturbs them least (measured: 2.79 ms per step against
        # 2.83 ms with the library's stand-alone choice, the one-warp TMA pipeline): a per-call hint, not a global option
        self.cif_hint_overlapped = 3 if w["T"] >= 64 and w["H"] % 4 == 0 else 0

    def _ctc_args(self):
        w, i, p = self.w, self.inp, self.lib.ptr
        return (p(i["logits"]), p(i["targets"]), p(i["in_len"]), p(i["tgt_len"]), w["B"], w["T"], w["V"], w["S"],
                w["V"] - 1, p(self.nll), p(self.g_logits), p(self.ws), self.ws_bytes)

    def ctc(self, stages):
        self.lib.check(self.L.asr_ctc_stages_f32(*self._ctc_args(), stages, self.lib.stream_ptr()), "asr_ctc_stages_f32")

    def cif_fwd(self, hint=0):
        w, i, p = self.w, self.inp, self.lib.ptr
        self.lib.check(self.L.asr_cif_fwd_hint_f32(
            p(i["hidden"]), p(self.alphas), 0.95, w["B"], w["T"], w["H"], self.Lout, p(self.out), p(self.fire_t),
            p(self.n_fired), p(self.cur), p(self.rem), p(self.sched), p(self.asum), p(self.num), p(self.qua), hint,
            self.lib.stream_ptr()), "asr_cif_fwd_hint_f32")

    def cif_bwd(self):
        w, i, p = self.w, self.inp, self.lib.ptr
        self.lib.check(self.L.asr_cif_bwd_f32(
            p(i["hidden"]), p(self.g_out), p(self.n_fired), p(self.cur), p(self.rem), p(self.sched), w["B"], w["T"],
            w["H"], self.Lout, p(self.g_hidden), p(self.g_alpha), p(self.cif_ws), self.cif_ws.numel() * 4,
            self.lib.stream_ptr()), "asr_cif_bwd_f32")

    def step_overlapped(self, ev=None):
        """One hot-path pass the way the library is meant to be driven: one stream, the CTC call in its two phases with
        the CIF forward/backward pair queued in between, where it runs next to the last slice's latency-bound lattice.
        The gradient all-reduce of the training step (N > 1) goes out behind the CTC row pass - the point of the step
        from which gradients exist (the row pass writes the dense part of d loss / d logits) - runs on a side stream next
        to the lattices, the CIF pair and the apply pass, and is waited for at the end of the step.
        ev = (before, after): CUDA events around the row kernels (all slices; they are the only work begin puts on this
        stream), i.e. the dominant kernel timed inside the timed region."""
        args = self._ctc_args() + (self.lib.stream_ptr(),)
        ticket = ctypes.c_int(0)
        if self.buckets is not None and self.allreduce_at == "start":
            self.buckets.launch()
        if ev is not None:
            ev[0].record()
        self.lib.check(self.L.asr_ctc_begin_f32(*args, ctypes.byref(ticket)), "asr_ctc_begin_f32")
        if ev is not None:
            ev[1].record()
        if self.buckets is not None and self.allreduce_at == "rows":
            self.buckets.launch()      # behind the row pass on this stream: the dense CTC gradient exists from here on
        self.cif_fwd(self.cif_hint_overlapped)
        self.cif_bwd()
        self.lib.check(self.L.asr_ctc_finish_f32(*args, ticket.value), "asr_ctc_finish_f32")
        if self.buckets is not None:
            if self.allreduce_at == "end":
                self.buckets.launch()
            self.buckets.wait()

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
        if self.buckets is not None:
            self.buckets.launch()
            self.buckets.wait()

    def snapshot(self):
        """What the self-check compares: per-utterance nll, fire positions and counts, the CIF outputs and gradients of the
        first and last utterances, the CTC gradient rows of the first and last utterance."""
        return {"nll": self.nll.clone(), "fire_t": self.fire_t.clone(), "n_fired": self.n_fired.clone(),
                "cif_out": torch.cat([self.out[:2], self.out[-2:]]).clone(), "g_alpha": self.g_alpha.clone(),
                "g_hidden": torch.cat([self.g_hidden[:1], self.g_hidden[-1:]]).clone(),
                "g_logits": torch.cat([self.g_logits[:1], self.g_logits[-1:]]).clone()}

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


def self_check(hp):
    """Bit-for-bit comparison of the timed (overlapped, sliced, hinted) schedule with the serial one-stream schedule, plus
    sanity of the values themselves; raises on any difference."""
    torch.cuda.synchronize()
    hp.step_overlapped()
    torch.cuda.synchronize()
    a = hp.snapshot()
    for t in (hp.nll, hp.fire_t, hp.n_fired, hp.out, hp.g_alpha, hp.g_hidden):
        t.zero_()
    hp.g_logits[:1].zero_()
    hp.g_logits[-1:].zero_()
    hp.step()
    torch.cuda.synchronize()
    b = hp.snapshot()
    res = {k: bool(torch.equal(a[k].view(torch.int32) if a[k].dtype == torch.float32 else a[k],
                               b[k].view(torch.int32) if b[k].dtype == torch.float32 else b[k])) for k in a}
    finite = bool(torch.isfinite(a["nll"]).all())
    fired_ok = bool((a["n_fired"] <= hp.Lout).all() and (a["n_fired"] > 0).all())
    res.update(nll_finite=finite, fires_within_L=fired_ok, mean_nll=float(a["nll"].mean()))
    bad = [k for k, v in res.items() if v is False]
    if bad:
        raise SystemExit("bench.py self-check failed: the timed schedule differs from the serial one in %s" % bad)
    return res


class E2EPipe:
    """Public-API step with host inputs.  Every step's inputs come from pinned host memory: the copies of step k+1 run on a
    copy stream while step k computes (two device buffer sets), the way a training input pipeline prefetches; the losses are
    read back (device -> host, synchronising) every step.
    fused=True (default): the step starts where the reference's does, from the encoder outputs [B,T,H] - the vocabulary
    projection `ctc_fc` (cif_model.py:38; its weight is a resident parameter) runs fused with the CTC loss
    (ops.ctc_fc_loss), so the [B,T,V] logits never cross PCIe (0.84 GB per step instead of 7.8 GB) and the
    step also produces d loss / d ctc_fc.weight.  fused=False: round 1's leg, the logits themselves come from the host."""

    def __init__(self, w, inp, g_out, fused=True):
        self.ops = pkg("ops")
        self.w, self.g_out, self.fused = w, g_out, fused
        self.KEYS = ("hidden", "alphas", "targets", "in_len", "noise") + (() if fused else ("logits",))
        if fused:
            g = torch.Generator(device=inp["hidden"].device).manual_seed(77)
            self.weight = (torch.randn(w["V"], w["H"], device=inp["hidden"].device, generator=g) * w["H"] ** -0.5).requires_grad_(True)
        self.host = {k: inp[k].cpu().pin_memory() for k in self.KEYS}
        self.bufs = [{k: torch.empty_like(inp[k]) for k in self.KEYS} for _ in range(2)]
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.copy_stream = torch.cuda.Stream()
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self.k = 0
        self.d2h_bytes = 8

    def prefetch(self, slot):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])          # the step that used this buffer set has finished
            for k in self.KEYS:
                self.bufs[slot][k].copy_(self.host[k], non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def start(self):
        for s in (0, 1):
            self.free[s].record()
        self.k = 0
        self.prefetch(0)

    def step(self):
        slot = self.k & 1
        self.prefetch(slot ^ 1)                                   # next step's inputs, under this step's kernels
        torch.cuda.current_stream().wait_event(self.ready[slot])
        buf = self.bufs[slot]
        grads = ("hidden", "alphas") + (() if self.fused else ("logits",))
        for k in grads:
            buf[k].requires_grad_(True)
        hidden, alphas_raw = buf["hidden"], buf["alphas"]
        _num, num, alphas = scale_alphas(alphas_raw, buf["targets"], buf["noise"])
        fired = self.ops.cif(hidden, alphas, 0.95)
        qua = torch.pow(_num - num, 2).mean()
        if self.fused:
            ctc = self.ops.ctc_fc_loss(hidden, self.weight, buf["in_len"], buf["targets"])
        else:
            ctc = self.ops.ctc_loss(buf["logits"], buf["in_len"], buf["targets"])
        total = ctc + 0.001 * qua + (fired * self.g_out[:, :fired.size(1)]).sum()
        total.backward()
        losses = torch.stack([ctc.detach(), qua.detach()])
        self.free[slot].record()
        losses = losses.cpu()                                     # D2H, synchronises
        for k in grads:
            buf[k].grad = None
            buf[k].requires_grad_(False)
        if self.fused:
            self.weight.grad = None
        self.k += 1
        return losses


# ---------------------------------------------------------------------------------------
# microbenches of the other hot-path kernels (rank 0, outside the timed region)
# ---------------------------------------------------------------------------------------
def assigner_microbench(w, inp, device, iters=5):
    """CIF weight producer (assigner tail + scaling) on the bench shape: x = the encoder
    output of the step [B,T,H], ragged lengths.  HBM-bound: forward reads the valid rows of x once,
    backward reads them once more and writes g_x once."""
    lib = pkg("_lib")
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
        ms = cuda_time(fn, iters)
        res[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / (ms * 1e-3) / 1e9}
    return res


def ctc_fc_microbench(w, inp, device, peaks, iters=3):
    """the vocabulary projection fused with the CTC loss on the bench batch - encoder outputs [B,T,H] ->
    loss, d hidden, d weight, the logits never leaving the call - and its three fp32 tensor-core GEMMs on their own."""
    ops = pkg("ops")
    B, T, V, H = w["B"], w["T"], w["V"], w["H"]
    g = torch.Generator(device=device).manual_seed(78)
    weight = (torch.randn(V, H, device=device, generator=g) * H ** -0.5).requires_grad_(True)
    hidden = inp["hidden"].detach().clone().requires_grad_(True)

    def whole():
        loss = ops.ctc_fc_loss(hidden, weight, inp["in_len"], inp["targets"])
        loss.backward()
        hidden.grad = weight.grad = None
    ms = cuda_time(whole, iters, warm=1)
    M = B * T
    flop = 2.0 * M * V * H
    res = {"ctc_fc_loss": {"ms": ms, "utts_per_s": B / ms * 1e3, "TFLOPs": 3 * flop / ms / 1e9,
                           "note": "whole call: projection GEMM + CTC (in place on the padded logits) + d hidden and d weight GEMMs; "
                                   "fp32 in / out, three TF32 products per K step"}}
    h2 = hidden.detach().reshape(M, H)
    buf = torch.empty(M, (V + 3) // 4 * 4, device=device)
    wd = weight.detach()
    for name, fn in (("gemm_f32_fwd (h W^T)", lambda: ops.gemm_f32(h2, wd, out=buf, split_k=False)),
                     ("gemm_f32_dx (g W, W MN-major)", lambda: ops.gemm_f32(buf[:, :V], wd, b_mn_major=True, split_k=False)),
                     ("gemm_f32_dw (g^T h, both MN-major, split-K)", lambda: ops.gemm_f32(buf[:, :V], h2, a_mn_major=True, b_mn_major=True))):
        t = cuda_time(fn, iters, warm=1)
        res[name] = {"ms": t, "TFLOPs": flop / t / 1e9}
    del buf
    # torch's own path for the same three products (cuBLAS SIMT sgemm in fp32)
    lg = torch.empty(M, V, device=device)
    t = cuda_time(lambda: torch.mm(h2, wd.t(), out=lg), 2, warm=1)
    res["torch_fp32_fwd (cuBLAS)"] = {"ms": t, "TFLOPs": flop / t / 1e9}
    del lg
    return res


def spec_aug_microbench(device, iters=5):
    """SpecAugment on the device on a batch of LFR-stacked fbank of the bench size
    (B=256 x T=1600 x 320 bins, ragged, two bands + two spans per utterance).  HBM-bound: the batch is read
    once for the two means; only the masked cells are written."""
    ops = pkg("ops")
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
    ms = cuda_time(lambda: ops.spec_aug_apply(x, lens, f0, fw, t0, tw), iters)
    return {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / (ms * 1e-3) / 1e9,
            "shape": "B=%d T=%d V=%d, %d bands + %d spans per utterance" % (B, T, V, R, R)}


def layer_norm_microbench(device, iters=10):
    """LayerNorm(dropout(y) + residual) * non-pad mask, forward and backward (csrc/ln.cu) at the bench batch
    (M = 64 x 1600 frames, d_model 512; y bf16 as under autocast, residual / out fp32, dropout 0.1): HBM-bound, every
    tensor (105 - 210 MB) larger than L2.  Algorithmic bytes per element: forward 2 (y) + 4 (residual) + 4 (z, saved) +
    4 (out); backward 4 (g) + 4 (z) + 4 (dz) + 2 (dy); the gamma / beta partial rows and their column sum are noise."""
    ops = pkg("ops")
    M, D = 64 * 1600, 512
    g = torch.Generator(device=device).manual_seed(10)
    y = torch.randn(M, D, device=device, generator=g).bfloat16().requires_grad_(True)
    res = torch.randn(M, D, device=device, generator=g).requires_grad_(True)
    gam = torch.ones(D, device=device, requires_grad=True)
    bet = torch.zeros(D, device=device, requires_grad=True)
    mask = (torch.rand(M, device=device, generator=g) < 0.8).float()
    go = torch.randn(M, D, device=device, generator=g)
    out = [None]

    def fwd():
        out[0] = ops.residual_layer_norm(y, res, gam, bet, 1e-5, dropout_p=0.1, seed=11, row_scale=mask)

    def bwd():
        out[0].backward(go, retain_graph=True)
        y.grad = res.grad = gam.grad = bet.grad = None
    f = cuda_time(fwd, iters)
    b = cuda_time(bwd, iters)
    nb = 14 * M * D
    return {"ln_fwd": {"ms": f, "algorithmic_bytes": nb, "GBps": nb / (f * 1e-3) / 1e9},
            "ln_bwd": {"ms": b, "algorithmic_bytes": nb, "GBps": nb / (b * 1e-3) / 1e9},
            "shape": "M=%d D=%d, y bf16, residual / out fp32, dropout 0.1, row mask" % (M, D)}


def linear_microbench(device, iters=10):
    """Fused tcgen05 linear layers on the feed-forward block of the encoder at the bench batch
    (M = 64 x 1600 frames, d_model 512, d_inner 2048), next to torch's cuBLAS + eager epilogue on the same tensors."""
    ops = pkg("ops")
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
            ("linear_residual_layernorm_one_kernel", lambda: ops.linear_residual_layernorm(h, w2, b2, x, gam, bet, one_kernel=True),
             lambda: F.layer_norm(F.linear(h, w2, b2h) + x, (d,), gamh, beth)),
            ("linear_residual_layernorm", lambda: ops.linear_residual_layernorm(h, w2, b2, x, gam, bet),
             lambda: F.layer_norm(F.linear(h, w2, b2h) + x, (d,), gamh, beth))]
    res = {}
    flop = 2.0 * M * d * di
    for name, ours, ref in rows:
        t = [cuda_time(fn, iters, warm=3) for fn in (ours, ref)]
        res[name] = {"ms": t[0], "TFLOPs": flop / t[0] / 1e9, "torch_ms": t[1], "torch_TFLOPs": flop / t[1] / 1e9,
                     "shape": "M=%d K=%d N=%d bf16" % ((M, d, di) if name == "linear_bias_relu" else (M, di, d))}
    # fp32 in / fp32 out on the tensor cores (three TF32 products per tile) next to torch's fp32 GEMM (cuBLAS SIMT sgemm)
    xf, wf = x.float(), w1.float()
    t = [cuda_time(fn, iters, warm=3) for fn in (lambda: ops.linear_f32(xf, wf, b1), lambda: F.linear(xf, wf, b1))]
    res["linear_f32_3xtf32"] = {"ms": t[0], "TFLOPs": flop / t[0] / 1e9, "torch_ms": t[1], "torch_TFLOPs": flop / t[1] / 1e9,
                                "shape": "M=%d K=%d N=%d f32" % (M, d, di)}
    return res


def attention_microbench(device, peaks, iters=5):
    """tcgen05 attention core: the SURVEY 8(d) microbench shape (B*heads = 128, L = 2048, d = 64), forward / backward,
    without and with the training-mode dropout; the causal + key-padding case; and the model shapes (encoder self
    attention L = 21 and 167 at B = 64 / 90, decoder self / cross attention U = 15) as latencies.  CUDA events."""
    lib = pkg("_lib")
    L = lib.lib()
    p, sp = lib.ptr, lib.stream_ptr
    g = torch.Generator(device=device).manual_seed(5)
    peak = peaks["bf16_tflops"]
    out = {"shapes": [], "model_shapes": []}

    def run(B, Lq, Lk, H, causal=0, ragged=False, drop=0.0, it=iters):
        q, do = (torch.randn(B, Lq, H, 64, device=device, generator=g).to(torch.bfloat16) for _ in range(2))
        k, v = (torch.randn(B, Lk, H, 64, device=device, generator=g).to(torch.bfloat16) for _ in range(2))
        o = torch.empty_like(q)
        lse = torch.empty(B, H, Lq, device=device)
        gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        wsb = L.asr_mha_bwd_workspace_bytes(B, H, Lq, Lk, 64)
        ws = torch.empty(wsb // 4 + 1, device=device)
        kv = torch.randint(int(0.6 * Lk), Lk + 1, (B,), device=device, generator=g).to(torch.int32) if ragged else None

        def fwd():
            lib.check(L.asr_mha_fwd_dropout_bf16(p(q), p(k), p(v), p(kv), None, causal, B, H, Lq, Lk, 64, 0.125, drop, 1234,
                                                 p(o), p(lse), sp()), "mha_fwd")

        def bwd():
            lib.check(L.asr_mha_bwd_dropout_bf16(p(q), p(k), p(v), p(o), p(do), p(lse), p(kv), None, causal, B, H, Lq, Lk, 64,
                                                 0.125, drop, 1234, p(gq), p(gk), p(gv), p(ws), wsb, sp()), "mha_bwd")
        # useful products only: the causal triangle and the valid keys
        if ragged:
            pairs = float((kv.double() * Lq).sum().item()) if not causal else float(sum(
                min(int(n), Lq) * (min(int(n), Lq) + 1) / 2 + max(0, Lq - int(n)) * int(n) for n in kv.tolist()))
        else:
            pairs = B * (Lq * (Lq + 1) / 2 if causal else Lq * Lk)
        # best of three runs of `it` calls each (the peak these are compared with is a best-of-ten figure: MEASURED_PEAKS.json)
        t_f = min(cuda_time(fwd, it, warm=3) for _ in range(3))
        t_b = min(cuda_time(bwd, it, warm=3) for _ in range(3))
        return t_f, t_b, 4.0 * H * 64 * pairs, 10.0 * H * 64 * pairs

    for name, kw in (("L=2048", {}), ("L=2048 dropout 0.1", dict(drop=0.1)), ("L=2048 causal + key padding", dict(causal=1, ragged=True)),
                     ("L=4096", dict(Lq=4096, B=8)), ("L=512 (persistent forward kernel)", dict(Lq=512, B=64)),
                     ("L=1024 (persistent forward kernel)", dict(Lq=1024, B=32))):
        B, Lq = kw.pop("B", 16), kw.pop("Lq", 2048)
        t_f, t_b, ff, fb = run(B, Lq, Lq, 8, **kw)
        out["shapes"].append({"shape": "B=%d heads=8 %s d=64 bf16" % (B, name), "fwd_ms": t_f, "bwd_ms": t_b,
                              "fwd_TFLOPs": ff / t_f / 1e9, "bwd_TFLOPs": fb / t_b / 1e9,
                              "fwd_frac_of_bf16_peak": ff / t_f / 1e9 / peak, "bwd_frac_of_bf16_peak": fb / t_b / 1e9 / peak})
    for name, (B, Lq, Lk, causal) in (("encoder self, conv front end (B=64, L=21)", (64, 21, 21, 0)),
                                      ("encoder self, LFR frames (B=90, L=167)", (90, 167, 167, 0)),
                                      ("decoder self (B=90, U=15, causal)", (90, 15, 15, 1)),
                                      ("decoder cross (B=90, U=15 x L=167)", (90, 15, 167, 0))):
        t_f, t_b, ff, fb = run(B, Lq, Lk, 8, causal=causal, ragged=True, drop=0.1, it=20)
        out["model_shapes"].append({"shape": name + ", dropout 0.1", "fwd_us": t_f * 1e3, "bwd_us": t_b * 1e3,
                                    "fwd_TFLOPs": ff / t_f / 1e9, "bwd_TFLOPs": fb / t_b / 1e9})
    return out


def ctc_sweep(device, peaks):
    """BASELINE config 2: whole CTC call (rows + lattice + apply, the library's own slicing) at the sweep shapes,
    algorithmic bytes = 8 V per valid frame."""
    lib = pkg("_lib")
    L = lib.lib()
    p = lib.ptr
    V = 4233
    rows = []
    for B, T, S in CTC_SWEEP:
        inp = make_inputs(dict(B=B, T=T, S=S, V=V, H=8), device, 1236)
        nll = torch.empty(B, device=device)
        g = torch.empty_like(inp["logits"])
        wsb = L.asr_ctc_workspace_bytes(B, T, V, S)
        ws = torch.empty(wsb // 4 + 1, device=device)
        # inputs below the 126 MB L2 would be served from it on a repeat: rotate over enough copies to exceed it
        n_copies = max(1, int(2 * 126e6 // (inp["logits"].numel() * 4)) + 1)
        copies = [inp["logits"]] + [inp["logits"].clone() for _ in range(n_copies - 1)]
        state = {"i": 0}

        def call():
            lg = copies[state["i"] % n_copies]
            state["i"] += 1
            lib.check(L.asr_ctc_fwd_bwd_f32(p(lg), p(inp["targets"]), p(inp["in_len"]), p(inp["tgt_len"]), B, T, V, S, V - 1,
                                            p(nll), p(g), p(ws), wsb, lib.stream_ptr()), "asr_ctc_fwd_bwd_f32")
        ms = cuda_time(call, 10, warm=3)
        nbytes = 8 * V * int(inp["in_len"].sum().item())
        rows.append({"B": B, "T": T, "S": S, "V": V, "ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / ms / 1e6,
                     "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"], "utts_per_s": B / ms * 1e3,
                     "l2": "rotating over %d input copies (> 2 x L2)" % n_copies if n_copies > 1 else "input exceeds L2"})
        del copies, g, inp
    return rows


# ---------------------------------------------------------------------------------------
# full-model training steps: config 5 (CIF_Model, fp32 shell) and config 3 (Transformer, bf16)
# ---------------------------------------------------------------------------------------
def _train_inputs(w, device, seed):
    B, T, S, V, D = w["B"], w["T"], w["S"], w["V"], w["D"]
    g = torch.Generator(device=device).manual_seed(seed)
    feats = torch.randn(B, T, D, device=device, generator=g)
    lens = torch.randint(int(0.7 * T), T + 1, (B,), device=device, generator=g)
    lens[0] = T
    feats = feats * (torch.arange(T, device=device)[None, :, None] < lens[:, None, None]).float()
    targets = torch.randint(4, V - 1, (B, S), device=device, generator=g)
    tl = torch.randint(max(1, (2 * S) // 3), S + 1, (B,), device=device, generator=g)
    tl[0] = S
    targets = targets * (torch.arange(S, device=device)[None, :] < tl[:, None]).long()
    return feats, lens, targets


def run_train(args, wname, rank, world, device, steps=None, emit=True, tf32=False, graph=True, torch_linear=False):
    """One optimiser step of a whole model per "step": forward, the reference's losses, backward, gradient all-reduce
    (NCCL, mean over ranks), Adam.  wname = "train" / "train_long": CIF_Model (config 5) with the fp32 shell of the
    reference; "transformer_bf16": Transformer (config 3) under bf16 autocast.
    graph=True: forward + backward are captured once in a CUDA graph and replayed (no host syncs inside: static_shapes,
    device-side dropout seeds), the all-reduce of the flat gradient buckets and the fused Adam step are launched after each
    replay; graph=False: eager launches with the all-reduce overlapped from gradient hooks (dp.GradAllReduce).
    tf32=True lets torch run the shell's fp32 Linear / Conv GEMMs with TF32 inputs - reported next to the
    reference-faithful fp32 step, never instead."""
    import torch.distributed as dist
    w = dict(WORKLOADS[wname])
    torch.backends.cudnn.benchmark = True      # the conv front end is cuDNN's (out of scope): let it pick its fastest algorithms
    tf32_before = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    if tf32:      # the plain run keeps torch's defaults (fp32 matmul; cuDNN may use TF32 for the convolutions, as in the reference's own torch)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    lossm, dp, ops, lib = pkg("transformer.loss"), pkg("dp"), pkg("ops"), pkg("_lib")
    module = pkg("transformer.module")
    switches_before = (module.USE_TENSOR_CORE_FP32, module.USE_TENSOR_CORE_BF16)
    if torch_linear:      # A/B: the shell's Linear layers on torch's F.linear (cuBLAS) instead of this package's GEMMs
        module.USE_TENSOR_CORE_FP32 = module.USE_TENSOR_CORE_BF16 = False
    bf16 = wname == "transformer_bf16"
    torch.manual_seed(1234)
    if bf16:
        model = pkg("transformer.transformer").Transformer.create_model(_model_args(w)).to(device).train()
        model.decoder.assume_full_width = graph
    else:
        model = pkg("transformer.cif_model").CIF_Model.create_model(_model_args(w)).to(device).train()
        model.static_shapes = graph
        model.fused_ctc_fc = True          # ctc_fc projection fused with the CTC loss
    dp.broadcast_parameters(model, 0)
    sync = dp.GradAllReduce(model, bucket_mb=25, overlap=not graph)
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, betas=(0.9, 0.98), eps=1e-9, fused=True)
    feats, lens, targets = _train_inputs(w, device, 1240 + rank)
    torch.manual_seed(100 + rank)          # per-rank noise / dropout streams
    n_params = sum(p.numel() for p in model.parameters())
    static_in = [feats.clone(), lens.clone(), targets.clone()]

    def fwd_bwd():
        sync.reset()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            if bf16:
                logits, targets_eos = model(*static_in)
                loss = lossm.cal_ce_loss(logits.float(), targets_eos, smoothing=0.1)
            else:
                ctc_logits, len_ctc, _num, num, logits = model(*static_in)
                qua, ctc, ce = lossm.cal_ctc_qua_ce_loss(ctc_logits, len_ctc, _num, num, logits, static_in[2], smoothing=0.1)
                loss = 0.001 * qua + ctc + ce
        loss.backward()
        return loss.detach()

    K, W = (steps or args.steps), max(args.warmup, 3)
    if graph:
        ds = ops.DropoutSeed(device, seed=4242 + rank)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), ops.device_dropout_seed(ds):
            for _ in range(3):          # allocator / cuBLAS / cuDNN warm-up outside the capture
                fwd_bwd()
                ds.advance()
                sync.finish()
                opt.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        cg = torch.cuda.CUDAGraph()
        with ops.device_dropout_seed(ds), torch.cuda.graph(cg):
            static_loss = fwd_bwd()
            ds.advance()

        def step(f, l, t):
            if f is not static_in[0]:
                for dst, src in zip(static_in, (f, l, t)):
                    dst.copy_(src, non_blocking=True)
            cg.replay()
            sync.finish()
            opt.step()
            return static_loss
    else:
        def step(f, l, t):
            if f is not static_in[0]:
                for dst, src in zip(static_in, (f, l, t)):
                    dst.copy_(src, non_blocking=True)
            loss = fwd_bwd()
            sync.finish()
            opt.step()
            return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(W):
        step(*static_in)
    barrier()
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        loss = step(*static_in)
    e1.record()
    barrier()
    clocks = sampler.stop()
    # kernels of this repository per step: counted on one eager step (a graph replay launches the same kernels
    # without passing through the library's host entry points)
    l0 = lib.launch_count()
    if graph:
        with ops.device_dropout_seed(ds):
            fwd_bwd()
    else:
        fwd_bwd()
    launches = (lib.launch_count() - l0) * K
    sync.finish()
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    # end to end: features / targets from pinned host memory every step, loss read back
    host = [x.cpu().pin_memory() for x in (feats, lens, targets)]
    Ke = max(3, min(K, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        lv = float(step(*host).item())
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    sync.remove()
    line = None
    if rank == 0:
        model_name = ("Transformer 6 enc / 6 dec, d_model 512, 8 heads, d_inner 2048 (BASELINE config 3)" if bf16 else
                      "CIF_Model reference recipe defaults (BASELINE config 5)")
        line = {"metric": METRIC, "value": world * w["B"] / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16 autocast (fp32 master weights, fp32 softmax statistics)" if bf16 else "f32 (attention core bf16)",
                "data": "synthetic",
                "config": dict(workload=wname, model=model_name, params=n_params, grad_allreduce_bytes=sync.grad_bytes(), **w,
                               launch="CUDA graph of forward + backward, then bucket all-reduce + fused Adam" if graph else
                               "eager launches, all-reduce overlapped from gradient hooks",
                               parallelism="dp%d by utterance, gradient all-reduce (%s) inside every timed step" % (
                                   world, "this package's peer-memory kernel" if getattr(sync, "peer", None) is not None else "NCCL"),
                               note="training mode: attention-probability dropout 0.1 applied inside the tcgen05 kernels"),
                "clocks": clocks,
                "e2e": {"value": world * w["B"] * Ke / float(dt.item()), "unit": UNIT,
                        "h2d_bytes_per_step": sum(h.numel() * h.element_size() for h in host), "d2h_bytes_per_step": 4,
                        "steps": Ke, "api": "model.forward + reference loss function + backward + all-reduce + Adam"},
                "gpu_launches": int(launches), "last_loss": lv, "roofline": None, "cpu_baseline": None}
        if emit:
            print(json.dumps(line), flush=True)
    del model, opt, sync
    module.USE_TENSOR_CORE_FP32, module.USE_TENSOR_CORE_BF16 = switches_before
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


# ---------------------------------------------------------------------------------------
# reference arms: the reference's own path on the host cores (or, eager, on the GPU)
# ---------------------------------------------------------------------------------------
def reference_step_fn(w, sample, seed, device="cpu"):
    """-> (run, kind): run() times one pass of the reference's hot path on `sample` utterances of the workload."""
    from oracle import ref_loader, torch_port
    wc = dict(w, B=sample)
    inp = make_inputs(wc, device, seed)
    g = torch.Generator(device=device).manual_seed(seed + 1)
    g_fired = torch.randn(sample, w["S"] + 2, w["H"], device=device, generator=g)
    if ref_loader.available():
        ns = ref_loader.load(cpu_shim=(device == "cpu"))
        fn = lambda: ref_loader.joint_hot_path_step(ns, inp["hidden"], inp["alphas"], inp["logits"], inp["in_len"],  # noqa: E731
                                                    inp["targets"], inp["noise"], g_fired)
        kind = "reference"
    else:
        fn = lambda: torch_port.joint_hot_path_step(inp["hidden"], inp["alphas"], inp["logits"], inp["in_len"],  # noqa: E731
                                                    inp["targets"], inp["noise"], g_fired)
        kind = "port"

    def run():
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        if device != "cpu":
            torch.cuda.synchronize()
        return time.perf_counter() - t0, r
    return run, kind


def _ref_src():
    from oracle import ref_loader
    return ref_loader.ref_src() or ""


def run_reference(args, w, rank, world, device="cpu"):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gpu = device != "cpu"
    if gpu and not torch.cuda.is_available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "no CUDA device"}), flush=True)
        return
    sample = args.cpu_sample if not gpu else max(args.cpu_sample, 32)
    run, kind = reference_step_fn(w, sample, 1236, device)
    W = max(args.warmup, 1)
    for _ in range(W):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    what = ("the reference's eager-GPU path (Python CIF loop, ATen ctc_loss, autograd)" if gpu else
            "the reference's CPU path (torch-CPU, %d threads)" % cores)
    line = {
        "impl": "reference-gpu" if gpu else "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": W, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload=args.workload, **w, per_step_sample=sample,
                       note="%s, %s; each step = %d utterances of the workload shape" % (
                           what, "unmodified modules from %s" % os.path.relpath(_ref_src(), ROOT) if kind == "reference" else
                           "op-for-op port oracle/torch_port.py (reference sources not staged)", sample)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "per_core": value / cores,
                         "sample": "%d utterances x T=%d, S=%d, V=%d, H=%d per step" % (sample, w["T"], w["S"], w["V"], w["H"])},
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

    if args.impl != "ours":
        hot = dict(WORKLOADS["cif_ctc_joint"]) if "D" in w else w      # the reference arms time the hot path
        run_reference(args, hot, rank, world, "cpu" if args.impl == "reference" else "cuda")
        return

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    for kv in args.opt:
        key, val = kv.split("=")
        pkg("_lib").set_option(key, int(val))
    if "D" in w:      # full-model workloads
        run_train(args, args.workload, rank, world, device, graph=not args.no_graph)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    lib = pkg("_lib")
    if args.ctc_chunks:
        lib.set_option("ctc_chunks", args.ctc_chunks)
    inp = make_inputs(w, device, 1236 + rank)
    n_params = cif_model_param_count(w)
    buckets = GradBuckets(n_params, device, world, args.allreduce) if (world > 1 and not args.no_allreduce) else None
    hp = HotPath(w, inp, buckets, args.allreduce_at)
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
    l0 = lib.launch_count()
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
    timed_launches = lib.launch_count() - l0
    total_ms = t_start.elapsed_time(t_end)
    check = self_check(hp)
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
        pipe = E2EPipe(w, inp, hp.g_out)
        Ke = max(3, min(K, 5))
        pipe.start()
        for _ in range(2):
            pipe.step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            losses = pipe.step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * w["B"] * Ke / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": int(losses.numel() * losses.element_size()), "steps": Ke,
               "api": "ops.cif + ops.ctc_fc_loss (ctc_fc projection fused with the CTC loss: the step starts from the encoder "
                      "outputs, like the reference's) + autograd.backward; pinned-host inputs copied every step on a copy "
                      "stream (step k+1's copies run under step k's kernels), losses read back every step",
               "last_losses": [float(x) for x in losses]}
        del pipe
        # round 1's leg for comparison: the [B,T,V] logits themselves shipped from the host every step (PCIe-bound)
        if rank == 0 and not args.no_extras:
            pipe = E2EPipe(w, inp, hp.g_out, fused=False)
            pipe.start()
            pipe.step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                pipe.step()
            torch.cuda.synchronize()
            e2e["logits_from_host"] = {"value": w["B"] * 3 / (time.perf_counter() - t0), "unit": UNIT,
                                       "h2d_bytes_per_step": pipe.h2d_bytes, "n_gpus": 1,
                                       "note": "rank 0 only: ops.ctc_loss on logits copied from the host (round-1 leg)"}
            del pipe

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
    mha = sweep = None
    extras = rank == 0 and not args.no_extras
    if extras:
        asg = assigner_microbench(w, inp, device)
        for n in ("cif_alpha_fwd", "cif_alpha_bwd"):
            kernels.append({"kernel": n, "bound": "hbm", "ms": asg[n]["ms"], "algorithmic_bytes": asg[n]["algorithmic_bytes"],
                            "GBps": asg[n]["GBps"], "frac_of_hbm_peak": asg[n]["GBps"] / peaks["hbm_gbs"],
                            "in_timed_step": False, "note": "SURVEY 8(f2): assigner tail + alpha scaling, next-row kernel"})
        fc = ctc_fc_microbench(w, inp, device, peaks)
        for n, v in fc.items():
            kernels.append(dict(kernel=n, bound="tensor (tf32 x3, shared-memory bandwidth)", in_timed_step=False,
                                shape="M=%d (B x T) K=%d N=%d f32" % (w["B"] * w["T"], w["H"], w["V"]), **v))
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
    del hp, inp
    torch.cuda.empty_cache()
    if extras:
        sa = spec_aug_microbench(device)
        kernels.append({"kernel": "spec_aug", "bound": "hbm", "ms": sa["ms"], "algorithmic_bytes": sa["algorithmic_bytes"],
                        "GBps": sa["GBps"], "frac_of_hbm_peak": sa["GBps"] / peaks["hbm_gbs"], "shape": sa["shape"],
                        "in_timed_step": False, "note": "SURVEY 8(f4): SpecAugment, three launches incl. the means"})
        lnb = layer_norm_microbench(device)
        for n in ("ln_fwd", "ln_bwd"):
            kernels.append({"kernel": n, "bound": "hbm", "ms": lnb[n]["ms"], "algorithmic_bytes": lnb[n]["algorithmic_bytes"],
                            "GBps": lnb[n]["GBps"], "frac_of_hbm_peak": lnb[n]["GBps"] / peaks["hbm_gbs"], "shape": lnb["shape"],
                            "in_timed_step": False,
                            "note": "SURVEY 8(f3): dropout + residual + LayerNorm (+ non-pad mask) in one kernel each way; "
                                    "the backward figure includes the column sum of the per-CTA gamma / beta partial rows"})
        lin = linear_microbench(device)
        for n in ("linear_bias_relu", "linear_residual_layernorm", "linear_residual_layernorm_one_kernel", "linear_f32_3xtf32"):
            kernels.append({"kernel": n, "bound": "tensor", "ms": lin[n]["ms"], "TFLOPs": lin[n]["TFLOPs"],
                            "frac_of_bf16_peak": lin[n]["TFLOPs"] / peaks["bf16_tflops"], "shape": lin[n]["shape"],
                            "torch_cublas_plus_eager_TFLOPs": lin[n]["torch_TFLOPs"], "in_timed_step": False,
                            "note": "SURVEY 8(f3): tcgen05 GEMM with the epilogue fused, next-row kernel (linear_residual_layernorm: persistent GEMM + "
                                    "bf16 LayerNorm kernel; _one_kernel: the fp32 row normalised inside tensor memory)"})
        mha = attention_microbench(device, peaks)
        sweep = ctc_sweep(device, peaks)
    mha_roofline = None
    if mha is not None:
        top = mha["shapes"][0]
        mha_roofline = {
            "fwd": {"bound": "tensor", "kernel": "asr::mha_fwdp_kernel<false> (persistent two-tile forward)", "achieved": top["fwd_TFLOPs"],
                    "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": top["fwd_frac_of_bf16_peak"],
                    "avg_launch_ms": top["fwd_ms"], "shape": top["shape"], "flops": "4 B h Lq Lk d"},
            "bwd": {"bound": "tensor", "kernel": "asr::mha_bwdp_kernel<false> (persistent; + delta and dQ-convert kernels)",
                    "achieved": top["bwd_TFLOPs"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": top["bwd_frac_of_bf16_peak"], "avg_launch_ms": top["bwd_ms"], "shape": top["shape"],
                    "flops": "10 B h Lq Lk d"},
            "peak_source": peak_src, "other_shapes": mha["shapes"][1:], "model_shapes": mha["model_shapes"]}

    # ---- the whole models next to the hot-path number (same launch, a few steps each) --------------------
    train_step = transformer_step = None
    if not args.no_extras:
        def brief(tl, **extra):
            d = {k: tl[k] for k in ("value", "unit", "ms_per_step", "e2e", "gpu_launches", "dtype")}
            d.update(workload=tl["config"]["workload"], model=tl["config"]["model"], params=tl["config"]["params"],
                     grad_allreduce_bytes=tl["config"]["grad_allreduce_bytes"], launch=tl["config"]["launch"],
                     parallelism=tl["config"]["parallelism"], **extra)
            return d
        tl = run_train(args, "train", rank, world, device, steps=10, emit=False, graph=not args.no_graph)
        te = run_train(args, "train", rank, world, device, steps=5, emit=False, graph=False)
        t2 = run_train(args, "train", rank, world, device, steps=10, emit=False, tf32=True, graph=not args.no_graph)
        if tl is not None:
            train_step = brief(tl, per_gpu=dict(WORKLOADS["train"]))
            train_step["eager_launches"] = {"value": te["value"], "ms_per_step": te["ms_per_step"],
                                            "note": "same step without the CUDA graph (round-1 configuration: host-bound)"}
            train_step["with_tf32_matmul"] = {"value": t2["value"], "unit": t2["unit"], "ms_per_step": t2["ms_per_step"],
                                              "note": "same step with torch.backends.*.allow_tf32 = True for the model shell's "
                                                      "Linear / Conv GEMMs (reduced precision: informational, not the reported value)"}
        tt = run_train(args, "transformer_bf16", rank, world, device, steps=10, emit=False, graph=not args.no_graph)
        tc = run_train(args, "transformer_bf16", rank, world, device, steps=10, emit=False, graph=not args.no_graph, torch_linear=True)
        tf = run_train(args, "train", rank, world, device, steps=10, emit=False, graph=not args.no_graph, torch_linear=True)
        if tt is not None:
            transformer_step = brief(tt, per_gpu=dict(WORKLOADS["transformer_bf16"]))
            transformer_step["with_torch_linear"] = {"value": tc["value"], "ms_per_step": tc["ms_per_step"],
                                                     "note": "same step with the shell's Linear layers on torch's F.linear (cuBLAS bf16) "
                                                             "instead of csrc/gemm2.cu"}
            train_step["with_torch_linear"] = {"value": tf["value"], "ms_per_step": tf["ms_per_step"],
                                               "note": "same step with the shell's Linear layers on torch's F.linear (cuBLAS SIMT sgemm) "
                                                       "instead of csrc/gemm2.cu"}

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            run, kind = reference_step_fn(w, args.cpu_sample, 1236)
            run()
            dt_cpu, _ = run()
            cpu_baseline = {"value": args.cpu_sample / dt_cpu, "unit": UNIT, "cores": cores, "kind": kind,
                            "per_core": args.cpu_sample / dt_cpu / cores,
                            "sample": "%d utterances x T=%d, S=%d, V=%d, H=%d, one step after one warm-up (%.1f s)" % (
                                args.cpu_sample, w["T"], w["S"], w["V"], w["H"], dt_cpu)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload=args.workload, **w, L=L_out, valid_frames=valid_frames,
                           grad_allreduce_bytes_per_step=(buckets.bytes if buckets is not None else 0),
                           grad_allreduce=(buckets.how + ", launched behind the CTC row pass, waited for at the end of every timed step")
                                          if buckets is not None else "none at N = 1 (one rank: nothing to exchange)",
                           parallelism="dp%d by utterance; gradients only over NVLink" % world,
                           schedule="serial, one stream" if args.serial else
                           "one stream; CTC begin (rows + lattices on library streams) / CIF pair (forward: kernel hint 3) / CTC finish (apply)",
                           l2="inputs (%.1f GB logits + %.1f GB hidden per GPU) exceed the 126 MB L2; no flush needed" % (
                               w["B"] * w["T"] * w["V"] * 4 / 1e9, w["B"] * w["T"] * w["H"] * 4 / 1e9)),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(timed_launches), "self_check": check,
            "roofline": roofline, "mha_roofline": mha_roofline, "kernels": kernels, "ctc_sweep": sweep,
            "cpu_baseline": cpu_baseline, "train_step": train_step, "transformer_step": transformer_step,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
