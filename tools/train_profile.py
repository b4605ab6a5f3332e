#!/usr/bin/env python3
"""Where a full-model training step spends its GPU time (torch profiler, eager launches, 3 steps):
    python tools/train_profile.py [train|transformer_bf16] [--torch-linear] [--fused-ctc]"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from torch.profiler import profile, ProfilerActivity
wname = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "train"
mod = bench.pkg("transformer.module")
if "--torch-linear" in sys.argv:
    mod.USE_TENSOR_CORE_FP32 = mod.USE_TENSOR_CORE_BF16 = False
lossm = bench.pkg("transformer.loss")
w = dict(bench.WORKLOADS[wname])
bf16 = wname == "transformer_bf16"
dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True          # as bench.py runs the step
torch.manual_seed(1234)
if bf16:
    model = bench.pkg("transformer.transformer").Transformer.create_model(bench._model_args(w)).to(dev).train()
else:
    model = bench.pkg("transformer.cif_model").CIF_Model.create_model(bench._model_args(w)).to(dev).train()
    model.fused_ctc_fc = "--fused-ctc" in sys.argv
opt = torch.optim.Adam(model.parameters(), lr=2e-4, betas=(0.9, 0.98), eps=1e-9, fused=True)
feats, lens, targets = bench._train_inputs(w, dev, 1240)

def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
        if bf16:
            logits, targets_eos = model(feats, lens, targets)
            loss = lossm.cal_ce_loss(logits.float(), targets_eos, smoothing=0.1)
        else:
            ctc_logits, len_ctc, _num, num, logits = model(feats, lens, targets)
            qua, ctc, ce = lossm.cal_ctc_qua_ce_loss(ctc_logits, len_ctc, _num, num, logits, targets, smoothing=0.1)
            loss = 0.001 * qua + ctc + ce
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize(); print("%s %s: wall per eager step %.2f ms" % (wname, sys.argv[2:], (time.perf_counter() - t0) / 5 * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
    for _ in range(3): step()
    torch.cuda.synchronize()
ka = p.key_averages()
# kernels only (device_type CUDA events), per step
from torch.autograd import DeviceType
rows = [(e.key, e.self_device_time_total / 3e3, e.count / 3) for e in ka if e.device_type == DeviceType.CUDA]
rows.sort(key=lambda r: -r[1])
total = sum(r[1] for r in rows)
print("GPU busy per step %.2f ms in %.0f launches" % (total, sum(r[2] for r in rows)))
ours = sum(r[1] for r in rows if "asr::" in r[0])
print("this package's kernels: %.2f ms (%.0f %%)" % (ours, 100 * ours / total))
for name, ms, n in rows[:45]:
    print("%8.3f ms %5.1f %% %6.1f x %7.1f us  %s" % (ms, 100 * ms / total, n, ms / n * 1e3, name[:110]))
