#!/usr/bin/env python3
"""Where the full-model training step spends its time: python tools/train_profile.py  (torch profiler, 3 steps)"""
import argparse, importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from torch.profiler import profile, ProfilerActivity
cm = importlib.import_module("end-to-end_asr_pytorch_b200.transformer.cif_model")
lossm = importlib.import_module("end-to-end_asr_pytorch_b200.transformer.loss")
w = {"B": 64, "T": 167, "S": 14, "V": 4233, "H": 512, "D": 320}
dev = torch.device("cuda")
torch.manual_seed(1234)
model = cm.CIF_Model.create_model(bench._model_args(w)).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=2e-4, betas=(0.9, 0.98), eps=1e-9)
feats, lens, targets = bench._train_inputs(w, dev, 1240)

def step():
    opt.zero_grad(set_to_none=True)
    ctc_logits, len_ctc, _num, num, logits = model(feats, lens, targets)
    qua, ctc, ce = lossm.cal_ctc_qua_ce_loss(ctc_logits, len_ctc, _num, num, logits, targets, smoothing=0.1)
    (0.001 * qua + ctc + ce).backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize(); print("wall per step %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
    for _ in range(3): step()
    torch.cuda.synchronize()
ka = p.key_averages()
cuda_total = sum(e.self_device_time_total for e in ka) / 3e3
print("GPU busy per step %.2f ms" % cuda_total)
print(ka.table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=70))
