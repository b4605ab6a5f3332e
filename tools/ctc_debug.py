import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200
from helpers import make_ctc_inputs
ops = asr_b200.ops
B, T, S, V = 64, 400, 20, 4233
logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=1236)
lg = logits.clone().requires_grad_(True)
loss, nll = ops.ctc_loss(lg, in_len, targets, return_nll=True); loss.backward(); g = lg.grad
lg2 = logits.double().clone().requires_grad_(True)
tl = targets.ne(0).int().sum(1)
lp = F.log_softmax(lg2, dim=-1).transpose(0, 1)
l2 = F.ctc_loss(lp, targets, in_len, tl, blank=V - 1); l2.backward(); d = lg2.grad
err = (g.double() - d).abs()
gs = d.abs().max().item()
print("gscale", gs, "max err", err.max().item())
per_b = err.amax(dim=(1, 2))
bad = (per_b > 1e-4 * gs).nonzero().flatten().tolist()
print("bad utterances", bad)
for b in bad[:4]:
    eb = err[b]
    tt = eb.amax(dim=1)
    rows = (tt > 1e-4 * gs).nonzero().flatten().tolist()
    print("b", b, "in_len", int(in_len[b]), "tgt_len", int(tl[b]), "targets", targets[b].tolist())
    print("  bad rows", rows[:10], "... n=", len(rows))
    for t in rows[:3]:
        cols = (eb[t] > 1e-4 * gs).nonzero().flatten().tolist()
        print("   t", t, "cols", cols, "ours", g[b, t, cols].tolist(), "ref", d[b, t, cols].tolist())
