#!/usr/bin/env python3
"""Error of each lattice variant against the fp64 oracle on one test shape: python tools/ctc_case_probe.py B T V S seed"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200, oracle
from helpers import make_ctc_inputs, to_np
B, T, V, S, seed = (int(x) for x in sys.argv[1:6])
logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=seed)
o_loss, o_nll, o_grad = oracle.ctc_loss_and_grad(to_np(logits), to_np(targets), to_np(in_len))
fin = np.isfinite(o_nll)
print("in_len", in_len.tolist(), "tgt_len", targets.ne(0).sum(1).tolist(), "oracle nll", o_nll)
for variant in (0, 1):
    asr_b200._lib.set_option("ctc_lattice_variant", variant)
    lg = logits.clone().requires_grad_(True)
    loss, nll = asr_b200.ops.ctc_loss(lg, in_len, targets, return_nll=True)
    loss.backward()
    g = to_np(lg.grad)
    gs = np.nanmax(np.abs(o_grad[fin]))
    err = np.abs(g[fin] - o_grad[fin])
    bi = np.unravel_index(np.argmax(err), err.shape)
    print("variant", variant, "nll", to_np(nll), "grad err/scale %.3e" % (err.max() / gs), "at", bi,
          "per-frame max err", np.round(err.max(axis=2)[0][:: max(1, T // 12)] / gs, 7))
