"""Shared helpers for the GPU parity tests: package import + the seeded synthetic
input generators of SURVEY.md 8(d)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "end-to-end_asr_pytorch_b200"


def pkg(sub=None):
    return importlib.import_module(PKG if sub is None else PKG + "." + sub)


def gen(seed):
    return torch.Generator().manual_seed(seed)


def make_cif_inputs(B, T, H, n_labels, seed, ragged=True, device="cuda"):
    """cfg-4 style: alphas = sigmoid(N(0,1)) rescaled per row to sum to U_b ~ U{2/3 n .. n},
    zero beyond a ragged length ~ U[0.8T, T]."""
    g = gen(seed)
    hidden = torch.randn(B, T, H, generator=g)
    alphas = torch.sigmoid(torch.randn(B, T, generator=g))
    if ragged:
        lens = torch.randint(max(1, int(0.8 * T)), T + 1, (B,), generator=g)
        alphas = alphas * (torch.arange(T)[None, :] < lens[:, None]).float()
    target = torch.randint(max(1, (2 * n_labels) // 3), n_labels + 1, (B,), generator=g).float()
    noise = torch.rand(B, generator=g) - 0.5
    alphas = alphas * ((target + noise) / alphas.sum(-1))[:, None]
    return hidden.to(device), alphas.to(device)


def make_ctc_inputs(B, T, V, S, seed, device="cuda", full_len=False):
    """cfg-2 style: logits ~ N(0,1), targets randint(1, V-1) with ~10% forced repeats,
    input_lengths ~ U[0.6T, T], target_lengths ~ U[0.5S, S] (0-padded), blank = V-1."""
    g = gen(seed)
    logits = torch.randn(B, T, V, generator=g)
    targets = torch.randint(1, V - 1, (B, S), generator=g)
    rep = torch.rand(B, S, generator=g) < 0.1
    for s in range(1, S):
        targets[:, s] = torch.where(rep[:, s], targets[:, s - 1], targets[:, s])
    if full_len:
        in_len = torch.full((B,), T, dtype=torch.int32)
        tgt_len = torch.full((B,), S, dtype=torch.int64)
    else:
        in_len = torch.randint(int(0.6 * T), T + 1, (B,), generator=g).to(torch.int32)
        tgt_len = torch.randint(max(1, S // 2), S + 1, (B,), generator=g)
    targets = targets * (torch.arange(S)[None, :] < tgt_len[:, None]).long()
    return logits.to(device), targets.to(device), in_len.to(device)


def to_np(t):
    return t.detach().cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)
