"""Attention-mask builders and SpecAugment - drop-in for /root/reference/src/utils/utils.py:125-194.

Integer / boolean host-side glue (row a7 of SURVEY.md 8a).  The fused attention
kernel takes (kv_len, causal) instead of a dense mask when the mask is one of
these forms; `mask_to_kv_len` recognises them.  `spec_aug` (row f4) draws its masks
exactly like the reference and applies them with the sm_100a kernels of csrc/specaug.cu.
"""
import importlib

import torch


def sequence_mask(lengths, maxlen=None, dtype=torch.float):
    """1 on frames < length (reference utils.py:125-133)."""
    if maxlen is None:
        maxlen = lengths.max()
    steps = torch.arange(1, int(maxlen) + 1, device=lengths.device)
    return (steps.unsqueeze(0) <= lengths.unsqueeze(1)).type(dtype)


def get_subsequent_mask(seq):
    """Strict upper triangle, uint8, [B,L,L] (reference utils.py:136-144)."""
    sz_b, len_s = seq.size()
    tri = torch.triu(torch.ones((len_s, len_s), device=seq.device, dtype=torch.uint8), diagonal=1)
    return tri.unsqueeze(0).expand(sz_b, -1, -1)


def get_attn_key_pad_mask(seq_k, seq_q, pad_idx):
    """True at keys whose token id <= pad_idx, [B,Lq,Lk] (reference utils.py:147-154)."""
    len_q = seq_q.size(1)
    return seq_k.le(pad_idx).unsqueeze(1).expand(-1, len_q, -1)


def get_attn_pad_mask(input_lengths, expand_length):
    """True at key frames >= length, [B,expand_length,Lk] (reference utils.py:157-165)."""
    pad = sequence_mask(input_lengths) < 1.0
    return pad.unsqueeze(1).expand(-1, expand_length, -1)


def spec_aug_draw(B, V, feature_lengths, config, device):
    """The random draws of the reference's spec_aug (utils.py:170,176-181,185-189), same torch.rand calls in
    the same order, hence the same masks for the same generator state.  Returns (f0, fw, t0, tw), each [R,B]
    int64 with R = time_mask_num - the reference loops `time_mask_num` times for BOTH families (utils.py:176)
    and never reads freq_mask_num; that is kept."""
    freq_mask_num, freq_mask_width, time_mask_num, time_mask_width = (int(i) for i in config.split('-'))
    f0, fw, t0, tw = [], [], [], []
    for _ in range(time_mask_num):
        fs = (freq_mask_width * torch.rand(size=[B], device=device, requires_grad=False)).long()
        f0s = ((V - fs).float() * torch.rand(size=[B], device=device, requires_grad=False)).long()
        fw.append(fs)
        f0.append(f0s)
    for _ in range(time_mask_num):
        ts = (time_mask_width * torch.rand(size=[B], device=device, requires_grad=False)).long()
        t0s = ((feature_lengths - ts).float() * torch.rand(size=[B], device=device, requires_grad=False)).long()
        tw.append(ts)
        t0.append(t0s)
    empty = torch.zeros((0, B), dtype=torch.long, device=device)
    stack = lambda xs: torch.stack(xs) if xs else empty
    return stack(f0), stack(fw), stack(t0), stack(tw)


def spec_aug(padded_features, feature_lengths, config):
    """Drop-in for the reference's spec_aug (utils.py:168-194): masks `padded_features` [B,T,V] IN PLACE and
    returns (padded_features, feature_lengths).  CUDA tensors only - there is no CPU fallback."""
    B, T, V = padded_features.shape
    f0, fw, t0, tw = spec_aug_draw(B, V, feature_lengths, config, padded_features.device)
    ops = importlib.import_module(__name__.rsplit(".", 2)[0] + ".ops")
    if padded_features.is_contiguous():
        ops.spec_aug_apply(padded_features, feature_lengths, f0, fw, t0, tw)
    else:
        padded_features.copy_(ops.spec_aug_apply(padded_features.contiguous(), feature_lengths, f0, fw, t0, tw))
    return padded_features, feature_lengths
