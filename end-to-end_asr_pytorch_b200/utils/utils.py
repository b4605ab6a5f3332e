"""Attention-mask builders - drop-in for /root/reference/src/utils/utils.py:125-165.

Integer / boolean host-side glue (row a7 of SURVEY.md 8a).  The fused attention
kernel takes (kv_len, causal) instead of a dense mask when the mask is one of
these forms; `mask_to_kv_len` recognises them.
"""
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
