"""Attention-mask builders oracle (numpy).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates /root/reference/src/utils/utils.py:125-165.  All masks are integer /
boolean work, so parity is bit-exact.  Convention: True (1) = masked position.
"""
import numpy as np


def sequence_mask(lengths, maxlen=None, dtype=np.float32):
    """utils.py:125-133: ones.cumsum(1) <= lengths[:,None]  -> 1 on valid frames."""
    lengths = np.asarray(lengths)
    if maxlen is None:
        maxlen = int(lengths.max())
    pos = np.arange(1, maxlen + 1)[None, :]
    return (pos <= lengths[:, None]).astype(dtype)


def get_subsequent_mask(seq):
    """utils.py:136-144: strict upper triangle, broadcast over the batch."""
    seq = np.asarray(seq)
    B, L = seq.shape
    tri = np.triu(np.ones((L, L), dtype=np.uint8), k=1)
    return np.broadcast_to(tri[None], (B, L, L)).copy()


def get_attn_key_pad_mask(seq_k, seq_q, pad_idx):
    """utils.py:147-154: key positions with token id <= pad_idx, repeated over Lq."""
    seq_k = np.asarray(seq_k)
    len_q = np.asarray(seq_q).shape[1]
    pad = seq_k <= pad_idx
    return np.broadcast_to(pad[:, None, :], (seq_k.shape[0], len_q, seq_k.shape[1])).copy()


def get_attn_pad_mask(input_lengths, expand_length):
    """utils.py:157-165: True at key frames >= length, repeated over expand_length queries."""
    valid = sequence_mask(input_lengths)
    pad = valid < 1.0
    return np.broadcast_to(pad[:, None, :], (pad.shape[0], expand_length, pad.shape[1])).copy()
