"""Caller of the attention hot path: Transformer encoder (plain PyTorch around the
tcgen05 attention core).  Names follow /root/reference/src/transformer/encoder.py
(`linear_in`, `layer_norm_in`, `layer_stack.N.{slf_attn,pos_ffn}`)."""
import torch.nn as nn

from .module import Linear, dropout_residual_layer_norm

from .attention import MultiheadAttention
from .module import PositionalEncoding, PositionwiseFeedForward
from ..utils.utils import sequence_mask


class EncoderLayer(nn.Module):
    """Self-attention + position-wise FFN, each followed by the non-pad mask."""

    def __init__(self, d_model, d_inner, n_head, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiheadAttention(d_model, n_head, dropout=dropout, return_attn=False)      # every caller here discards attn
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, dropout=dropout)

    def forward(self, enc_input, non_pad_mask=None, slf_attn_mask=None, kv_len=None, causal=False):
        out, _ = self.slf_attn(enc_input, enc_input, enc_input, mask=slf_attn_mask, kv_len=kv_len, causal=causal,
                               out_scale=non_pad_mask)
        return self.pos_ffn(out, out_scale=non_pad_mask)


class Encoder(nn.Module):
    def __init__(self, d_input, n_layers, n_head, d_model, d_inner, dropout=0.1):
        super().__init__()
        self.d_input, self.n_layers, self.n_head = d_input, n_layers, n_head
        self.d_model = self.d_output = d_model
        self.d_inner, self.dropout_rate = d_inner, dropout
        self.linear_in = Linear(d_input, d_model)
        self.layer_norm_in = nn.LayerNorm(d_model)
        self.positional_encoding = PositionalEncoding(d_model)
        self.dropout = nn.Dropout(dropout)
        self.layer_stack = nn.ModuleList([EncoderLayer(d_model, d_inner, n_head, dropout=dropout)
                                          for _ in range(n_layers)])

    def forward(self, padded_input, input_lengths):
        """N x T x D, N -> N x T x d_model.  The key-padding mask of the reference
        (get_attn_pad_mask, utils.py:157-165) is handed to the attention kernel in its
        structured form, as per-utterance key lengths."""
        x = self.dropout(dropout_residual_layer_norm(self.layer_norm_in, None, self.linear_in(padded_input)) + self.positional_encoding(padded_input))
        # in the activations' dtype: a float32 mask would promote bf16 activations (and break the fused bf16 layers)
        non_pad_mask = sequence_mask(input_lengths, padded_input.size(1), dtype=x.dtype).unsqueeze(-1)
        for layer in self.layer_stack:
            x = layer(x, non_pad_mask=non_pad_mask, kv_len=input_lengths)
        return x
