"""Caller: the CIF weight predictor (plain PyTorch).  Names follow
/root/reference/src/transformer/attentionAssigner.py (`conv.conv.assigner/conv1d_i`, `linear`).
`forward_scaled` is the fused form of the tail (linear -> sigmoid -> pad mask) together with
CIF_Model.forward's scaling glue: one sm_100a kernel pass over the activations (ops.cif_alpha)."""
import torch
import torch.nn as nn

from .conv_encoder import Conv1d
from ..ops import cif_alpha
from ..utils.utils import sequence_mask


class Attention_Assigner(nn.Module):
    """Conv1d stack -> dropout -> Linear(.,1) -> sigmoid, zero on padded frames."""

    def __init__(self, d_input, d_hidden, w_context, n_layers, dropout=0.1):
        super().__init__()
        self.d_input, self.d_hidden, self.n_layers, self.w_context = d_input, d_hidden, n_layers, w_context
        self.conv = Conv1d(d_input, d_hidden, n_layers, w_context, pad='same', name='assigner')
        self.dropout = nn.Dropout(p=dropout)
        self.linear = nn.Linear(d_hidden, 1)

    def forward(self, padded_input, input_lengths):
        x, input_lengths = self.conv(padded_input, input_lengths)
        alphas = torch.sigmoid(self.linear(self.dropout(x)).squeeze(-1))
        return alphas * sequence_mask(input_lengths, padded_input.size(1))

    def forward_scaled(self, padded_input, input_lengths, num_noise=None):
        """(alpha [B,T] scaled so that it sums to num_noise, _num [B] = sum of the unscaled weights);
        attentionAssigner.py:36-40 + cif_model.py:43-48 in one kernel pass."""
        x, input_lengths = self.conv(padded_input, input_lengths)
        return cif_alpha(self.dropout(x), self.linear.weight, self.linear.bias, input_lengths, num_noise)
