"""Drop-in for /root/reference/src/transformer/attention.py.

`MultiheadAttention(d_model, n_head, d_k=64, d_v=64, dropout=0.1)` keeps the
reference's parameters, initialisation and state_dict keys (`w_qs, w_ks, w_vs, fc,
layer_norm`), and `forward(q, k, v, mask=None) -> (output, attn)`.

What changes: the scaled-dot-product core (bmm / scale / masked_fill / softmax / bmm
and the three head-major permute copies, reference :47-57 and :74-86) is one
tcgen05 kernel working in bf16 with fp32 accumulation directly on the
[B, L, heads, 64] projection outputs.  The core never materialises the attention
probabilities; the second return value `attn` ([(heads*B), Lq, Lk] fp32, head-major
rows like attention.py:47,62) comes from a separate kernel.  It is returned by default,
like the reference.  Every training caller discards it (encoder.py:72,
decoder.py:628-633), so the shells of this package build their layers with
`return_attn=False` (attn = None, no second kernel) and `patch.install()` turns the
class default off the same way.  In evaluation mode `attn` equals the reference's; in
training mode the reference returns the probabilities AFTER dropout (attention.py:83-86)
while this kernel returns them before dropout.

Dropout on the probabilities (reference :83) happens inside the kernel in training
mode: a counter-based generator keyed by a per-call seed (drawn from torch's CPU
generator, so `torch.manual_seed` reproduces a run) decides per (b, head, q, k), and
backward regenerates the same mask.  The rate is quantised to 1/256 (0.1 -> 26/256)
with the exactly matching rescale, so the expectation is unbiased.
"""
import numpy as np
import torch
import torch.nn as nn

from ..ops import mha_core, mha_probs, linear_act, linear_residual_layernorm
from .module import fused_linear_ok, Linear, dropout_residual_layer_norm


class MultiheadAttention(nn.Module):
    ''' Multi-Head Attention module (same parameters as the reference) '''

    RETURN_ATTN_DEFAULT = True      # class-wide default of `return_attn` (patch.install() turns it off)

    def __init__(self, d_model, n_head, d_k=64, d_v=64, dropout=0.1, return_attn=None):
        super().__init__()
        if d_k != 64 or d_v != 64:
            raise ValueError("the sm_100a attention core is built for d_k = d_v = 64 (every reference recipe)")
        self.n_head = n_head
        self.d_k = d_k
        self.d_v = d_v
        self.return_attn = return_attn

        self.w_qs = Linear(d_model, n_head * d_k)
        self.w_ks = Linear(d_model, n_head * d_k)
        self.w_vs = Linear(d_model, n_head * d_v)
        nn.init.normal_(self.w_qs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))

        self.attn_dropout_p = dropout
        self.temperature = np.power(d_k, 0.5)
        self.layer_norm = nn.LayerNorm(d_model)

        self.fc = Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)

        self.dropout = nn.Dropout(dropout)

    def forward(self, q, k, v, mask=None, kv_len=None, causal=False, out_scale=None):
        """q [B,Lq,d_model], k,v [B,Lk,d_model], mask [B,Lq,Lk] (True = masked) or None.
        `kv_len` / `causal` are optional structured forms of the reference's masks
        (get_attn_pad_mask / get_subsequent_mask); they are OR-ed with `mask`.
        `out_scale` [B,Lq,1] (optional): the non-pad mask the calling layer multiplies the output with
        (encoder.py:76, decoder.py:628-632), folded into the LayerNorm kernel on the training path."""
        n_head, d_k, d_v = self.n_head, self.d_k, self.d_v
        sz_b, len_q, _ = q.size()
        len_k = k.size(1)

        residual = q
        fused = fused_linear_ok(self, q, self.w_qs, self.w_ks, self.w_vs, self.fc) and k.dtype == q.dtype and v.dtype == q.dtype
        if fused:      # evaluation in bf16: the projections as tcgen05 GEMMs with the bias in the epilogue
            qh = linear_act(q, self.w_qs.weight, self.w_qs.bias).view(sz_b, len_q, n_head, d_k)
            kh = linear_act(k, self.w_ks.weight, self.w_ks.bias).view(sz_b, len_k, n_head, d_k)
            vh = linear_act(v, self.w_vs.weight, self.w_vs.bias).view(sz_b, len_k, n_head, d_v)
        else:
            qi, ki, vi = q, k, v
            bf16_autocast = (q.is_cuda and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16)
            if bf16_autocast and q.dtype == torch.float32:
                # autocast casts every Linear input on its own: a self-attention layer would convert the same activations
                # three times (and its backward convert three gradients back); once is enough
                from ..ops import bf16_copy_of
                cast = lambda t: (lambda c: t.to(torch.bfloat16) if c is None else c)(bf16_copy_of(t))  # noqa: E731
                qi = cast(q)
                ki = qi if k is q else cast(k)
                vi = ki if v is k else (qi if v is q else cast(v))
            qh = self.w_qs(qi).view(sz_b, len_q, n_head, d_k)
            kh = self.w_ks(ki).view(sz_b, len_k, n_head, d_k)
            vh = self.w_vs(vi).view(sz_b, len_k, n_head, d_v)

        p_attn = self.attn_dropout_p if self.training else 0.0
        ctx = mha_core(qh, kh, vh, kv_len=kv_len, mask=mask, causal=causal, scale=1.0 / float(self.temperature),
                       dropout_p=p_attn)
        attn = None
        if self.RETURN_ATTN_DEFAULT if self.return_attn is None else self.return_attn:
            attn = mha_probs(qh, kh, kv_len=kv_len, mask=mask, causal=causal, scale=1.0 / float(self.temperature))

        output = ctx.reshape(sz_b, len_q, n_head * d_v)
        if fused:      # fc + bias + residual + LayerNorm in one kernel (the 512-wide row stays in tensor memory)
            out = linear_residual_layernorm(output.to(q.dtype), self.fc.weight, self.fc.bias, residual, self.layer_norm.weight,
                                            self.layer_norm.bias, self.layer_norm.eps)
            return (out if out_scale is None else out * out_scale), attn
        if not (q.is_cuda and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16):
            output = output.to(q.dtype)       # under bf16 autocast the projection takes the kernel's bf16 output as it is
        return dropout_residual_layer_norm(self.layer_norm, self.dropout, self.fc(output), residual, out_scale), attn
