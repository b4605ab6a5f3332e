"""Caller-side building blocks (plain PyTorch; not on the kernel hot path).

Same parameter / buffer names as /root/reference/src/transformer/module.py so that
checkpoints interchange: `PositionalEncoding.pe`, `PositionwiseFeedForward.{w_1,w_2,layer_norm}`.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# Linear layers of the model shell on this package's tensor-core GEMMs (csrc/gemm2.cu), forward AND backward, operands as
# torch stores them (no transposed copies):
#   fp32 (the reference's precision): three TF32 products per K step, ~5e-6 relative error; torch's own fp32 path is
#        cuBLAS's SIMT sgemm, ~45 % of the GPU time of the training step on B200.
#   bf16 under torch.autocast(bf16) (BASELINE config 3): bias / ReLU in the epilogue, fp32 weight gradients.
# Both switches can be turned off to fall back to torch's F.linear (cuBLAS), e.g. for A/B measurements.
USE_TENSOR_CORE_FP32 = True
USE_TENSOR_CORE_BF16 = True
# LayerNorm(dropout(y) + residual) at the end of every sub-layer as ONE kernel each way (csrc/ln.cu) while gradients are
# recorded on CUDA; off: nn.Dropout, +, nn.LayerNorm as the reference runs them.
USE_FUSED_LAYER_NORM = True


def dropout_residual_layer_norm(layer_norm, dropout, y, residual=None, row_scale=None):
    """layer_norm(dropout(y) + residual) [* row_scale] - module.py:50-52 / attention.py:59-60 / encoder.py:49 of the reference,
    and the non-pad mask the layers apply next (encoder.py:76-80) - with the parameters of the given nn.LayerNorm /
    nn.Dropout (dropout may be None).  row_scale: [..., 1] or None.
    Under bf16 autocast the returned fp32 tensor carries a bf16 copy written by the same kernel (ops.bf16_copy_of), which
    the shell's Linear / MultiheadAttention read instead of converting it again.  The copy belongs to that tensor OBJECT: any
    out-of-place op yields a tensor without it; do not modify the returned tensor in place before handing it to a Linear."""
    if USE_FUSED_LAYER_NORM and y.is_cuda and torch.is_grad_enabled() and layer_norm.elementwise_affine:
        from .. import ops
        yc = y.contiguous()
        rc = residual.contiguous() if residual is not None else None
        if len(layer_norm.normalized_shape) == 1 and ops.residual_layer_norm_available(yc, rc, layer_norm.weight):
            p = dropout.p if (dropout is not None and dropout.training) else 0.0
            # under bf16 autocast the consumer is a bf16 GEMM: let the kernel write the bf16 copy it will read
            shadow = torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
            return ops.residual_layer_norm(yc, rc, layer_norm.weight, layer_norm.bias, layer_norm.eps, p, True, row_scale=row_scale,
                                           bf16_copy=shadow)
    if dropout is not None:
        y = dropout(y)
    out = layer_norm(y + residual if residual is not None else y)
    return out if row_scale is None else out * row_scale


class Linear(nn.Linear):
    """nn.Linear (same parameters, same state_dict keys) whose CUDA forward and backward run on the GEMMs of this package.
    `relu=True` fuses max(., 0) (feed-forward block, module.py:50)."""

    def forward(self, x, relu=False):
        if x.is_cuda and x.numel() > 0 and self.weight.dtype == torch.float32:
            from .. import ops
            if torch.is_autocast_enabled("cuda"):
                if (USE_TENSOR_CORE_BF16 and torch.get_autocast_dtype("cuda") == torch.bfloat16
                        and ops.linear_bf16_ok(x, self.weight)):
                    xb = ops.bf16_copy_of(x)          # written by the LayerNorm kernel that produced x
                    return ops.linear_bf16_autograd(x if xb is None else xb, self.weight, self.bias, relu=relu)
            elif USE_TENSOR_CORE_FP32 and x.dtype == torch.float32 and ops.linear_f32_ok(x, self.weight):
                y = ops.linear_f32_autograd(x, self.weight, self.bias)
                return F.relu(y) if relu else y
        y = F.linear(x, self.weight, self.bias)
        return F.relu(y) if relu else y


class PositionalEncoding(nn.Module):
    """Sinusoidal table, returned for the first T positions of the input."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
        inv_freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(pos * inv_freq)
        table[:, 1::2] = torch.cos(pos * inv_freq)
        self.register_buffer('pe', table.unsqueeze(0))

    def forward(self, input):
        return self.pe[:, :input.size(1)]


class PositionwiseFeedForward(nn.Module):
    """LayerNorm(x + W2 relu(W1 x))."""

    def __init__(self, d_model, d_ff, dropout=0.1):
        super().__init__()
        self.w_1 = Linear(d_model, d_ff)
        self.w_2 = Linear(d_ff, d_model)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(d_model)

    def forward(self, x, out_scale=None):
        """out_scale [..., 1] (optional): the non-pad mask the calling layer multiplies the result with, folded into the
        LayerNorm kernel on the training path."""
        if fused_linear_ok(self, x, self.w_1, self.w_2):
            # evaluation in bf16: two tcgen05 GEMMs, bias + ReLU and bias + residual + LayerNorm in their epilogues
            from ..ops import linear_act, linear_residual_layernorm
            h = linear_act(x, self.w_1.weight, self.w_1.bias, relu=True)
            out = linear_residual_layernorm(h, self.w_2.weight, self.w_2.bias, x, self.layer_norm.weight,
                                            self.layer_norm.bias, self.layer_norm.eps)
            return out if out_scale is None else out * out_scale
        return dropout_residual_layer_norm(self.layer_norm, self.dropout, self.w_2(self.w_1(x, relu=True)), x, out_scale)


def fused_linear_ok(module, x, *linears):
    """The fused tcgen05 linear layers are forward-only and bf16: used when no gradient is being recorded, the
    dropout between the projection and the LayerNorm is off, and every shape fits (N % 128 == 0, K % 64 == 0,
    LayerNorm width 512)."""
    if torch.is_grad_enabled() or not x.is_cuda or x.dtype != torch.bfloat16:
        return False
    if module.training and getattr(module, "dropout", None) is not None and module.dropout.p > 0:
        return False
    if module.layer_norm.normalized_shape != (512,):
        return False
    return all(l.weight.dtype == torch.bfloat16 and l.out_features % 128 == 0 and l.in_features % 64 == 0 for l in linears)
