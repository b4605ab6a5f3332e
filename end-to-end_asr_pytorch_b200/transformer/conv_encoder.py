"""Caller-side convolutional front ends (plain PyTorch / cuDNN; SURVEY.md 2 marks them
out of the kernel scope).  Module and parameter names follow
/root/reference/src/transformer/conv_encoder.py (`conv.subsample/conv0`, `affine`,
`conv.<name>/conv1d_<i>`) so checkpoints interchange."""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from .module import Linear
import torch.nn.functional as F


class Conv1d(nn.Module):
    """n_layers x (Conv1d(w_context) + ReLU) over time, 'same' length by right padding."""

    def __init__(self, d_input, d_hidden, n_layers, w_context, pad='same', name=''):
        super().__init__()
        assert n_layers >= 1 and pad == 'same'
        self.n_layers, self.w_context = n_layers, w_context
        stack = OrderedDict()
        for i in range(n_layers):
            stack["{}/conv1d_{}".format(name, i)] = nn.Conv1d(d_input if i == 0 else d_hidden, d_hidden, w_context, 1)
            stack["{}/relu_{}".format(name, i)] = nn.ReLU()
        self.conv = nn.Sequential(stack)

    def forward(self, feats, feat_lengths):
        n_frames = feats.size(1)
        x = F.pad(feats, (0, 0, 0, self.n_layers * self.w_context))
        x = self.conv(x.transpose(1, 2)).transpose(1, 2)
        return x[:, :n_frames, :], feat_lengths


class Conv2dSubsample(nn.Module):
    """n_layers x (Conv2d(1|32 -> 32, 3x3, stride (2,1)) + ReLU), then an affine map to d_model.
    Time is halved (ceil) per layer."""

    def __init__(self, d_input, d_model, n_layers=2, pad='same'):
        super().__init__()
        assert n_layers >= 1 and pad == 'same'
        self.n_layers, self.d_input = n_layers, d_input
        stack = OrderedDict()
        for i in range(n_layers):
            stack["subsample/conv{}".format(i)] = nn.Conv2d(1 if i == 0 else 32, 32, 3, (2, 1))
            stack["subsample/relu{}".format(i)] = nn.ReLU()
        self.conv = nn.Sequential(stack)
        self.d_conv_out = int(math.ceil(d_input / 2))
        self.affine = Linear(32 * self.d_conv_out, d_model)

    def forward(self, feats, feat_lengths):
        n_frames = feats.size(1)
        x = F.pad(feats, (0, 10, 0, 20)).unsqueeze(1)                  # [B, 1, T+20, D+10]
        if x.is_cuda:
            # NHWC activations on the GPU: cuDNN's tensor-core convolutions are NHWC kernels, and with NCHW tensors every
            # call is wrapped in layout-conversion kernels (measured in the config-5 training step: 1.2 ms of nchwToNhwc /
            # nhwcToNchw per step around 1.3 ms of convolutions).  The parameters keep torch's default layout (so do their
            # gradients, which the flat all-reduce buckets and the fused optimizer rely on); the 9 K-element NHWC copy of
            # a weight is made per call.  With one input channel NCHW and NHWC are the same bytes, but torch reads the
            # layout off the strides and a size-1 dimension is ambiguous: spell the NHWC strides out.
            _, _, hh, ww = x.shape
            x = x.as_strided(x.shape, (hh * ww, 1, ww, 1))
            for layer in self.conv:
                if isinstance(layer, nn.Conv2d):
                    x = F.conv2d(x, layer.weight.contiguous(memory_format=torch.channels_last), layer.bias, layer.stride)
                else:
                    x = layer(x)
        else:
            x = self.conv(x)
        x = x[:, :, :, :self.d_conv_out]
        B, C, T, D = x.size()
        x = x.permute(0, 2, 1, 3).contiguous().view(B, T, C * D)
        out_len = feat_lengths
        for _ in range(self.n_layers):
            out_len = torch.ceil(out_len / 2.0).int()
            n_frames = int(math.ceil(n_frames / 2.0))
        return self.affine(x[:, :n_frames, :]), out_len
