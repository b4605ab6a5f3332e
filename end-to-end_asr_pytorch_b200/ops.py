"""torch.autograd.Function wrappers over the C ABI (include/asr_sm100.h).

PyTorch is plumbing here: it owns device memory and streams; every hot-path
computation below is one call into libasr_sm100.so.  No CPU fallback: tensors
must live on a CUDA (sm_100) device.
"""
import ctypes

import torch

from . import _lib
from ._lib import ptr, stream_ptr, check


def _require_cuda(name, t, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: libasr_sm100 has no CPU path" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))


# ---------------------------------------------------------------------------------
# CIF  (reference: src/transformer/cif_model.py:57-106)
# ---------------------------------------------------------------------------------
class _CifFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, alphas, threshold, L, target_num):
        B, T, H = hidden.shape
        dev = hidden.device
        out = torch.empty((B, L, H), dtype=torch.float32, device=dev)
        fire_t = torch.empty((B, L), dtype=torch.int32, device=dev)
        n_fired = torch.empty((B,), dtype=torch.int32, device=dev)
        cur = torch.empty((B, T), dtype=torch.float32, device=dev)
        rem = torch.empty((B, T), dtype=torch.float32, device=dev)
        sched = torch.empty((B, T), dtype=torch.int32, device=dev)
        alpha_sum = torch.empty((B,), dtype=torch.float32, device=dev)
        qua_term = torch.empty((B,), dtype=torch.float32, device=dev) if target_num is not None else None
        with torch.cuda.device(dev):
            check(_lib.lib().asr_cif_fwd_f32(
                ptr(hidden), ptr(alphas), ctypes.c_float(threshold), B, T, H, L,
                ptr(out) if L > 0 else None, ptr(fire_t) if L > 0 else None, ptr(n_fired),
                ptr(cur), ptr(rem), ptr(sched), ptr(alpha_sum), ptr(target_num), ptr(qua_term),
                stream_ptr()), "asr_cif_fwd_f32")
        ctx.save_for_backward(hidden, n_fired, cur, rem, sched)
        ctx.dims = (B, T, H, L)
        if qua_term is None:
            qua_term = alpha_sum.new_empty((0,))
        ctx.mark_non_differentiable(fire_t, n_fired, alpha_sum, qua_term)
        return out, fire_t, n_fired, alpha_sum, qua_term

    @staticmethod
    def backward(ctx, g_out, *unused):
        hidden, n_fired, cur, rem, sched = ctx.saved_tensors
        B, T, H, L = ctx.dims
        dev = hidden.device
        g_out = g_out.contiguous()
        if g_out.dtype != torch.float32:
            g_out = g_out.float()
        g_hidden = torch.empty_like(hidden)
        g_alphas = torch.empty((B, T), dtype=torch.float32, device=dev)
        ws_bytes = _lib.lib().asr_cif_bwd_workspace_bytes(B, T)
        ws = torch.empty((max(ws_bytes, 4) // 4,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().asr_cif_bwd_f32(
                ptr(hidden), ptr(g_out) if L > 0 else None, ptr(n_fired), ptr(cur), ptr(rem), ptr(sched),
                B, T, H, L, ptr(g_hidden), ptr(g_alphas), ptr(ws), ws_bytes, stream_ptr()), "asr_cif_bwd_f32")
        return g_hidden, g_alphas, None, None, None


class _CifAlphaFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, lens, num_noise):
        B, T, D = x.shape
        dev = x.device
        alpha = torch.empty((B, T), dtype=torch.float32, device=dev)
        a_raw = torch.empty((B, T), dtype=torch.float32, device=dev)
        num_raw = torch.empty((B,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().asr_cif_alpha_fwd_f32(ptr(x), ptr(weight), ptr(bias), ptr(lens), ptr(num_noise), B, T, D,
                                                   ptr(alpha), ptr(a_raw), ptr(num_raw), stream_ptr()),
                  "asr_cif_alpha_fwd_f32")
        ctx.save_for_backward(x, weight, lens, num_noise, a_raw, num_raw)
        ctx.wshape = weight.shape
        return alpha, num_raw

    @staticmethod
    def backward(ctx, g_alpha, g_num):
        x, weight, lens, num_noise, a_raw, num_raw = ctx.saved_tensors
        B, T, D = x.shape
        dev = x.device
        g_alpha = g_alpha.contiguous().float()
        g_num = g_num.contiguous().float() if g_num is not None else None
        g_x = torch.empty_like(x)
        g_w = torch.empty((D,), dtype=torch.float32, device=dev)
        g_b = torch.empty((1,), dtype=torch.float32, device=dev)
        ws_bytes = _lib.lib().asr_cif_alpha_bwd_workspace_bytes(B, T, D)
        ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().asr_cif_alpha_bwd_f32(ptr(x), ptr(weight), ptr(lens), ptr(num_noise), ptr(a_raw), ptr(num_raw),
                                                   ptr(g_alpha), ptr(g_num), B, T, D, ptr(g_x), ptr(g_w), ptr(g_b),
                                                   ptr(ws), ws_bytes, stream_ptr()), "asr_cif_alpha_bwd_f32")
        return g_x, g_w.view(ctx.wshape), g_b, None, None


def cif_alpha(x, weight, bias, lens, num_noise=None):
    """The CIF weight producer: tail of the attention assigner + the scaling of CIF_Model.forward.

    x [B,T,D] f32 (assigner activations), weight [1,D] or [D], bias [1] (the assigner's `linear`),
    lens [B] valid frames, num_noise [B] (= #labels + U[0,1) - 0.5; None: no scaling)
    -> (alpha [B,T] scaled weights, _num [B] = sum_t of the unscaled weights).
    attentionAssigner.py:36-40 and cif_model.py:43-48 in one pass over x."""
    _require_cuda("x", x, torch.float32)
    if x.dim() != 3:
        raise ValueError("cif_alpha: x must be [B, T, D]")
    x = x.contiguous()
    weight = weight.contiguous().float()
    bias = bias.contiguous().float().view(1)
    lens = lens.to(device=x.device, dtype=torch.int32).contiguous()
    if num_noise is not None:
        num_noise = num_noise.to(device=x.device, dtype=torch.float32).contiguous()
    if weight.numel() != x.shape[-1]:
        raise ValueError("cif_alpha: weight must have D elements")
    return _CifAlphaFunction.apply(x, weight, bias, lens, num_noise)


def build_lfr_features(features, lens, m, n):
    """Low-frame-rate stacking of a padded batch on the device (utils/data.py:191-218 applied to every
    utterance): features [B,T,D] f32, lens [B] -> (out [B, ceil(T/n), m*D], out_lens [B] = ceil(lens/n)).
    Frame i of utterance b is frames i*n .. i*n+m-1 side by side, the last frame repeated past the end;
    rows beyond out_lens[b] are zero.  A pure copy, no gradient."""
    _require_cuda("features", features, torch.float32)
    if features.dim() != 3:
        raise ValueError("build_lfr_features: features must be [B, T, D]")
    features = features.detach().contiguous()
    B, T, D = features.shape
    lens = lens.to(device=features.device, dtype=torch.int32).contiguous()
    To = (T + n - 1) // n
    out = torch.empty((B, To, m * D), dtype=torch.float32, device=features.device)
    out_lens = torch.empty((B,), dtype=torch.int32, device=features.device)
    with torch.cuda.device(features.device):
        check(_lib.lib().asr_lfr_f32(ptr(features), ptr(lens), B, T, D, int(m), int(n), ptr(out), ptr(out_lens),
                                     stream_ptr()), "asr_lfr_f32")
    return out, out_lens


def spec_aug_apply(features, lens, f0, fw, t0, tw):
    """The masking part of spec_aug (utils/utils.py:168-194) on the device, IN PLACE: features [B,T,V] f32,
    lens [B], and the drawn frequency bands (f0, fw) / time spans (t0, tw) as [R,B] integer tensors.  Cells in
    a time span get the per-bin mean over time (sum / lens), other cells in a frequency band the per-frame
    mean over bins, both of the batch as it was on entry.  Returns features.  No gradient (the reference
    masks the input features, which carry none)."""
    _require_cuda("features", features, torch.float32)
    if features.dim() != 3 or not features.is_contiguous():
        raise ValueError("spec_aug_apply: features must be a contiguous [B, T, V] tensor (masked in place)")
    B, T, V = features.shape
    dev = features.device
    lens = lens.to(device=dev, dtype=torch.int32).contiguous()
    R = f0.shape[0] if f0.dim() == 2 else 0
    for m in (f0, fw, t0, tw):
        if m.dim() != 2 or tuple(m.shape) != (R, B):
            raise ValueError("spec_aug_apply: the mask arrays must all be [R, B]")
    # one [4,R,B] i32 block (one stack + one cast instead of four casts)
    masks = torch.stack([m.to(dev) for m in (f0, fw, t0, tw)]).to(torch.int32).contiguous()
    wsb = _lib.lib().asr_spec_aug_workspace_bytes(B, T, V)
    ws = torch.empty((wsb + 3) // 4, dtype=torch.float32, device=dev)
    base, step = masks.data_ptr(), R * B * 4
    with torch.cuda.device(dev):
        check(_lib.lib().asr_spec_aug_f32(ptr(features), ptr(lens), base, base + step, base + 2 * step, base + 3 * step,
                                          int(R), B, T, V, ptr(ws), wsb, stream_ptr()), "asr_spec_aug_f32")
    return features


def linear_act(x, weight, bias=None, relu=False):
    """y = act(x W^T + b) on the tensor cores with the bias / ReLU fused into the epilogue (module.py:50,
    attention.py:40-45): x [..., K] bf16, weight [N, K] bf16, bias [N] (any float dtype) -> [..., N] bf16.
    Forward only (evaluation path); N % 128 == 0 and K % 64 == 0."""
    _require_cuda("x", x, torch.bfloat16)
    _require_cuda("weight", weight, torch.bfloat16)
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.dim() != 2 or weight.shape[1] != K:
        raise ValueError("linear_act: weight must be [N, K] with K = x.shape[-1]")
    x2 = x.detach().reshape(-1, K).contiguous()
    w = weight.detach().contiguous()
    b = None if bias is None else bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
    y = torch.empty((x2.shape[0], N), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().asr_linear_act_bf16(ptr(x2), ptr(w), ptr(b), x2.shape[0], N, K, 1 if relu else 0, ptr(y), stream_ptr()),
              "asr_linear_act_bf16")
    return y.reshape(*x.shape[:-1], N)


def linear_f32(x, weight, bias=None):
    """y = x W^T + b in fp32 on the tensor cores with fp32-level accuracy (three TF32 products per tile):
    x [..., K] f32, weight [N, K] f32 -> [..., N] f32.  Forward only; K % 4 == 0."""
    _require_cuda("x", x, torch.float32)
    _require_cuda("weight", weight, torch.float32)
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.dim() != 2 or weight.shape[1] != K:
        raise ValueError("linear_f32: weight must be [N, K] with K = x.shape[-1]")
    x2 = x.detach().reshape(-1, K).contiguous()
    w = weight.detach().contiguous()
    b = None if bias is None else bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
    y = torch.empty((x2.shape[0], N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().asr_linear_f32(ptr(x2), ptr(w), ptr(b), x2.shape[0], N, K, ptr(y), stream_ptr()), "asr_linear_f32")
    return y.reshape(*x.shape[:-1], N)


def _gemm_ws(M, N, K, device):
    nbytes = _lib.lib().asr_gemm_workspace_bytes(M, N, K)
    return torch.empty((nbytes // 4 + 1,), dtype=torch.float32, device=device), nbytes


def gemm_f32(a, b, a_mn_major=False, b_mn_major=False, bias=None, out=None, n_valid=None, split_k=True, row_len=None,
             group_rows=0, skip_dead_output=False):
    """C = A B (+ bias) in fp32 on the tensor cores at fp32-level accuracy (three TF32 products per K step), operands as stored:
    a: [M,K] (a_mn_major False) or [K,M] (True); b: [N,K] (False) or [K,N] (True); 2-D, last dim contiguous, row strides
    multiples of 4.  `out` may be a [M, ld >= N] buffer (its first N columns are written, plus zeros up to the next
    multiple of 4).  The building block of the linear layers' backward products and of ops.ctc_fc_loss.
    row_len [groups] i32 + group_rows: ragged rows (padded utterances of `group_rows` frames each, row_len[g] valid) -
    tiles / K steps that lie entirely in padding are skipped, see asr_gemm_f32_ragged."""
    _require_cuda("a", a, torch.float32)
    _require_cuda("b", b, torch.float32)
    if a.dim() != 2 or b.dim() != 2 or a.stride(1) != 1 or b.stride(1) != 1:
        raise ValueError("gemm_f32: 2-D operands with a contiguous last dimension expected")
    M, K = (a.shape[1], a.shape[0]) if a_mn_major else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn_major else (b.shape[0], b.shape[1])
    if n_valid is not None:
        if b_mn_major:
            N = n_valid
        else:
            raise ValueError("gemm_f32: n_valid applies to an MN-major b")
    if K != Kb:
        raise ValueError("gemm_f32: contraction lengths differ (%d vs %d)" % (K, Kb))
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    elif out.dim() != 2 or out.shape[0] != M or out.shape[1] < N or out.stride(1) != 1:
        raise ValueError("gemm_f32: out must be [M, >= N] with a contiguous last dimension")
    ws, wsb = _gemm_ws(M, N, K, a.device) if split_k else (None, 0)
    if bias is not None:
        bias = bias.detach().to(device=a.device, dtype=torch.float32).contiguous()
    with torch.cuda.device(a.device):
        check(_lib.lib().asr_gemm_f32_ragged(ptr(a), int(a_mn_major), a.stride(0), ptr(b), int(b_mn_major), b.stride(0), ptr(bias),
                                             M, N, K, ptr(out), out.stride(0), ptr(row_len), int(group_rows),
                                             int(skip_dead_output), ptr(ws), wsb, stream_ptr()), "asr_gemm_f32")
    return out


def gemm_bf16(a, b, a_mn_major=False, b_mn_major=False, bias=None, relu=False, out_dtype=torch.bfloat16, split_k=True):
    """C = A B (+ bias, ReLU) with bf16 operands and fp32 accumulation on the tensor cores; same operand conventions as
    gemm_f32 (row strides multiples of 8).  out_dtype bf16 or fp32 (weight gradients for fp32 master weights)."""
    _require_cuda("a", a, torch.bfloat16)
    _require_cuda("b", b, torch.bfloat16)
    if a.dim() != 2 or b.dim() != 2 or a.stride(1) != 1 or b.stride(1) != 1:
        raise ValueError("gemm_bf16: 2-D operands with a contiguous last dimension expected")
    M, K = (a.shape[1], a.shape[0]) if a_mn_major else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn_major else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError("gemm_bf16: contraction lengths differ (%d vs %d)" % (K, Kb))
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    ws, wsb = _gemm_ws(M, N, K, a.device) if split_k else (None, 0)
    if bias is not None:
        bias = bias.detach().to(device=a.device, dtype=torch.float32).contiguous()
    with torch.cuda.device(a.device):
        check(_lib.lib().asr_gemm_bf16(ptr(a), int(a_mn_major), a.stride(0), ptr(b), int(b_mn_major), b.stride(0), ptr(bias),
                                       int(relu), M, N, K, ptr(out), out.stride(0), int(out_dtype == torch.float32), ptr(ws), wsb,
                                       stream_ptr()), "asr_gemm_bf16")
    return out


def colsum(x):
    """Column sums of a 2-D fp32 / bf16 matrix (last dim contiguous) in fp32 - the bias gradient of a linear layer - in one
    pass with a fixed summation order (deterministic)."""
    _require_cuda("x", x)
    if x.dim() != 2 or x.stride(1) != 1 or x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("colsum: 2-D fp32 / bf16 matrix with a contiguous last dimension expected")
    M, N = x.shape
    out = torch.empty((N,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().asr_colsum(ptr(x), int(x.dtype == torch.bfloat16), M, N, x.stride(0), ptr(out), stream_ptr()), "asr_colsum")
    return out


class _LinearF32Function(torch.autograd.Function):
    """y = x W^T + b with all three products (y, dx = gy W, dW = gy^T x) on the fp32 tensor-core GEMM, every operand read
    as torch stores it (K-major or MN-major descriptors): no transposed copies.  Reference: what autograd does for an
    nn.Linear of the model shell (module.py:46-53, attention.py:40-45,59-60) - there as cuBLAS SIMT sgemm."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        if x2.stride(1) != 1 or x2.stride(0) % 4 != 0 or x2.data_ptr() % 16 != 0:
            x2 = x2.contiguous()
        w = weight if weight.is_contiguous() else weight.contiguous()
        ctx.save_for_backward(x2, w)
        ctx.x_shape = x.shape
        ctx.has_bias = bias is not None
        return gemm_f32(x2, w, bias=bias).reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        N, K = w.shape
        gy2 = gy.reshape(-1, N)
        if N % 4 != 0:      # odd widths (a 4233-class vocabulary): rows padded to 16 bytes, the operand is the [:, :N] view
            buf = gy2.new_empty((gy2.shape[0], (N + 3) // 4 * 4))
            buf[:, :N] = gy2
            gy2 = buf[:, :N]
        elif gy2.stride(1) != 1 or gy2.stride(0) % 4 != 0 or gy2.data_ptr() % 16 != 0:
            gy2 = gy2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm_f32(gy2, w, b_mn_major=True).reshape(ctx.x_shape)      # contraction over N
        if ctx.needs_input_grad[1]:
            gw = gemm_f32(gy2, x2, a_mn_major=True, b_mn_major=True)                         # contraction over the rows
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = colsum(gy2)
        return gx, gw, gb


def linear_f32_ok(x, weight):
    """Shapes the fp32 tensor-core linear layer takes: 16-byte rows for every operand of its three products; an odd
    output width (the 4233-wide vocabulary projections) costs one padding copy of the incoming gradient, which is
    worth it from ~64 columns on (the assigner's 1-wide head stays with torch)."""
    N, K = weight.shape
    return K % 4 == 0 and (N % 4 == 0 or N >= 64) and x.numel() > 0


def linear_f32_autograd(x, weight, bias=None):
    """Differentiable fp32 linear layer on the tensor cores (three TF32 products per tile, see gemm_f32)."""
    return _LinearF32Function.apply(x, weight, bias)


class _LinearBf16Function(torch.autograd.Function):
    """bf16 linear layer for mixed-precision training (fp32 master weights, bf16 activations): y = act(x W^T + b) with the
    bias / ReLU in the GEMM epilogue; backward dx = (gy o relu') W in bf16 and dW = gy^T x accumulated and written in fp32
    (the master weights' dtype), both on operands as stored.  The ReLU mask comes from the saved output (y > 0)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        if x2.dtype != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)
        if x2.stride(1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
            x2 = x2.contiguous()
        w16 = weight.detach().to(torch.bfloat16).contiguous()
        y = gemm_bf16(x2, w16, bias=bias, relu=relu)
        ctx.save_for_backward(x2, w16, y if relu else None)
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        ctx.has_bias, ctx.relu = bias is not None, relu
        ctx.w_dtype = weight.dtype
        return y.reshape(*x.shape[:-1], w16.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w16, y = ctx.saved_tensors
        N, K = w16.shape
        gy2 = gy.reshape(-1, N)
        if ctx.relu:
            if gy2.dtype != torch.bfloat16:
                gy2 = gy2.to(torch.bfloat16)
            if gy2.is_contiguous() and y.is_contiguous() and gy2.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0:
                masked = torch.empty_like(gy2)
                with torch.cuda.device(gy2.device):
                    check(_lib.lib().asr_relu_bwd_bf16(ptr(gy2), ptr(y), ptr(masked), gy2.numel(), stream_ptr()), "asr_relu_bwd_bf16")
                gy2 = masked
            else:
                gy2 = gy2 * (y > 0).to(gy2.dtype)
        if N % 8 != 0:      # odd widths (the 4233-class vocabulary): rows padded to 16 bytes, the operand is the [:, :N] view
            buf = torch.empty((gy2.shape[0], (N + 7) // 8 * 8), dtype=torch.bfloat16, device=gy2.device)
            buf[:, :N] = gy2          # converts on the way
            gy2 = buf[:, :N]
        else:
            if gy2.dtype != torch.bfloat16:
                gy2 = gy2.to(torch.bfloat16)
            if gy2.stride(1) != 1 or gy2.stride(0) % 8 != 0 or gy2.data_ptr() % 16 != 0:
                gy2 = gy2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm_bf16(gy2, w16, b_mn_major=True).reshape(ctx.x_shape).to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            gw = gemm_bf16(gy2, x2, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32).to(ctx.w_dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = colsum(gy2)
        return gx, gw, gb, None


def linear_bf16_ok(x, weight):
    """Shapes the bf16 tensor-core linear layer takes: 16-byte rows for every operand of its three products; an odd output
    width (the 4233-wide vocabulary projection) costs one padding copy of the incoming gradient, as in linear_f32_ok."""
    N, K = weight.shape
    return K % 8 == 0 and (N % 8 == 0 or N >= 64) and x.numel() > 0


def linear_bf16_autograd(x, weight, bias=None, relu=False):
    """Differentiable bf16 linear layer (+ fused bias / ReLU) on the tensor cores, fp32 weight gradients."""
    return _LinearBf16Function.apply(x, weight, bias, bool(relu))


def linear_residual_layernorm(x, weight, bias, residual, ln_weight, ln_bias, eps=1e-5, one_kernel=False):
    """y = LayerNorm(x W^T + b + residual) (module.py:50-52, attention.py:59-60 with dropout off): x [..., K] bf16,
    weight [512, K] bf16, residual [..., 512] bf16, LayerNorm weight / bias [512] -> [..., 512] bf16.  Forward only.
    Default route: the persistent bf16 GEMM (bias in its epilogue, bf16 out - what torch's bf16 F.linear rounds to as well)
    followed by the bf16 LayerNorm kernel: 0.22 ms at M = 102400, K = 2048.  one_kernel=True: round 1's single kernel,
    which keeps the whole 512-wide fp32 row in tensor memory (the pre-norm activations are never rounded): 0.31 ms."""
    _require_cuda("x", x, torch.bfloat16)
    _require_cuda("weight", weight, torch.bfloat16)
    _require_cuda("residual", residual, torch.bfloat16)
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.dim() != 2 or weight.shape[1] != K or residual.shape[-1] != N or residual.shape[:-1] != x.shape[:-1]:
        raise ValueError("linear_residual_layernorm: weight [N, K], residual [..., N] with x [..., K]")
    x2 = x.detach().reshape(-1, K).contiguous()
    r2 = residual.detach().reshape(-1, N).contiguous()
    w = weight.detach().contiguous()
    f32 = lambda t: None if t is None else t.detach().to(device=x.device, dtype=torch.float32).contiguous()
    b, g, be = f32(bias), f32(ln_weight), f32(ln_bias)
    y = torch.empty((x2.shape[0], N), dtype=torch.bfloat16, device=x.device)
    if not one_kernel and N in LN_WIDTHS and K % 8 == 0:
        pre = gemm_bf16(x2, w, bias=b)
        with torch.cuda.device(x.device):
            check(_lib.lib().asr_ln_eval_bf16(ptr(pre), ptr(r2), ptr(g), ptr(be), x2.shape[0], N, ctypes.c_float(eps), ptr(y),
                                              stream_ptr()), "asr_ln_eval_bf16")
        return y.reshape(*x.shape[:-1], N)
    with torch.cuda.device(x.device):
        check(_lib.lib().asr_linear_residual_layernorm_bf16(ptr(x2), ptr(w), ptr(b), ptr(r2), ptr(g), ptr(be), float(eps),
                                                            x2.shape[0], N, K, ptr(y), stream_ptr()),
              "asr_linear_residual_layernorm_bf16")
    return y.reshape(*x.shape[:-1], N)


def cif_label_len(alphas):
    """L of cif_model.py:95-96: max_b int(round(sum_t alphas)) - one host sync, like the reference."""
    return int(torch.round(alphas.sum(-1)).int().max().item())


def cif(hidden, alphas, threshold, L=None, target_num=None, check_overflow=True, return_aux=False):
    """Integrate-and-fire.  hidden [B,T,H] f32, alphas [B,T] f32 -> [B,L,H] f32.

    L defaults to the reference's max_b round(sum_t alphas).  With return_aux the
    result is (out, aux) where aux holds fire_t [B,L] (frame index of each fire),
    n_fired [B], alpha_sum [B] and, when target_num [B] is given, qua_term [B] =
    (alpha_sum - target_num)^2.
    """
    _require_cuda("hidden", hidden, torch.float32)
    _require_cuda("alphas", alphas, torch.float32)
    if hidden.dim() != 3 or alphas.dim() != 2 or hidden.shape[:2] != alphas.shape:
        raise ValueError("cif: hidden [B,T,H] and alphas [B,T] expected, got %s and %s"
                         % (tuple(hidden.shape), tuple(alphas.shape)))
    hidden_c = hidden.contiguous()
    alphas_c = alphas.contiguous()
    if L is None:
        L = cif_label_len(alphas_c)
    if target_num is not None:
        _require_cuda("target_num", target_num, torch.float32)
        target_num = target_num.contiguous()
    out, fire_t, n_fired, alpha_sum, qua_term = _CifFunction.apply(hidden_c, alphas_c, float(threshold), int(L), target_num)
    if check_overflow:
        worst = int(n_fired.max().item()) if n_fired.numel() else 0
        if worst > L:
            # the reference dies in torch.zeros([max_label_len - l.size(0), H]) (cif_model.py:100)
            raise RuntimeError("cif: an utterance fired %d times but the output holds L=%d rows "
                               "(L = max round(sum alphas), cif_model.py:95-100)" % (worst, L))
    if return_aux:
        aux = {"fire_t": fire_t, "n_fired": n_fired, "alpha_sum": alpha_sum,
               "qua_term": qua_term if target_num is not None else None}
        return out, aux
    return out


# ---------------------------------------------------------------------------------
# CTC loss  (reference: src/transformer/loss.py:39-43, src/ctcModel/loss.py:7-11)
# ---------------------------------------------------------------------------------
class _CtcLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, in_len, tgt_len, blank):
        B, T, V = logits.shape
        S = targets.shape[1]
        dev = logits.device
        need_grad = ctx.needs_input_grad[0]
        nll = torch.empty((B,), dtype=torch.float32, device=dev)
        g = torch.empty_like(logits) if need_grad else None
        ws_bytes = _lib.lib().asr_ctc_workspace_bytes(B, T, V, S)
        ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().asr_ctc_fwd_bwd_f32(
                ptr(logits), ptr(targets) if S > 0 else None, ptr(in_len), ptr(tgt_len),
                B, T, V, S, int(blank), ptr(nll), ptr(g), ptr(ws), ws_bytes, stream_ptr()),
                "asr_ctc_fwd_bwd_f32")
        # reduction='mean' of F.ctc_loss: mean_b(nll_b / clamp(target_len_b, 1))
        loss = (nll / tgt_len.clamp(min=1).to(nll.dtype)).mean()
        ctx.g = g
        ctx.mark_non_differentiable(nll)
        return loss, nll

    @staticmethod
    def backward(ctx, g_loss, g_nll_unused):
        g = ctx.g
        if g is None:
            raise RuntimeError("ctc_loss: backward called twice (the fused gradient buffer is handed out once)")
        ctx.g = None
        g_loss = g_loss.reshape(1).to(dtype=torch.float32, device=g.device).contiguous()
        with torch.cuda.device(g.device):
            # scales in place on the device only when the incoming gradient is not exactly 1
            check(_lib.lib().asr_scale_inplace_f32(ptr(g), g.numel(), ptr(g_loss), stream_ptr()),
                  "asr_scale_inplace_f32")
        return g, None, None, None, None


def ctc_loss(logits, len_logits, targets, blank=None, return_nll=False):
    """Mean CTC loss of raw logits [B,T,V] (log-softmax fused), targets [B,S]
    0-padded int64, blank = V-1 by default - the reference's call convention."""
    _require_cuda("logits", logits)
    if logits.dim() != 3:
        raise ValueError("ctc_loss: logits [B,T,V] expected")
    if logits.dtype != torch.float32:
        logits = logits.float()
    B, T, V = logits.shape
    if blank is None:
        blank = V - 1
    logits_c = logits.contiguous()
    targets_c = targets.to(device=logits.device, dtype=torch.int64).contiguous()
    tgt_len = targets_c.ne(0).sum(1).to(torch.int32)
    in_len = len_logits.to(device=logits.device, dtype=torch.int32).contiguous()
    loss, nll = _CtcLossFunction.apply(logits_c, targets_c, in_len, tgt_len, blank)
    if return_nll:
        return loss, nll
    return loss


# ---------------------------------------------------------------------------------
# Vocabulary projection fused with the CTC loss  (reference: ctc_fc at src/transformer/cif_model.py:38,
# transformer.py:148, ctcModel/decoder.py:32-36 followed by the loss call of loss.py:39-43)
# ---------------------------------------------------------------------------------
class _CtcFcLossFunction(torch.autograd.Function):
    """loss = ctc(h W^T): the [B,T,V] logits live only inside this call, in rows padded to a multiple of 4 floats.
    One fp32 tensor-core GEMM writes them, the CTC kernels replace them IN PLACE by d loss / d logits (every row is
    read once and written once), and two more GEMMs read that gradient as stored - K-major for d h = g W, MN-major
    for d W = g^T h - so no [B,T,V] tensor is ever copied, transposed or kept for backward."""

    @staticmethod
    def forward(ctx, hidden, weight, targets, in_len, tgt_len, blank):
        B, T, K = hidden.shape
        V = weight.shape[0]
        dev = hidden.device
        ld = (V + 3) // 4 * 4
        h2 = hidden.reshape(B * T, K)
        if not h2.is_contiguous():
            h2 = h2.contiguous()
        w = weight if weight.is_contiguous() else weight.contiguous()
        buf = torch.empty((B * T, ld), dtype=torch.float32, device=dev)
        # frames beyond an utterance's length: their logits are never read and their gradient rows are exactly zero, so all
        # three products skip the row tiles / K steps that lie entirely in the padding
        rag = dict(row_len=in_len, group_rows=T)
        gemm_f32(h2, w, out=buf, split_k=False, skip_dead_output=True, **rag)     # logits, columns [0, V)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        nll = torch.empty((B,), dtype=torch.float32, device=dev)
        S = targets.shape[1]
        ws_bytes = _lib.lib().asr_ctc_workspace_bytes(B, T, V, S)
        ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().asr_ctc_fwd_bwd_ld_f32(ptr(buf), ptr(targets) if S > 0 else None, ptr(in_len), ptr(tgt_len),
                                                    B, T, V, ld, S, int(blank), ptr(nll), ptr(buf) if need_grad else None,
                                                    ptr(ws), ws_bytes, stream_ptr()), "asr_ctc_fwd_bwd_ld_f32")
        loss = (nll / tgt_len.clamp(min=1).to(nll.dtype)).mean()
        gh = gw = None
        if need_grad:
            g = buf[:, :V]                                                        # [B*T, V], row stride ld
            if ctx.needs_input_grad[0]:
                gh = gemm_f32(g, w, b_mn_major=True, split_k=False, **rag).reshape(B, T, K)
            if ctx.needs_input_grad[1]:
                gw = gemm_f32(g, h2, a_mn_major=True, b_mn_major=True, **rag)
        ctx.grads = (gh, gw)
        ctx.mark_non_differentiable(nll)
        return loss, nll

    @staticmethod
    def backward(ctx, g_loss, g_nll_unused):
        gh, gw = ctx.grads
        ctx.grads = (None, None)
        ref = gh if gh is not None else gw
        if ref is None:
            return None, None, None, None, None, None
        scale = g_loss.reshape(1).to(dtype=torch.float32, device=ref.device).contiguous()
        with torch.cuda.device(ref.device):
            for t in (gh, gw):
                if t is not None:      # scales in place on the device only when the incoming gradient is not exactly 1
                    check(_lib.lib().asr_scale_inplace_f32(ptr(t), t.numel(), ptr(scale), stream_ptr()), "asr_scale_inplace_f32")
        return gh, gw, None, None, None, None


def ctc_fc_loss(hidden, weight, len_logits, targets, blank=None, return_nll=False):
    """Mean CTC loss of the projected logits hidden [B,T,K] x weight [V,K]^T (no bias, like ctc_fc) without materialising
    them for the caller: see _CtcFcLossFunction.  Same conventions as ctc_loss (targets [B,S] 0-padded int64, blank = V-1)."""
    _require_cuda("hidden", hidden, torch.float32)
    _require_cuda("weight", weight, torch.float32)
    if hidden.dim() != 3 or weight.dim() != 2 or weight.shape[1] != hidden.shape[2]:
        raise ValueError("ctc_fc_loss: hidden [B,T,K] and weight [V,K] expected")
    if hidden.shape[2] % 4 != 0:
        raise ValueError("ctc_fc_loss: K must be a multiple of 4")
    V = weight.shape[0]
    if blank is None:
        blank = V - 1
    targets_c = targets.to(device=hidden.device, dtype=torch.int64).contiguous()
    tgt_len = targets_c.ne(0).sum(1).to(torch.int32)
    in_len = len_logits.to(device=hidden.device, dtype=torch.int32).contiguous()
    loss, nll = _CtcFcLossFunction.apply(hidden, weight, targets_c, in_len, tgt_len, blank)
    return (loss, nll) if return_nll else loss


class ProjectedLogits:
    """Stand-in for the `ctc_logits` tensor a model returns when the projection is fused with the loss: it carries the
    projection's input and weight to `cal_ctc_ce_loss` / `cal_ctc_qua_ce_loss` / `cal_loss`, which call ctc_fc_loss.
    `materialize()` gives the actual [B,T,V] tensor to any other consumer."""

    def __init__(self, hidden, weight):
        self.hidden, self.weight = hidden, weight

    def size(self, dim=None):
        shape = tuple(self.hidden.shape[:-1]) + (self.weight.shape[0],)
        return shape if dim is None else shape[dim]

    @property
    def shape(self):
        return self.size()

    def materialize(self):
        return torch.nn.functional.linear(self.hidden, self.weight)


# ---------------------------------------------------------------------------------
# Multi-head attention core  (reference: src/transformer/attention.py:74-86)
# ---------------------------------------------------------------------------------
class DropoutSeed:
    """A dropout seed that lives on the device, for training steps captured in a CUDA graph.

    Host-side arguments are frozen when a graph is captured, so a seed drawn on the host would replay the same
    dropout masks for ever.  With an active DropoutSeed every attention call of the step uses the mask of
    `*seed_dev + seed_add`: `seed_add` is a per-call constant handed out by `take()` while the step is traced,
    and `advance()` - a device-side add, captured with the rest of the step - moves `*seed_dev` on for the next replay.

        ds = ops.DropoutSeed(device)
        with ops.device_dropout_seed(ds):
            with torch.cuda.graph(g):
                loss = step(); loss.backward(); ds.advance()
    """
    _STRIDE = 0x9E3779B97F4A7C15        # odd: distinct calls never meet for any realistic number of steps

    def __init__(self, device, seed=None):
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self.tensor = torch.tensor([seed], dtype=torch.int64, device=device)
        self.calls = 0

    def take(self):
        self.calls += 1
        return self.tensor, (self.calls * self._STRIDE) & 0xFFFFFFFFFFFFFFFF

    def advance(self):
        self.tensor.add_(1)
        self.calls = 0


_active_seed = None


class device_dropout_seed:
    """Context manager: attention dropout inside draws its seeds from `ds` (see DropoutSeed)."""

    def __init__(self, ds):
        self.ds = ds

    def __enter__(self):
        global _active_seed
        self.prev, _active_seed = _active_seed, self.ds
        return self.ds

    def __exit__(self, *exc):
        global _active_seed
        _active_seed = self.prev
        return False


class _MhaCoreFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, kv_len, dense_mask, causal, scale, p_drop, seed, seed_dev):
        B, Lq, Hh, D = q.shape
        Lk = k.shape[1]
        out = torch.empty((B, Lq, Hh, D), dtype=torch.bfloat16, device=q.device)
        lse = torch.empty((B, Hh, Lq), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            if seed_dev is None:
                check(_lib.lib().asr_mha_fwd_dropout_bf16(ptr(q), ptr(k), ptr(v), ptr(kv_len), ptr(dense_mask), int(causal),
                                                          B, Hh, Lq, Lk, D, ctypes.c_float(scale), ctypes.c_float(p_drop),
                                                          ctypes.c_uint64(seed), ptr(out), ptr(lse), stream_ptr()),
                      "asr_mha_fwd_dropout_bf16")
            else:
                check(_lib.lib().asr_mha_fwd_dropout_dev_bf16(ptr(q), ptr(k), ptr(v), ptr(kv_len), ptr(dense_mask), int(causal),
                                                              B, Hh, Lq, Lk, D, ctypes.c_float(scale), ctypes.c_float(p_drop),
                                                              ptr(seed_dev), ctypes.c_uint64(seed), ptr(out), ptr(lse),
                                                              stream_ptr()), "asr_mha_fwd_dropout_dev_bf16")
        ctx.save_for_backward(q, k, v, out, lse, kv_len, dense_mask, seed_dev)
        ctx.meta = (causal, scale, p_drop, seed)
        return out

    @staticmethod
    def backward(ctx, g_out):
        q, k, v, out, lse, kv_len, dense_mask, seed_dev = ctx.saved_tensors
        causal, scale, p_drop, seed = ctx.meta
        B, Lq, Hh, D = q.shape
        Lk = k.shape[1]
        g_out = g_out.contiguous()
        if g_out.dtype != torch.bfloat16:
            g_out = g_out.to(torch.bfloat16)
        g_q, g_k, g_v = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ws_bytes = _lib.lib().asr_mha_bwd_workspace_bytes(B, Hh, Lq, Lk, D)
        ws = torch.empty((ws_bytes // 4 + 1,), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            if seed_dev is None:
                check(_lib.lib().asr_mha_bwd_dropout_bf16(ptr(q), ptr(k), ptr(v), ptr(out), ptr(g_out), ptr(lse), ptr(kv_len),
                                                          ptr(dense_mask), int(causal), B, Hh, Lq, Lk, D, ctypes.c_float(scale),
                                                          ctypes.c_float(p_drop), ctypes.c_uint64(seed),
                                                          ptr(g_q), ptr(g_k), ptr(g_v), ptr(ws), ws_bytes, stream_ptr()),
                      "asr_mha_bwd_dropout_bf16")
            else:
                check(_lib.lib().asr_mha_bwd_dropout_dev_bf16(ptr(q), ptr(k), ptr(v), ptr(out), ptr(g_out), ptr(lse), ptr(kv_len),
                                                              ptr(dense_mask), int(causal), B, Hh, Lq, Lk, D,
                                                              ctypes.c_float(scale), ctypes.c_float(p_drop), ptr(seed_dev),
                                                              ctypes.c_uint64(seed), ptr(g_q), ptr(g_k), ptr(g_v), ptr(ws),
                                                              ws_bytes, stream_ptr()), "asr_mha_bwd_dropout_dev_bf16")
        return g_q, g_k, g_v, None, None, None, None, None, None, None


def _check_mha_args(what, q, k, v, kv_len, mask):
    """Shapes the kernels take as raw pointers: k / v [B,Lk,Hh,D] matching q [B,Lq,Hh,D] with D = 64, kv_len [B],
    mask broadcastable to [B,Lq,Lk] (expanded here - the reference's masked_fill accepts such masks too)."""
    B, Lq, Hh, D = q.shape
    if D != 64:
        raise ValueError("%s: head dimension %d, the sm_100a attention kernels are built for 64" % (what, D))
    for name, t in (("k", k), ("v", v)):
        if t is not None and (t.shape[0] != B or t.shape[2] != Hh or t.shape[3] != D or t.shape[1] != k.shape[1]):
            raise ValueError("%s: %s %s does not match q %s" % (what, name, tuple(t.shape), tuple(q.shape)))
    Lk = k.shape[1]
    if kv_len is not None:
        if kv_len.numel() != B:
            raise ValueError("%s: kv_len must have one entry per batch row (%d), got %s" % (what, B, tuple(kv_len.shape)))
        kv_len = kv_len.reshape(B).to(device=q.device, dtype=torch.int32).contiguous()
    if mask is not None:
        if mask.dim() != 3:
            raise ValueError("%s: mask must be [B, Lq, Lk] (or broadcastable to it), got %s" % (what, tuple(mask.shape)))
        try:
            mask = mask.to(device=q.device).expand(B, Lq, Lk)
        except RuntimeError:
            raise ValueError("%s: mask %s is not broadcastable to [%d, %d, %d]" % (what, tuple(mask.shape), B, Lq, Lk)) from None
        mask = mask.ne(0).to(torch.uint8).contiguous()
    return kv_len, mask


def mha_core(q, k, v, kv_len=None, mask=None, causal=False, scale=None, dropout_p=0.0, seed=None):
    """softmax(mask(q k^T * scale)) v on the tensor cores.

    q [B,Lq,Hh,64], k,v [B,Lk,Hh,64] (any float dtype; computed in bf16, fp32 accumulate)
    -> [B,Lq,Hh,64] bf16.  Masking: kv_len [B] (keys >= kv_len[b] masked), causal, and/or a
    dense mask [B,Lq,Lk] (True / non-zero = masked), combined with OR.
    dropout_p > 0 applies the reference's dropout to the probabilities (attention.py:83) inside
    the kernel; `seed` (default: drawn from torch's CPU generator, so torch.manual_seed makes it
    reproducible; inside `device_dropout_seed(...)`: read from the device, see DropoutSeed) selects the
    mask, which backward regenerates."""
    _require_cuda("q", q)
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise ValueError("mha_core: q, k, v must be [B, L, heads, 64]")
    if not 0.0 <= dropout_p < 1.0:
        raise ValueError("mha_core: dropout_p must be in [0, 1)")
    if scale is None:
        scale = 1.0 / (q.shape[-1] ** 0.5)
    qb, kb, vb = (t.to(torch.bfloat16).contiguous() for t in (q, k, v))
    kv_len, mask = _check_mha_args("mha_core", q, k, v, kv_len, mask)
    seed_dev = None
    if dropout_p > 0.0 and seed is None:
        if _active_seed is not None:
            seed_dev, seed = _active_seed.take()
        else:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return _MhaCoreFunction.apply(qb, kb, vb, kv_len, mask, bool(causal), float(scale), float(dropout_p),
                                  int(seed or 0), seed_dev)


def mha_dropout_keep(B, Hh, Lq, Lk, dropout_p, seed, device="cuda"):
    """(keep mask [B,Hh,Lq,Lk] bool, keep probability) of mha_core's dropout for (dropout_p, seed)."""
    keep = torch.empty((B, Hh, Lq, Lk), dtype=torch.uint8, device=device)
    with torch.cuda.device(keep.device):
        check(_lib.lib().asr_mha_dropout_keep_u8(B, Hh, Lq, Lk, ctypes.c_float(dropout_p), ctypes.c_uint64(seed),
                                                 ptr(keep), stream_ptr()), "asr_mha_dropout_keep_u8")
    return keep.bool(), float(_lib.lib().asr_mha_dropout_keep_prob(ctypes.c_float(dropout_p)))


def mha_probs(q, k, kv_len=None, mask=None, causal=False, scale=None):
    """Attention probabilities [(heads*B), Lq, Lk] f32 in the reference's head-major row order
    (attention.py:47,62).  Not on the training path; for callers that want `attn`."""
    _require_cuda("q", q)
    B, Lq, Hh, D = q.shape
    Lk = k.shape[1]
    if scale is None:
        scale = 1.0 / (D ** 0.5)
    qb, kb = q.detach().to(torch.bfloat16).contiguous(), k.detach().to(torch.bfloat16).contiguous()
    kv_len, mask = _check_mha_args("mha_probs", q, k, None, kv_len, mask)
    attn = torch.empty((Hh * B, Lq, Lk), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        check(_lib.lib().asr_mha_probs_f32(ptr(qb), ptr(kb), ptr(kv_len), ptr(mask), int(causal), B, Hh, Lq, Lk, D,
                                           ctypes.c_float(scale), ptr(attn), stream_ptr()), "asr_mha_probs_f32")
    return attn


# ---------------------------------------------------------------------------------------
# dropout + residual + LayerNorm of the training step (csrc/ln.cu)
# ---------------------------------------------------------------------------------------
LN_WIDTHS = (256, 512, 1024)


class _ResidualLayerNormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, residual, weight, bias, eps, p_drop, seed, seed_dev, row_scale, want_bf16):
        D = y.shape[-1]
        y2 = y.reshape(-1, D)
        r2 = residual.reshape(-1, D) if residual is not None else None
        M = y2.shape[0]
        dev = y.device
        out = torch.empty((M, D), dtype=torch.float32, device=dev)
        mean = torch.empty((M,), dtype=torch.float32, device=dev)
        rstd = torch.empty((M,), dtype=torch.float32, device=dev)
        z_is_y = r2 is None and p_drop == 0.0 and y2.dtype == torch.float32
        z = None if z_is_y else torch.empty((M, D), dtype=torch.float32, device=dev)
        out16 = torch.empty((M, D), dtype=torch.bfloat16, device=dev) if want_bf16 else None
        w32, b32 = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        with torch.cuda.device(dev):
            check(_lib.lib().asr_ln_fwd(ptr(y2), int(y2.dtype == torch.bfloat16), ptr(r2), ptr(w32), ptr(b32), ptr(row_scale), M, D,
                                        ctypes.c_float(eps), ctypes.c_float(p_drop), ctypes.c_uint64(seed), ptr(seed_dev),
                                        ptr(z), ptr(out), ptr(out16), ptr(mean), ptr(rstd), stream_ptr()), "asr_ln_fwd")
        ctx.save_for_backward(y2 if z_is_y else z, mean, rstd, w32, seed_dev, row_scale)
        ctx.meta = (y.shape, y.dtype, residual is not None, p_drop, seed)
        if want_bf16:
            return out.reshape(y.shape), out16.reshape(y.shape)
        return out.reshape(y.shape), None

    @staticmethod
    def backward(ctx, g_out, g_out16):
        z, mean, rstd, w32, seed_dev, row_scale = ctx.saved_tensors
        shape, y_dtype, has_res, p_drop, seed = ctx.meta
        M, D = z.shape
        dev = z.device
        g2 = g16 = None
        if g_out is not None:
            g2 = g_out.reshape(M, D)
            if g2.dtype != torch.float32 or not g2.is_contiguous():
                g2 = g2.float().contiguous()
        if g_out16 is not None:          # what came back through the bf16 copy (the next layer's dX): added inside the kernel
            g16 = g_out16.reshape(M, D)
            if g16.dtype != torch.bfloat16 or not g16.is_contiguous():
                g16 = g16.to(torch.bfloat16).contiguous()
        if g2 is None and g16 is None:
            g2 = torch.zeros((M, D), dtype=torch.float32, device=dev)
        need_y, need_r = ctx.needs_input_grad[0], has_res and ctx.needs_input_grad[1]
        same = p_drop == 0.0 and y_dtype == torch.float32          # dy is dz: one tensor serves both
        g_z = torch.empty((M, D), dtype=torch.float32, device=dev) if (need_r or (need_y and same)) else None
        g_y = torch.empty((M, D), dtype=y_dtype, device=dev) if (need_y and not same) else None
        if g_z is None and g_y is None:
            g_z = torch.empty((M, D), dtype=torch.float32, device=dev)
        g_wb = torch.empty((2, D), dtype=torch.float32, device=dev)
        ws_bytes = _lib.lib().asr_ln_bwd_workspace_bytes(M, D)
        ws = torch.empty((ws_bytes // 4 + 4,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().asr_ln_bwd(ptr(g2), ptr(g16), ptr(z), ptr(mean), ptr(rstd), ptr(w32), ptr(row_scale), M, D, ctypes.c_float(p_drop),
                                        ctypes.c_uint64(seed), ptr(seed_dev), ptr(g_z), ptr(g_y), int(y_dtype == torch.bfloat16),
                                        ptr(g_wb), ptr(ws), ws_bytes, stream_ptr()), "asr_ln_bwd")
        gy = (g_z if same else g_y).reshape(shape) if need_y else None
        gr = g_z.reshape(shape) if need_r else None
        return gy, gr, g_wb[0], g_wb[1], None, None, None, None, None, None


def residual_layer_norm_available(y, residual, weight):
    """True when residual_layer_norm takes these operands (CUDA, width 256 / 512 / 1024, y fp32 / bf16, residual fp32)."""
    return (y.is_cuda and y.shape[-1] in LN_WIDTHS and y.dtype in (torch.float32, torch.bfloat16) and y.is_contiguous()
            and weight.shape == (y.shape[-1],)
            and (residual is None or (residual.dtype == torch.float32 and residual.shape == y.shape and residual.is_contiguous())))


BF16_COPY_ATTR = "_asr_bf16"          # a tensor attribute: the bf16 copy residual_layer_norm wrote next to its fp32 output


def bf16_copy_of(x):
    """The bf16 copy that residual_layer_norm(bf16_copy=True) attached to its output (same values, same autograd node), or
    None.  The bf16 linear layers of the shell take it instead of converting x again."""
    c = getattr(x, BF16_COPY_ATTR, None)
    return c if (c is not None and c.shape == x.shape and c.device == x.device) else None


def residual_layer_norm(y, residual, weight, bias, eps=1e-5, dropout_p=0.0, training=True, seed=None, row_scale=None,
                        bf16_copy=False):
    """LayerNorm(dropout(y) + residual) [* row_scale] -> fp32, with autograd (module.py:50-52, attention.py:59-60, encoder.py:49 of the
    reference): one kernel forward, one (+ a column sum of per-CTA partials) backward.  y [..., D] fp32 or bf16, residual
    [..., D] fp32 or None, weight / bias [D] (gamma / beta).  The dropout mask comes from the library's Philox stream
    (`seed`: default drawn from torch's CPU generator; inside `device_dropout_seed(...)` read from the device) and is
    regenerated in the backward - see ln_dropout_keep.  row_scale: one constant factor per row ([...] or [..., 1], e.g. the
    non-pad mask the layers multiply every sub-layer output with - encoder.py:76-80, decoder.py:628-634).
    bf16_copy=True (bf16 autocast): the kernel also writes the output in bf16 and attaches it to the returned fp32 tensor
    (`bf16_copy_of`): the next bf16 GEMM reads it instead of running a conversion kernel, and the gradient that comes back
    through it is added to the fp32 one inside the backward kernel."""
    _require_cuda("y", y)
    if not residual_layer_norm_available(y, residual, weight):
        raise ValueError("residual_layer_norm: unsupported operands (width %d, dtypes %s / %s)" % (
            y.shape[-1], y.dtype, None if residual is None else residual.dtype))
    p = float(dropout_p) if training else 0.0
    if not 0.0 <= p < 1.0:
        raise ValueError("residual_layer_norm: dropout_p must be in [0, 1)")
    seed_dev = None
    if p > 0.0 and seed is None:
        if _active_seed is not None:
            seed_dev, seed = _active_seed.take()
        else:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    if row_scale is not None:
        if row_scale.numel() != y.numel() // y.shape[-1]:
            raise ValueError("residual_layer_norm: row_scale must hold one factor per row")
        row_scale = row_scale.detach().reshape(-1).to(device=y.device, dtype=torch.float32).contiguous()
    out, out16 = _ResidualLayerNormFunction.apply(y, residual, weight, bias, float(eps), p, int(seed or 0), seed_dev, row_scale,
                                                  bool(bf16_copy))
    if out16 is not None:
        setattr(out, BF16_COPY_ATTR, out16)
    return out


def ln_dropout_keep(M, D, dropout_p, seed, device="cuda"):
    """(keep mask [M, D] bool, keep probability) of residual_layer_norm's dropout for (dropout_p, seed)."""
    keep = torch.empty((M, D), dtype=torch.uint8, device=device)
    with torch.cuda.device(keep.device):
        check(_lib.lib().asr_ln_dropout_keep(ptr(keep), M, D, ctypes.c_float(dropout_p), ctypes.c_uint64(seed), stream_ptr()),
              "asr_ln_dropout_keep")
    return keep.bool(), float(_lib.lib().asr_ln_dropout_keep_prob(ctypes.c_float(dropout_p)))
