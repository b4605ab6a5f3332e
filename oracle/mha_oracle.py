"""Multi-head attention oracle (numpy).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates `MultiheadAttention.forward` and `ScaledDotProductAttention.forward`
(/root/reference/src/transformer/attention.py:33-62 and :74-86; the ctcModel
twin at src/ctcModel/attention.py is identical apart from the ctor order):

    q,k,v   = w_qs(q), w_ks(k), w_vs(v)                       :43-45
    split into heads, HEAD-MAJOR batch order (row = head * B + b)  :47-49
    attn    = softmax(masked_fill(q k^T / sqrt(d_k), mask, -inf), dim=2)   :76-82
    attn    = dropout(attn)   (training mode; here with an explicit keep mask)   :83
    out     = attn v                                          :84
    merge heads -> fc -> (dropout) -> LayerNorm(out + residual)    :56-60
A fully masked row is NaN in the reference (softmax over all -inf); so it is here.
"""
import numpy as np


def _softmax_lastdim(x):
    m = x.max(-1, keepdims=True)
    with np.errstate(invalid="ignore"):
        e = np.exp(x - m)
        return e / e.sum(-1, keepdims=True)


def mha_core_forward(q, k, v, mask=None, scale=None, dtype=np.float64, keep=None, keep_prob=1.0):
    """q [N,Lq,d], k [N,Lk,d], v [N,Lk,dv]; mask [N,Lq,Lk] bool (True = masked).
    keep [N,Lq,Lk] bool + keep_prob: the dropout of attention.py:83 with an explicit mask
    (nn.Dropout zeroes with probability p and scales the rest by 1/(1-p) = 1/keep_prob).
    Returns (out [N,Lq,dv], attn [N,Lq,Lk]); attn is the softmax BEFORE dropout."""
    q, k, v = (np.asarray(a).astype(dtype) for a in (q, k, v))
    if scale is None:
        scale = 1.0 / np.sqrt(q.shape[-1])
    s = np.einsum("nqd,nkd->nqk", q, k) * dtype(scale)
    if mask is not None:
        s = np.where(np.asarray(mask).astype(bool), dtype(-np.inf), s)
    attn = _softmax_lastdim(s)
    used = attn if keep is None else attn * np.asarray(keep).astype(dtype) / dtype(keep_prob)
    return np.einsum("nqk,nkd->nqd", used, v), attn


def mha_core_backward(q, k, v, g_out, mask=None, scale=None, dtype=np.float64, keep=None, keep_prob=1.0):
    """Gradients of mha_core_forward's `out` w.r.t. q, k, v."""
    q, k, v, g_out = (np.asarray(a).astype(dtype) for a in (q, k, v, g_out))
    if scale is None:
        scale = 1.0 / np.sqrt(q.shape[-1])
    _, p = mha_core_forward(q, k, v, mask, scale, dtype)
    m = 1.0 if keep is None else np.asarray(keep).astype(dtype) / dtype(keep_prob)
    g_v = np.einsum("nqk,nqd->nkd", p * m, g_out)
    g_p = np.einsum("nqd,nkd->nqk", g_out, v) * m        # through the dropout to the softmax output
    g_s = p * (g_p - (g_p * p).sum(-1, keepdims=True))
    g_q = np.einsum("nqk,nkd->nqd", g_s, k) * dtype(scale)
    g_k = np.einsum("nqk,nqd->nkd", g_s, q) * dtype(scale)
    return g_q, g_k, g_v


def split_heads(x, n_head):
    """[B,L,n_head*d] -> head-major [(n_head*B), L, d]   (attention.py:43-49)."""
    B, L, D = x.shape
    d = D // n_head
    return x.reshape(B, L, n_head, d).transpose(2, 0, 1, 3).reshape(n_head * B, L, d)


def merge_heads(x, n_head):
    """inverse of split_heads   (attention.py:56-57)."""
    N, L, d = x.shape
    B = N // n_head
    return x.reshape(n_head, B, L, d).transpose(1, 2, 0, 3).reshape(B, L, n_head * d)


def mha_module_forward(q, k, v, weights, n_head, mask=None, eps=1e-5, dtype=np.float64):
    """Whole MultiheadAttention.forward in eval mode.  `weights` uses the
    reference's state_dict keys (w_qs.weight, w_qs.bias, ..., fc.*, layer_norm.*).
    Returns (output [B,Lq,d_model], attn [n_head*B, Lq, Lk])."""
    w = {kk: np.asarray(vv).astype(dtype) for kk, vv in weights.items()}
    q, k, v = (np.asarray(a).astype(dtype) for a in (q, k, v))
    residual = q
    qh = split_heads(q @ w["w_qs.weight"].T + w["w_qs.bias"], n_head)
    kh = split_heads(k @ w["w_ks.weight"].T + w["w_ks.bias"], n_head)
    vh = split_heads(v @ w["w_vs.weight"].T + w["w_vs.bias"], n_head)
    d_k = qh.shape[-1]
    m = None if mask is None else np.tile(np.asarray(mask).astype(bool), (n_head, 1, 1))   # attention.py:52
    out, attn = mha_core_forward(qh, kh, vh, m, 1.0 / np.power(d_k, 0.5), dtype)
    out = merge_heads(out, n_head) @ w["fc.weight"].T + w["fc.bias"]
    x = out + residual
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    y = (x - mu) / np.sqrt(var + dtype(eps)) * w["layer_norm.weight"] + w["layer_norm.bias"]
    return y, attn
