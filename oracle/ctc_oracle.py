"""CTC loss + gradient oracle (numpy).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates what the reference computes at `cal_ctc_ce_loss`
(/root/reference/src/transformer/loss.py:34-48) and `cal_loss`
(/root/reference/src/ctcModel/loss.py:4-13):

    target_lengths = (targets != 0).sum(1)
    lp   = log_softmax(logits, -1)                      # [B,T,V]
    loss = F.ctc_loss(lp^T, targets, len_logits, target_lengths, blank=V-1)
           (defaults: reduction='mean', zero_infinity=False)

`F.ctc_loss` is torch (third-party to the reference; ATen LossCTC.cpp, torch
2.11.0 here).  Its published algorithm, restated: with the blank-extended
sequence ext = (blank, l1, blank, ..., lS, blank) of length 2S+1,
    log_alpha_0(0) = lp_0(blank), log_alpha_0(1) = lp_0(l1)
    log_alpha_t(s) = lp_t(ext_s) + LSE(log_alpha_{t-1}(s), log_alpha_{t-1}(s-1),
                                       [ext_s != blank and ext_s != ext_{s-2}] log_alpha_{t-1}(s-2))
    nll = -LSE(log_alpha_{T-1}(2S), log_alpha_{T-1}(2S-1))
beta is the mirror image from t = T-1 (both include lp_t(ext_s)), and
    d nll / d logits[t,c] = softmax_t(c) - exp(LSE_{s: ext_s=c}(log_alpha_t(s)+log_beta_t(s)) + nll - lp_t(c))
for t < input_length, exactly 0 beyond it.  reduction='mean' divides each
utterance by max(target_length, 1) and then by B.  An infeasible alignment
gives nll = +inf and NaN in the label/blank columns of d nll / d log_probs
(-inf + inf) with zero_infinity=False; because the reference differentiates
through log_softmax (row sum of the incoming gradient), every valid frame row of
such an utterance is NaN at the logits, rows beyond input_length stay 0.
"""
import numpy as np


def _lse(a, axis=None):
    a = np.asarray(a)
    m = np.max(a, axis=axis, keepdims=True)
    m_safe = np.where(np.isfinite(m), m, 0.0).astype(a.dtype)
    with np.errstate(divide="ignore"):
        r = np.log(np.sum(np.exp(a - m_safe), axis=axis, keepdims=True)) + m_safe
    return np.squeeze(r, axis=axis) if axis is not None else r.reshape(())


def _shift(a, k, fill):
    """out[i] = a[i-k] (k>0 shifts right, k<0 left), `fill` where out of range."""
    out = np.full_like(a, fill)
    n = a.shape[0]
    if k > 0 and k < n:
        out[k:] = a[:n - k]
    elif k < 0 and -k < n:
        out[:n + k] = a[-k:]
    return out


def log_softmax(x):
    m = x.max(-1, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(-1, keepdims=True))


def ctc_loss_and_grad(logits, targets, in_len, blank=None, dtype=np.float64, need_grad=True):
    """Returns (loss, nll [B], grad [B,T,V] of the MEAN loss w.r.t. logits or None).

    logits [B,T,V] float, targets [B,S] int (0 = padding), in_len [B] int."""
    logits = np.asarray(logits).astype(dtype)
    targets = np.asarray(targets).astype(np.int64)
    in_len = np.asarray(in_len).astype(np.int64)
    B, T, V = logits.shape
    if blank is None:
        blank = V - 1
    tgt_len = (targets != 0).sum(1)
    lp_all = log_softmax(logits)
    nll = np.zeros((B,), dtype=dtype)
    grad = np.zeros((B, T, V), dtype=dtype) if need_grad else None
    neg_inf = dtype(-np.inf)
    for b in range(B):
        Tb, Sb = int(in_len[b]), int(tgt_len[b])
        lab = targets[b, :Sb]
        ext = np.full((2 * Sb + 1,), blank, dtype=np.int64)
        ext[1::2] = lab
        Sx = ext.shape[0]
        skip = np.zeros((Sx,), dtype=bool)
        skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
        lp = lp_all[b]
        if Tb == 0:
            # ATen: no frames -> nll = 0 if the target is empty, else inf
            nll[b] = dtype(0) if Sb == 0 else dtype(np.inf)
            continue
        la = np.full((Tb, Sx), neg_inf, dtype=dtype)
        la[0, 0] = lp[0, blank]
        if Sx > 1:
            la[0, 1] = lp[0, ext[1]]
        for t in range(1, Tb):
            prev = la[t - 1]
            p1 = _shift(prev, 1, neg_inf)
            p2 = np.where(skip, _shift(prev, 2, neg_inf), neg_inf)
            la[t] = _lse(np.stack([prev, p1, p2]), axis=0) + lp[t, ext]
        tail = la[Tb - 1, Sx - 2:] if Sx > 1 else la[Tb - 1, Sx - 1:]
        nll[b] = -_lse(tail, axis=0)
        if not need_grad:
            continue
        lb = np.full((Tb, Sx), neg_inf, dtype=dtype)
        lb[Tb - 1, Sx - 1] = lp[Tb - 1, blank]
        if Sx > 1:
            lb[Tb - 1, Sx - 2] = lp[Tb - 1, ext[Sx - 2]]
        skip_fwd = np.zeros((Sx,), dtype=bool)     # may s jump to s+2 ?
        if Sx > 2:
            skip_fwd[:-2] = skip[2:]
        for t in range(Tb - 2, -1, -1):
            nxt = lb[t + 1]
            n1 = _shift(nxt, -1, neg_inf)
            n2 = np.where(skip_fwd, _shift(nxt, -2, neg_inf), neg_inf)
            lb[t] = _lse(np.stack([nxt, n1, n2]), axis=0) + lp[t, ext]
        lab_sum = la + lb                                   # [Tb, Sx]
        occ = np.full((Tb, V), neg_inf, dtype=dtype)
        for s in range(Sx):                                 # LSE-accumulate per class
            c = ext[s]
            occ[:, c] = np.logaddexp(occ[:, c], lab_sum[:, s])
        touched = np.zeros((V,), dtype=bool)
        touched[ext] = True
        g = np.exp(lp[:Tb])
        with np.errstate(invalid="ignore", over="ignore"):
            g[:, touched] = g[:, touched] - np.exp(occ[:, touched] + nll[b] - lp[:Tb][:, touched])
        if not np.isfinite(nll[b]):
            # the reference differentiates through log_softmax: one NaN column
            # poisons sum_c(grad_lp) and with it the whole row
            g[:] = np.nan
        grad[b, :Tb] = g / dtype(max(Sb, 1) * B)
    with np.errstate(invalid="ignore"):
        loss = (nll / np.maximum(tgt_len, 1).astype(dtype)).mean()
    return loss, nll, grad
