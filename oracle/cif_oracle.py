"""CIF integrate-and-fire oracle (numpy).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates `CIF_Model.cif` (/root/reference/src/transformer/cif_model.py:57-106)
and the alpha scaling glue (cif_model.py:43-48) with IEEE float32 numpy ops in
the reference's exact operation order, so fire decisions and fp32 outputs are
bit-identical to the reference's torch-CPU path, plus the analytic backward
(SURVEY.md 8a row a2') in a caller-chosen dtype.

Per utterance, per frame t (state: scalar `integrate`, vector `frame`):
    dc        = 1 - integrate                      cif_model.py:69
    integrate = integrate + alpha_t                cif_model.py:71   (pre-reset value is the "fire" signal, :72)
    fire      = integrate > threshold              cif_model.py:74
    integrate = fire ? integrate - 1 : integrate   cif_model.py:75-77
    cur       = fire ? dc : alpha_t                cif_model.py:78-80
    rem       = alpha_t - cur                      cif_model.py:81
    frame     = frame + cur * h_t                  cif_model.py:83   (pre-reset frame is what is emitted, :84)
    frame     = fire ? rem * h_t : frame           cif_model.py:85-87
Output row k of utterance b is the pre-reset frame at the k-th fired t, rows
beyond the number of fires are zero, L = max_b int(round(sum_t alpha_bt))
(round-half-even; cif_model.py:95-101).
"""
import numpy as np

_F32 = np.float32


def cif_scale_alphas(alpha, targets, rand):
    """cif_model.py:43-48.  `rand` replaces torch.rand(B) (values in [0,1)).

    Returns (_num, num, scaled_alpha): _num = sum_t alpha (before scaling; this
    is what the quantity loss sees), num = #(targets > 0), scaled = alpha *
    ((num + rand - 0.5) / _num).  NOTE numpy's float32 row sum is pairwise and
    may differ from torch's in the last bit; callers that need the reference's
    exact `_num` pass it back in through `cif_scale_with_num`.
    """
    alpha = np.asarray(alpha, dtype=_F32)
    _num = alpha.sum(-1, dtype=_F32)
    num = (np.asarray(targets) > 0).astype(_F32).sum(-1, dtype=_F32)
    return _num, num, cif_scale_with_num(alpha, _num, num, rand)


def cif_scale_with_num(alpha, _num, num, rand):
    num_noise = (num + np.asarray(rand, dtype=_F32)) - _F32(0.5)
    return (np.asarray(alpha, dtype=_F32) * (num_noise / _num)[:, None]).astype(_F32)


def build_lfr_features(inputs, m, n):
    """utils/data.py:191-218 for one utterance [T,D]: frame i = frames i*n .. i*n+m-1 side by side,
    the last frame repeated past the end.  Returns [ceil(T/n), m*D]."""
    inputs = np.asarray(inputs)
    T = inputs.shape[0]
    idx = np.minimum(np.arange((T + n - 1) // n)[:, None] * n + np.arange(m)[None, :], T - 1)
    return inputs[idx].reshape(idx.shape[0], -1)


def assigner_tail_forward(x, w, b, lens, num_noise=None, dtype=np.float64):
    """attentionAssigner.py:36-40 + cif_model.py:43-48.
    x [B,T,D], w [D] (or [1,D]), b scalar, lens [B]; num_noise [B] or None (no scaling).
    Returns (alpha_raw [B,T], _num [B], alpha [B,T])."""
    x = np.asarray(x).astype(dtype)
    w = np.asarray(w).astype(dtype).reshape(-1)
    z = x @ w + dtype(np.asarray(b).reshape(-1)[0])                      # linear(x).squeeze(-1)      :37
    s = 1.0 / (1.0 + np.exp(-z))                                         # sigmoid                    :38
    mask = (np.arange(x.shape[1])[None, :] < np.asarray(lens)[:, None])  # sequence_mask              :39
    a_raw = s * mask                                                     #                            :41
    num = a_raw.sum(-1)                                                  # _num                       cif_model.py:44
    if num_noise is None:
        return a_raw, num, a_raw
    r = np.asarray(num_noise).astype(dtype) / num
    return a_raw, num, a_raw * r[:, None]                                #                            :48


def assigner_tail_backward(x, w, b, lens, num_noise, g_alpha, g_num=None, dtype=np.float64):
    """Gradients of (sum(alpha * g_alpha) + sum(_num * g_num)) w.r.t. x, w, b (analytic)."""
    x = np.asarray(x).astype(dtype)
    w1 = np.asarray(w).astype(dtype).reshape(-1)
    a_raw, num, _ = assigner_tail_forward(x, w1, b, lens, num_noise, dtype)
    g_alpha = np.asarray(g_alpha).astype(dtype)
    mask = (np.arange(x.shape[1])[None, :] < np.asarray(lens)[:, None])
    c = np.zeros(x.shape[0], dtype) if g_num is None else np.asarray(g_num).astype(dtype).copy()
    r = np.ones(x.shape[0], dtype)
    if num_noise is not None:
        r = np.asarray(num_noise).astype(dtype) / num
        c = c - (r / num) * (g_alpha * a_raw).sum(-1)
    d_a = g_alpha * r[:, None] + c[:, None]
    dz = d_a * mask * a_raw * (1.0 - a_raw)
    g_x = dz[:, :, None] * w1[None, None, :]
    g_w = np.einsum("bt,btd->d", dz, x)
    return g_x, g_w, dz.sum()


def cif_schedule(alphas, threshold):
    """Scalar recurrence only.  Returns dict of [B,T] arrays: fire (bool),
    cur, rem (float32), seg (int32, number of fires strictly before t) and
    n_fired [B]."""
    alphas = np.asarray(alphas, dtype=_F32)
    B, T = alphas.shape
    thr = _F32(threshold)
    one = _F32(1.0)
    integrate = np.zeros((B,), dtype=_F32)
    fire = np.zeros((B, T), dtype=bool)
    cur = np.zeros((B, T), dtype=_F32)
    rem = np.zeros((B, T), dtype=_F32)
    seg = np.zeros((B, T), dtype=np.int32)
    count = np.zeros((B,), dtype=np.int32)
    for t in range(T):
        a = alphas[:, t]
        dc = one - integrate
        s = integrate + a
        f = s > thr
        integrate = np.where(f, s - one, s).astype(_F32)
        c = np.where(f, dc, a).astype(_F32)
        fire[:, t] = f
        cur[:, t] = c
        rem[:, t] = a - c
        seg[:, t] = count
        count = count + f.astype(np.int32)
    return {"fire": fire, "cur": cur, "rem": rem, "seg": seg, "n_fired": count}


def cif_label_len(alphas):
    """L of cif_model.py:95-96 (numpy float32 sum; round half to even)."""
    s = np.asarray(alphas, dtype=_F32).sum(-1, dtype=_F32)
    return int(np.rint(s).astype(np.int32).max()) if s.size else 0


def cif_forward(hidden, alphas, threshold, L=None):
    """Returns (out [B,L,H] float32, fire_t [B,max(L,1)] int32 (-1 padded), n_fired [B]).

    Raises ValueError when an utterance fires more than L times (the reference
    fails in torch.zeros with a negative size, cif_model.py:100)."""
    hidden = np.asarray(hidden, dtype=_F32)
    alphas = np.asarray(alphas, dtype=_F32)
    B, T, H = hidden.shape
    if L is None:
        L = cif_label_len(alphas)
    sch = cif_schedule(alphas, threshold)
    if int(sch["n_fired"].max(initial=0)) > L:
        raise ValueError("an utterance fires %d times but L=%d" % (int(sch["n_fired"].max()), L))
    out = np.zeros((B, L, H), dtype=_F32)
    fire_t = np.full((B, max(L, 1)), -1, dtype=np.int32)
    frame = np.zeros((B, H), dtype=_F32)
    k = np.zeros((B,), dtype=np.int64)
    for t in range(T):
        h = hidden[:, t, :]
        f = sch["fire"][:, t]
        pre = frame + sch["cur"][:, t, None] * h          # two roundings: mul then add (no FMA)
        for b in np.nonzero(f)[0]:
            out[b, k[b]] = pre[b]
            fire_t[b, k[b]] = t
            k[b] += 1
        frame = np.where(f[:, None], sch["rem"][:, t, None] * h, pre).astype(_F32)
    return out, fire_t, sch["n_fired"].copy()


def cif_backward(hidden, alphas, threshold, g_out, dtype=np.float32):
    """Analytic backward of cif_forward (SURVEY.md 8a row a2').

    With seg(t) = number of fires strictly before t, the carried frame gradient
    is constant inside a segment, so for every frame
        gpre_t = g_out[seg(t)]        (0 when seg(t) >= n_fired)
        G_t    = g_out[seg(t) + 1]    (only used on fired frames; 0 past the end)
        gh_t   = cur_t * gpre_t + (fire_t ? rem_t * G_t : 0)
        d1_t   = <gpre_t, h_t>,  d2_t = fire_t ? <G_t, h_t> : 0,  gcur_t = d1_t - d2_t
        galpha_t = -S_t + d2_t + (fire_t ? 0 : gcur_t),   S_t = sum_{s>t, fire_s} gcur_s
    (the comparison itself carries no gradient; dc_t = 1 - integrate_{t-1} is
    where S comes from).  Returns (g_hidden [B,T,H], g_alpha [B,T]) in `dtype`.
    """
    sch = cif_schedule(alphas, threshold)
    hidden = np.asarray(hidden).astype(dtype)
    g_out = np.asarray(g_out).astype(dtype)
    B, T, H = hidden.shape
    L = g_out.shape[1]
    cur = sch["cur"].astype(dtype)
    rem = sch["rem"].astype(dtype)
    g_hidden = np.zeros((B, T, H), dtype=dtype)
    g_alpha = np.zeros((B, T), dtype=dtype)
    zero = np.zeros((H,), dtype=dtype)
    for b in range(B):
        nf = min(int(sch["n_fired"][b]), L)
        S = dtype(0)
        for t in range(T - 1, -1, -1):
            sg = int(sch["seg"][b, t])
            fired = bool(sch["fire"][b, t])
            gpre = g_out[b, sg] if sg < nf else zero
            h = hidden[b, t]
            gh = cur[b, t] * gpre
            d1 = dtype(np.dot(gpre, h))
            d2 = dtype(0)
            if fired:
                G = g_out[b, sg + 1] if sg + 1 < nf else zero
                gh = gh + rem[b, t] * G
                d2 = dtype(np.dot(G, h))
            gcur = d1 - d2
            g_hidden[b, t] = gh
            g_alpha[b, t] = -S + d2 + (dtype(0) if fired else gcur)
            if fired:
                S = S + gcur
    return g_hidden, g_alpha


def spec_aug_apply(features, lens, f0, fw, t0, tw):
    """Masking loops of spec_aug, /root/reference/src/utils/utils.py:168-194, with the drawn bands / spans given
    ([R,B] integer arrays): both means from the batch as it was on entry (:171-173), all frequency bands first
    (:176-183), then all time spans (:185-192), each as the reference's slice assignment.  Returns a copy."""
    x = np.array(features, dtype=np.float32, copy=True)
    B, T, V = x.shape
    freq_means = x.mean(axis=-1, dtype=np.float32)
    time_means = x.sum(axis=1, dtype=np.float32) / np.asarray(lens, dtype=np.float32)[:, None]
    for r in range(np.asarray(f0).shape[0]):
        for b in range(B):
            a, w = int(f0[r][b]), int(fw[r][b])
            x[b, :, a:a + w] = freq_means[b][:, None]
    for r in range(np.asarray(t0).shape[0]):
        for b in range(B):
            a, w = int(t0[r][b]), int(tw[r][b])
            x[b, a:a + w, :] = time_means[b][None, :]
    return x
