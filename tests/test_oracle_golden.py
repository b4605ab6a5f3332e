"""Pin the CPU oracle against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import cif_oracle, ctc_oracle, mha_oracle
from conftest import load_golden

CIF = load_golden("cif")
CIF_CASES = sorted({k.split("_")[0] for k in CIF.files})
CTC = load_golden("ctc")
CTC_CASES = sorted({k.split("_")[0] for k in CTC.files})
MHA = load_golden("mha")
MHA_CASES = ["self_pad", "self_causal", "cross", "nomask"]


@pytest.mark.parametrize("case", CIF_CASES)
def test_cif_forward_bit_exact(case):
    hidden, alphas, thr = CIF[case + "_hidden"], CIF[case + "_alphas"], float(CIF[case + "_thr"])
    ref_out = CIF[case + "_out"]
    out, fire_t, n_fired = oracle.cif_forward(hidden, alphas, thr, L=ref_out.shape[1])
    # fire positions: integer work, must be identical
    np.testing.assert_array_equal(n_fired, CIF[case + "_n_fired"])
    np.testing.assert_array_equal(fire_t, CIF[case + "_fire_t"])
    # fp32 outputs in the reference's op order: identical bits
    assert out.shape == ref_out.shape
    np.testing.assert_array_equal(out.view(np.uint32), ref_out.view(np.uint32))
    # the numpy L agrees with torch.round(alphas.sum(-1)).int().max() on these cases
    assert cif_oracle.cif_label_len(alphas) == ref_out.shape[1]


@pytest.mark.parametrize("case", CIF_CASES)
def test_cif_backward_matches_reference_autograd(case):
    hidden, alphas, thr = CIF[case + "_hidden"], CIF[case + "_alphas"], float(CIF[case + "_thr"])
    g_out = CIF[case + "_g_out"]
    gh, ga = oracle.cif_backward(hidden, alphas, thr, g_out, dtype=np.float32)
    # g_hidden has no reductions: cur*gpre + rem*G, two terms -> bit exact
    np.testing.assert_array_equal(gh.view(np.uint32) & 0x7FFFFFFF,
                                  CIF[case + "_g_hidden"].view(np.uint32) & 0x7FFFFFFF)
    np.testing.assert_allclose(gh, CIF[case + "_g_hidden"], rtol=0, atol=0)
    # g_alpha has H-long dot products whose summation order differs from torch's
    gh64, ga64 = oracle.cif_backward(hidden, alphas, thr, g_out, dtype=np.float64)
    scale = np.abs(ga64).max() + 1e-30
    assert np.abs(ga - CIF[case + "_g_alpha"]).max() <= 2e-5 * scale
    assert np.abs(ga64 - CIF[case + "_g_alpha"]).max() <= 2e-5 * scale


def test_cif_overflowing_L_raises():
    hidden, alphas = CIF["a_hidden"], CIF["a_alphas"]
    with pytest.raises(ValueError):
        oracle.cif_forward(hidden, alphas, 0.95, L=3)


def test_cif_glue():
    g = load_golden("cif_glue")
    _num, num, scaled = oracle.cif_scale_alphas(g["alpha"], g["targets"], g["rand"])
    np.testing.assert_array_equal(num, g["num"])
    np.testing.assert_allclose(_num, g["_num"], rtol=1e-6)
    # with the reference's own _num the scaling is bit exact
    scaled2 = cif_oracle.cif_scale_with_num(g["alpha"], g["_num"], g["num"], g["rand"])
    np.testing.assert_array_equal(scaled2.view(np.uint32), g["scaled"].view(np.uint32))


def test_lfr_matches_reference():
    """utils/data.py:191-218 executed on small utterances (make_golden.py: make_lfr)."""
    g = load_golden("lfr")
    for name in sorted({k.split("_")[0] for k in g.files}):
        m, n = (int(v) for v in g[name + "_mn"])
        y = oracle.build_lfr_features(g[name + "_x"], m, n)
        assert y.shape == g[name + "_y"].shape
        np.testing.assert_array_equal(y.view(np.uint32), g[name + "_y"].view(np.uint32))


def _spec_aug_masks(g, name):
    """Bands / spans from the recorded torch.rand draws, with the reference's arithmetic (utils.py:178-181,186-189)."""
    import torch
    draws = torch.as_tensor(g[name + "_draws"])
    lens = torch.as_tensor(g[name + "_lens"])
    _, fwid, tnum, twid = (int(v) for v in g[name + "_cfg"])
    V = g[name + "_x"].shape[2]
    f0, fw, t0, tw = [], [], [], []
    for r in range(tnum):
        fs = (fwid * draws[2 * r]).long()
        fw.append(fs)
        f0.append(((V - fs).float() * draws[2 * r + 1]).long())
    for r in range(tnum):
        ts = (twid * draws[2 * tnum + 2 * r]).long()
        tw.append(ts)
        t0.append(((lens - ts).float() * draws[2 * tnum + 2 * r + 1]).long())
    return tuple(torch.stack(m).numpy() for m in (f0, fw, t0, tw))


def test_spec_aug_matches_reference():
    """utils/utils.py:168-194 executed on CPU with its torch.rand draws recorded
    (make_golden.py: make_spec_aug).  Unmasked cells bit-exact, the means to fp32 summation-order accuracy."""
    g = load_golden("spec_aug")
    for name in sorted({k.split("_")[0] for k in g.files}):
        f0, fw, t0, tw = _spec_aug_masks(g, name)
        y = oracle.spec_aug_apply(g[name + "_x"], g[name + "_lens"], f0, fw, t0, tw)
        ref = g[name + "_y"]
        same = ref == g[name + "_x"]
        assert 0.02 < 1.0 - same.mean() < 0.9                   # the fixture masks something, not everything
        np.testing.assert_array_equal(y[same].view(np.uint32), ref[same].view(np.uint32))
        np.testing.assert_allclose(y, ref, rtol=1e-5, atol=2e-6)


def test_spec_aug_draws_follow_the_reference():
    """The host mirror makes the reference's torch.rand calls in the reference's order: same generator state,
    same bands / spans (host logic only - nothing is launched)."""
    import torch
    from helpers import pkg
    uu = pkg("utils.utils")
    g = load_golden("spec_aug")
    for name in sorted({k.split("_")[0] for k in g.files}):
        x = g[name + "_x"]
        cfg = "-".join(str(int(v)) for v in g[name + "_cfg"])
        torch.manual_seed(int(g[name + "_seed"][0]))
        got = uu.spec_aug_draw(x.shape[0], x.shape[2], torch.as_tensor(g[name + "_lens"]), cfg, torch.device("cpu"))
        for a, b in zip(got, _spec_aug_masks(g, name)):
            np.testing.assert_array_equal(a.numpy(), b)


def test_assigner_tail_matches_reference():
    """attentionAssigner.py:36-40 + cif_model.py:43-48, values and autograd gradients
    produced by the reference's own ops (tests/golden/make_golden.py: make_assigner_tail)."""
    g = load_golden("assigner_tail")
    a_raw, num, alpha = oracle.assigner_tail_forward(g["x"], g["w"], g["b"], g["lens"], g["num_noise"])
    np.testing.assert_allclose(a_raw, g["alpha_raw"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(num, g["_num"], rtol=2e-6)
    np.testing.assert_allclose(alpha, g["scaled"], rtol=3e-6, atol=1e-7)
    assert not a_raw[1, 17:].any() and not a_raw[3, 1:].any()          # zero at padded frames
    g_x, g_w, g_b = oracle.assigner_tail_backward(g["x"], g["w"], g["b"], g["lens"], g["num_noise"], g["g_alpha"], g["g_num"])
    sx, sw = np.abs(g["g_x"]).max(), np.abs(g["g_w"]).max()
    assert np.abs(g_x - g["g_x"]).max() <= 1e-5 * sx
    assert np.abs(g_w - g["g_w"].reshape(-1)).max() <= 1e-5 * sw
    np.testing.assert_allclose(g_b, g["g_b"][0], rtol=1e-4)
    # without scaling (decoding path): alpha is the masked sigmoid itself
    a2, n2, al2 = oracle.assigner_tail_forward(g["x"], g["w"], g["b"], g["lens"], None)
    np.testing.assert_array_equal(a2, al2)


@pytest.mark.parametrize("case", CTC_CASES)
def test_ctc_matches_reference(case):
    logits, targets, in_len = CTC[case + "_logits"], CTC[case + "_targets"], CTC[case + "_in_len"]
    loss, nll, grad = oracle.ctc_loss_and_grad(logits, targets, in_len)
    ref_nll, ref_loss, ref_grad = CTC[case + "_nll"], CTC[case + "_loss"], CTC[case + "_grad"]
    fin = np.isfinite(ref_nll)
    np.testing.assert_array_equal(np.isfinite(nll), fin)
    np.testing.assert_allclose(nll[fin], ref_nll[fin], rtol=2e-6)
    if np.isfinite(ref_loss):
        np.testing.assert_allclose(loss, ref_loss, rtol=2e-6)
    else:
        assert np.isinf(loss) and loss > 0
    # gradients of feasible utterances; infeasible ones: same NaN pattern
    for b in range(logits.shape[0]):
        if fin[b]:
            np.testing.assert_allclose(grad[b], ref_grad[b], rtol=1e-4, atol=2e-7)
        else:
            np.testing.assert_array_equal(np.isnan(grad[b]), np.isnan(ref_grad[b]))
            ok = ~np.isnan(ref_grad[b])
            np.testing.assert_allclose(grad[b][ok], ref_grad[b][ok], rtol=1e-4, atol=2e-7)
    # beyond the input length the gradient is exactly zero
    for b in range(logits.shape[0]):
        assert not grad[b, int(in_len[b]):].any()


def test_ctc_fp32_mode_close_to_fp64():
    logits, targets, in_len = CTC["b_logits"], CTC["b_targets"], CTC["b_in_len"]
    l64, n64, g64 = oracle.ctc_loss_and_grad(logits, targets, in_len, dtype=np.float64)
    l32, n32, g32 = oracle.ctc_loss_and_grad(logits, targets, in_len, dtype=np.float32)
    np.testing.assert_allclose(n32, n64, rtol=1e-5)
    np.testing.assert_allclose(g32, g64, rtol=1e-3, atol=1e-6)


def test_quantity_loss():
    g = load_golden("qua")
    qua = np.mean((g["_number"].astype(np.float32) - g["number"].astype(np.float32)) ** 2, dtype=np.float32)
    np.testing.assert_allclose(qua, g["qua"], rtol=1e-6)
    loss, _, _ = oracle.ctc_loss_and_grad(g["logits"], g["targets"], g["in_len"], need_grad=False)
    np.testing.assert_allclose(loss, g["ctc"], rtol=2e-6)


def _weights():
    return {k[2:]: MHA[k] for k in MHA.files if k.startswith("w_")}


@pytest.mark.parametrize("case", MHA_CASES)
def test_mha_module_forward(case):
    q, kv = MHA[case + "_q"], MHA[case + "_kv"]
    mask = MHA[case + "_mask"]
    mask = None if mask.size == 0 else mask.astype(bool)
    y, attn = oracle.mha_module_forward(q, kv, kv, _weights(), n_head=2, mask=mask)
    np.testing.assert_allclose(y, MHA[case + "_y"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(attn, MHA[case + "_attn"], rtol=1e-4, atol=1e-6)
    # head-major row order: row = head * B + b
    assert attn.shape[0] == 2 * q.shape[0]


def test_mha_core_backward_against_finite_difference():
    rng = np.random.default_rng(0)
    N, Lq, Lk, d = 2, 3, 4, 5
    q, k, v = rng.normal(size=(N, Lq, d)), rng.normal(size=(N, Lk, d)), rng.normal(size=(N, Lk, d))
    g = rng.normal(size=(N, Lq, d))
    mask = np.zeros((N, Lq, Lk), bool)
    mask[0, :, 3] = True
    gq, gk, gv = oracle.mha_core_backward(q, k, v, g, mask)

    def f(qq, kk, vv):
        return (mha_oracle.mha_core_forward(qq, kk, vv, mask)[0] * g).sum()
    eps = 1e-6
    for arr, grad in ((q, gq), (k, gk), (v, gv)):
        num = np.zeros_like(arr)
        it = np.nditer(arr, flags=["multi_index"])
        for _ in it:
            i = it.multi_index
            old = arr[i]
            arr[i] = old + eps
            fp = f(q, k, v)
            arr[i] = old - eps
            fm = f(q, k, v)
            arr[i] = old
            num[i] = (fp - fm) / (2 * eps)
        np.testing.assert_allclose(grad, num, rtol=1e-5, atol=1e-7)


def test_mha_dropout_oracle_against_finite_difference():
    """The explicit-mask dropout of the oracle (attention.py:83 semantics) and its analytic backward."""
    rng = np.random.default_rng(1)
    N, Lq, Lk, d = 2, 3, 5, 4
    q, k, v = rng.normal(size=(N, Lq, d)), rng.normal(size=(N, Lk, d)), rng.normal(size=(N, Lk, d))
    g = rng.normal(size=(N, Lq, d))
    keep = rng.random((N, Lq, Lk)) > 0.3
    out, attn = oracle.mha_core_forward(q, k, v, keep=keep, keep_prob=0.7)
    np.testing.assert_allclose(out, np.einsum("nqk,nkd->nqd", attn * keep / 0.7, v))
    gq, gk, gv = oracle.mha_core_backward(q, k, v, g, keep=keep, keep_prob=0.7)

    def f(qq, kk, vv):
        return (mha_oracle.mha_core_forward(qq, kk, vv, keep=keep, keep_prob=0.7)[0] * g).sum()
    eps = 1e-6
    for arr, grad in ((q, gq), (k, gk), (v, gv)):
        num = np.zeros_like(arr)
        it = np.nditer(arr, flags=["multi_index"])
        for _ in it:
            i = it.multi_index
            old = arr[i]
            arr[i] = old + eps
            fp = f(q, k, v)
            arr[i] = old - eps
            fm = f(q, k, v)
            arr[i] = old
            num[i] = (fp - fm) / (2 * eps)
        np.testing.assert_allclose(grad, num, rtol=1e-5, atol=1e-7)


def test_masks():
    g = load_golden("masks")
    np.testing.assert_array_equal(oracle.sequence_mask(g["lens"]), g["sequence_mask"])
    np.testing.assert_array_equal(oracle.sequence_mask(g["lens"], 7), g["sequence_mask_7"])
    np.testing.assert_array_equal(oracle.get_attn_pad_mask(g["lens"], 3), g["attn_pad_mask"])
    np.testing.assert_array_equal(oracle.get_subsequent_mask(g["seq"]), g["subsequent_mask"])
    np.testing.assert_array_equal(oracle.get_attn_key_pad_mask(g["seq"], g["seq"], 0), g["key_pad_mask"])


# ---- the torch-CPU port used as the timed CPU baseline must agree with the same goldens ----
def test_torch_port_matches_goldens():
    import torch
    from oracle import torch_port
    for case in ("a", "c", "d"):
        hidden, alphas = torch.as_tensor(CIF[case + "_hidden"]), torch.as_tensor(CIF[case + "_alphas"])
        out = torch_port.cif_loop(hidden, alphas, float(CIF[case + "_thr"]))
        assert torch.equal(out, torch.as_tensor(CIF[case + "_out"]))
    for case in ("a", "b", "f"):
        loss = torch_port.ctc_mean_loss(torch.as_tensor(CTC[case + "_logits"]), torch.as_tensor(CTC[case + "_in_len"]),
                                        torch.as_tensor(CTC[case + "_targets"]))
        np.testing.assert_allclose(loss.numpy(), CTC[case + "_loss"], rtol=1e-6)
    w = {k: torch.as_tensor(v) for k, v in _weights().items()}
    for case in MHA_CASES:
        mask = MHA[case + "_mask"]
        mask = None if mask.size == 0 else torch.as_tensor(mask.astype(bool))
        q, kv = torch.as_tensor(MHA[case + "_q"]), torch.as_tensor(MHA[case + "_kv"])
        y, attn = torch_port.attention_block(q, kv, kv, w, 2, mask)
        np.testing.assert_allclose(y.numpy(), MHA[case + "_y"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(attn.numpy(), MHA[case + "_attn"], rtol=1e-5, atol=1e-7)
    g = load_golden("cif_glue")
    _num, num, scaled = torch_port.scale_alphas(torch.as_tensor(g["alpha"]), torch.as_tensor(g["targets"]),
                                                torch.as_tensor(g["rand"]))
    assert torch.equal(scaled, torch.as_tensor(g["scaled"])) and torch.equal(_num, torch.as_tensor(g["_num"]))
