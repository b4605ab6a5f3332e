"""CIF weight producer (assigner tail + scaling) vs the reference-generated golden
vector, the oracle and torch's own ops on the device."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden
from helpers import pkg, to_np

pytestmark = pytest.mark.gpu


def _torch_ref(x, w, b, lens, num_noise):
    """The reference's op sequence (attentionAssigner.py:36-40, cif_model.py:43-48) on this device."""
    alphas = torch.sigmoid(torch.nn.functional.linear(x, w.view(1, -1), b).squeeze(-1))
    mask = (torch.arange(x.size(1), device=x.device)[None, :] < lens[:, None]).float()
    alpha = alphas * mask
    _num = alpha.sum(-1)
    if num_noise is None:
        return alpha, _num
    return alpha * (num_noise / _num)[:, None], _num


def test_golden_forward_backward():
    ops = pkg("ops")
    g = load_golden("assigner_tail")
    x = torch.as_tensor(g["x"]).cuda().requires_grad_(True)
    w = torch.as_tensor(g["w"]).cuda().requires_grad_(True)
    b = torch.as_tensor(g["b"]).cuda().requires_grad_(True)
    lens = torch.as_tensor(g["lens"]).cuda()
    alpha, num = ops.cif_alpha(x, w, b, lens, torch.as_tensor(g["num_noise"]).cuda())
    np.testing.assert_allclose(to_np(num), g["_num"], rtol=1e-5)
    np.testing.assert_allclose(to_np(alpha), g["scaled"], rtol=1e-5, atol=1e-7)
    assert not to_np(alpha)[1, 17:].any() and not to_np(alpha)[3, 1:].any()
    ((alpha * torch.as_tensor(g["g_alpha"]).cuda()).sum() + (num * torch.as_tensor(g["g_num"]).cuda()).sum()).backward()
    for got, ref in ((x.grad, g["g_x"]), (w.grad, g["g_w"]), (b.grad, g["g_b"])):
        assert got.shape == tuple(ref.shape)
        assert np.abs(to_np(got) - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("B,T,D,scaled", [(3, 50, 512, True), (2, 33, 320, True), (4, 17, 100, True), (2, 40, 63, False),
                                         (1, 1, 8, True), (5, 300, 256, False)])
def test_random_vs_oracle_and_torch(B, T, D, scaled):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(B * 1000 + T + D)
    x = torch.randn(B, T, D, generator=gen)
    w = torch.randn(D, generator=gen) * (2.0 / D ** 0.5)
    b = torch.randn(1, generator=gen) * 0.1
    lens = torch.randint(1, T + 1, (B,), generator=gen)
    lens[0] = T
    noise = (torch.randint(1, 20, (B,), generator=gen).float() + torch.rand(B, generator=gen) - 0.5) if scaled else None
    g_alpha = torch.randn(B, T, generator=gen)
    g_num = torch.randn(B, generator=gen)
    xs, ws, bs = (t.clone().cuda().requires_grad_(True) for t in (x, w, b))
    alpha, num = ops.cif_alpha(xs, ws, bs, lens.cuda(), noise.cuda() if scaled else None)
    ((alpha * g_alpha.cuda()).sum() + (num * g_num.cuda()).sum()).backward()
    # fp64 oracle
    a_raw, o_num, o_alpha = oracle.assigner_tail_forward(x.numpy(), w.numpy(), b.numpy(), lens.numpy(),
                                                         noise.numpy() if scaled else None)
    np.testing.assert_allclose(to_np(num), o_num, rtol=1e-5)
    assert np.abs(to_np(alpha) - o_alpha).max() <= 1e-5 * np.abs(o_alpha).max()
    o_gx, o_gw, o_gb = oracle.assigner_tail_backward(x.numpy(), w.numpy(), b.numpy(), lens.numpy(),
                                                     noise.numpy() if scaled else None, g_alpha.numpy(), g_num.numpy())
    assert np.abs(to_np(xs.grad) - o_gx).max() <= 1e-5 * np.abs(o_gx).max()
    assert np.abs(to_np(ws.grad) - o_gw).max() <= 1e-5 * np.abs(o_gw).max() + 1e-7
    assert abs(float(bs.grad) - o_gb) <= 1e-5 * abs(o_gb) + 1e-6
    # padded frames: zero weight, zero gradient rows
    for i in range(B):
        assert not to_np(alpha)[i, int(lens[i]):].any()
        assert not to_np(xs.grad)[i, int(lens[i]):].any()
    # torch's own op sequence on the same device is no closer to fp64 than we are (x2)
    xr, wr, br = (t.clone().cuda().requires_grad_(True) for t in (x, w, b))
    r_alpha, r_num = _torch_ref(xr, wr, br, lens.cuda(), noise.cuda() if scaled else None)
    e_ours = np.abs(to_np(alpha) - o_alpha).max()
    e_ref = np.abs(to_np(r_alpha) - o_alpha).max()
    assert e_ours <= max(2 * e_ref, 1e-6 * np.abs(o_alpha).max())


def test_deterministic_and_no_grad_path():
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(6, 200, 512, generator=gen).cuda()
    w = torch.randn(512, generator=gen).cuda() * 0.05
    b = torch.zeros(1).cuda()
    lens = torch.tensor([200, 150, 200, 77, 1, 199]).cuda()
    noise = torch.tensor([9.3, 7.1, 12.2, 3.4, 0.7, 8.8]).cuda()
    outs = []
    for _ in range(2):
        xs = x.clone().requires_grad_(True)
        ws = w.clone().requires_grad_(True)
        alpha, num = ops.cif_alpha(xs, ws, b, lens, noise)
        alpha.square().sum().backward()
        outs.append((alpha.detach(), num.detach(), xs.grad, ws.grad))
    for a, c in zip(outs[0], outs[1]):
        assert torch.equal(a, c)                      # fixed-order reductions: bit-reproducible
    np.testing.assert_allclose(to_np(outs[0][0].sum(-1)), to_np(noise), rtol=1e-5)   # scaled weights sum to num_noise
    with torch.no_grad():
        alpha2, _ = ops.cif_alpha(x, w, b, lens, noise)
    assert torch.equal(alpha2, outs[0][0])


def test_cif_model_with_fused_alpha_trains():
    """CIF_Model(fused_alpha=True): same quantity / losses as the unfused shell up to fp32 reordering of
    the weights, finite gradients into the assigner."""
    import test_model_shell as tms
    tl = pkg("transformer.loss")
    model = tms._build()
    model.load_state_dict(tms._golden_state(), strict=False)
    model = model.cuda().eval()
    feats = torch.as_tensor(tms.G["feats"]).cuda()
    lens = torch.as_tensor(tms.G["lens"]).cuda()
    targets = torch.as_tensor(tms.G["targets"]).cuda()
    res = []
    for fused in (False, True):
        model.fused_alpha = fused
        model.zero_grad()
        torch.manual_seed(int(tms.G["rand_seed"]))
        ctc_logits, len_ctc, _num, num, logits = model(feats, lens, targets)
        qua, ctc, ce = tl.cal_ctc_qua_ce_loss(ctc_logits, len_ctc, _num, num, logits, targets, smoothing=0.1)
        (0.001 * qua + ctc + ce).backward()
        res.append((to_np(_num), float(qua.detach()), float(ctc.detach()), float(ce.detach()), logits.shape,
                    to_np(model.assigner.linear.weight.grad)))
    np.testing.assert_allclose(res[1][0], res[0][0], rtol=1e-5)
    np.testing.assert_allclose(res[1][1], res[0][1], rtol=1e-4)
    np.testing.assert_allclose(res[1][2], res[0][2], rtol=1e-5)
    assert res[1][4] == res[0][4]
    assert np.isfinite(res[1][5]).all() and np.abs(res[1][5]).sum() > 0


# ---- low-frame-rate stacking ------------------------------------------------------
def test_lfr_golden_bit_exact():
    ops = pkg("ops")
    g = load_golden("lfr")
    for name in sorted({k.split("_")[0] for k in g.files}):
        m, n = (int(v) for v in g[name + "_mn"])
        x = torch.as_tensor(g[name + "_x"]).cuda()[None]
        y, yl = ops.build_lfr_features(x, torch.tensor([x.size(1)]), m, n)
        ref = g[name + "_y"]
        assert int(yl[0]) == ref.shape[0] and tuple(y.shape[1:]) == ref.shape
        np.testing.assert_array_equal(to_np(y[0]).view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("B,T,D,m,n", [(5, 167, 80, 4, 3), (3, 50, 83, 4, 3), (2, 31, 40, 1, 1), (4, 64, 8, 3, 2)])
def test_lfr_ragged_batch_vs_oracle(B, T, D, m, n):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(T + D)
    x = torch.randn(B, T, D, generator=gen)
    lens = torch.randint(1, T + 1, (B,), generator=gen)
    lens[0] = T
    y, yl = ops.build_lfr_features(x.cuda(), lens, m, n)
    assert y.shape == (B, (T + n - 1) // n, m * D)
    for b in range(B):
        ref = oracle.build_lfr_features(x[b, :int(lens[b])].numpy(), m, n)
        assert int(yl[b]) == ref.shape[0]
        np.testing.assert_array_equal(to_np(y[b, :ref.shape[0]]).view(np.uint32), ref.view(np.uint32))
        assert not to_np(y[b, ref.shape[0]:]).any()          # zero beyond the utterance


# ---- SpecAugment -------------------------------------------------------------------
def test_spec_aug_golden():
    """The reference's spec_aug run on CPU with recorded draws (tests/golden/make_golden.py: make_spec_aug)."""
    from test_oracle_golden import _spec_aug_masks
    ops = pkg("ops")
    g = load_golden("spec_aug")
    for name in sorted({k.split("_")[0] for k in g.files}):
        masks = [torch.as_tensor(m) for m in _spec_aug_masks(g, name)]
        x = torch.as_tensor(g[name + "_x"]).cuda()
        y = ops.spec_aug_apply(x, torch.as_tensor(g[name + "_lens"]), *masks)
        assert y.data_ptr() == x.data_ptr()                     # in place, like the reference
        ref = g[name + "_y"]
        same = ref == g[name + "_x"]
        got = to_np(y)
        np.testing.assert_array_equal(got[same].view(np.uint32), ref[same].view(np.uint32))
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("B,T,V,R", [(6, 500, 320, 2), (3, 167, 80, 2), (2, 70, 33, 3), (4, 1600, 320, 1), (2, 64, 1024, 2)])
def test_spec_aug_ragged_vs_oracle(B, T, V, R):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(B * T + V)
    x = torch.randn(B, T, V, generator=gen)
    lens = torch.randint(T // 2, T + 1, (B,), generator=gen)
    lens[0] = T
    for b in range(B):
        x[b, int(lens[b]):] = 0.0
    fw = torch.randint(0, min(27, V), (R, B), generator=gen)
    f0 = (torch.rand(R, B, generator=gen) * (V - fw)).long()
    tw = torch.randint(0, 40, (R, B), generator=gen)
    t0 = (torch.rand(R, B, generator=gen) * (lens[None] - tw)).long()
    fw[0, 0] = 0                                               # an empty band and an empty span
    tw[-1, -1] = 0
    ref = oracle.spec_aug_apply(x.numpy(), lens.numpy(), f0.numpy(), fw.numpy(), t0.numpy(), tw.numpy())
    y = to_np(ops.spec_aug_apply(x.clone().cuda(), lens, f0, fw, t0, tw))
    same = ref == x.numpy()
    np.testing.assert_array_equal(y[same].view(np.uint32), ref[same].view(np.uint32))
    np.testing.assert_allclose(y, ref, rtol=1e-5, atol=3e-6)
    # run-to-run bit reproducibility (fixed summation order, no atomics)
    y2 = to_np(ops.spec_aug_apply(x.clone().cuda(), lens, f0, fw, t0, tw))
    np.testing.assert_array_equal(y.view(np.uint32), y2.view(np.uint32))


def test_spec_aug_drop_in_signature():
    """utils.spec_aug(padded_features, feature_lengths, config) as cif_model.py:33 calls it: in place, returns
    (features, lengths); its masks are the ones spec_aug_draw yields for the same generator state."""
    uu = pkg("utils.utils")
    B, T, V = 4, 120, 80
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(B, T, V, generator=gen).cuda()
    lens = torch.tensor([120, 100, 90, 64]).cuda()
    x0 = x.clone()
    torch.manual_seed(11)
    y, yl = uu.spec_aug(x, lens, "2-27-2-40")
    assert y is x and yl is lens
    torch.manual_seed(11)
    f0, fw, t0, tw = uu.spec_aug_draw(B, V, lens, "2-27-2-40", x.device)
    ref = oracle.spec_aug_apply(to_np(x0), to_np(lens), to_np(f0), to_np(fw), to_np(t0), to_np(tw))
    np.testing.assert_allclose(to_np(y), ref, rtol=1e-5, atol=3e-6)
    assert (to_np(y) != to_np(x0)).any()


def test_spec_aug_rejects_cpu_and_wide_features():
    ops = pkg("ops")
    z = torch.zeros(1, 1, dtype=torch.long)
    with pytest.raises((RuntimeError, ValueError, TypeError)):
        ops.spec_aug_apply(torch.zeros(1, 4, 8), torch.tensor([4]), z, z, z, z)
    with pytest.raises(RuntimeError):
        ops.spec_aug_apply(torch.zeros(1, 4, 1025).cuda(), torch.tensor([4]), z, z, z, z)
