"""tcgen05 attention core vs the oracle, the reference-generated golden vectors and a
plain torch fp32 reference of the same op (bf16 tolerance: rtol 2e-2 of the output scale)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import mha_oracle
from conftest import load_golden
from helpers import pkg, to_np

pytestmark = pytest.mark.gpu

MHA = load_golden("mha")
CASES = ["self_pad", "self_causal", "cross", "nomask"]


@pytest.fixture(autouse=True, params=[(0, 0), (3, 2), (21, 0), (40, 0), (0, 4), (0, 5)],
                ids=["auto", "one_tile_kernel_bwd8w", "two_tile_kernel", "persistent_two_tile_kernel", "bwd_16_warps_flush_dq", "bwd_one_cta_per_item"])
def fwd_variant(request):
    """The two forward kernels (auto picks by shape / dropout; 3 = always one 128-query tile per CTA, two threads per
    row; 21 = always two tiles per CTA, one thread per row; 40 = always the persistent kernel) and the softmax-backward
    layouts (default: persistent kernel, 16 warps + 4 dQ warps; 5 = the same with one CTA per key tile; 4 = 16 warps that
    flush dQ themselves; 2 = 8 warps)."""
    lib = pkg("_lib")
    lib.set_option("mha_variant", request.param[0])
    lib.set_option("mha_bwd_groups", request.param[1])
    yield request.param
    lib.set_option("mha_variant", 0)
    lib.set_option("mha_bwd_groups", 0)


def _torch_core(q, k, v, mask=None, scale=None):
    """fp32 reference of the core on [B,L,H,D] tensors (bmm / scale / masked_fill / softmax / bmm)."""
    B, Lq, H, D = q.shape
    scale = scale or 1.0 / D ** 0.5
    qh, kh, vh = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))     # [B,H,L,D]
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if mask is not None:
        s = s.masked_fill(mask[:, None].bool(), float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, vh).permute(0, 2, 1, 3)


def _rand_qkv(B, Lq, Lk, H, seed, std=1.0):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(B, Lq, H, 64, generator=g) * std).cuda().to(torch.bfloat16)
    k = (torch.randn(B, Lk, H, 64, generator=g) * std).cuda().to(torch.bfloat16)
    v = torch.randn(B, Lk, H, 64, generator=g).cuda().to(torch.bfloat16)
    return q, k, v


def _close(a, b, tol=2e-2, elementwise=True):
    """bf16 bar of north_star: rtol 2e-2.  Two readings, both asserted: the worst element against the output scale, and
    every element against its own magnitude with an absolute floor of one bf16 ulp of the scale (elements that cancel to
    ~0 cannot be held to a relative bound)."""
    a, b = a.float(), b.float()
    scale = b.abs().max().item() + 1e-6
    err = (a - b).abs()
    assert err.max().item() <= tol * scale, (err.max().item(), scale)
    if elementwise:
        assert (err <= tol * b.abs() + 2.0 ** -7 * scale).all(), ((err - tol * b.abs()).max().item(), scale)


@pytest.mark.parametrize("B,Lq,Lk,H", [(2, 128, 128, 2), (1, 21, 21, 8), (3, 167, 167, 8), (2, 16, 300, 4),
                                       (2, 300, 40, 4), (1, 512, 512, 2), (2, 129, 257, 3)])
def test_core_no_mask(B, Lq, Lk, H):
    ops = pkg("ops")
    q, k, v = _rand_qkv(B, Lq, Lk, H, seed=Lq + Lk)
    out = ops.mha_core(q, k, v)
    _close(out, _torch_core(q, k, v))


@pytest.mark.parametrize("B,L,H", [(3, 167, 8), (2, 300, 2), (4, 64, 4)])
def test_core_key_padding_lengths(B, L, H):
    ops = pkg("ops")
    q, k, v = _rand_qkv(B, L, L, H, seed=L)
    kv_len = torch.randint(L // 2, L + 1, (B,), generator=torch.Generator().manual_seed(1))
    kv_len[0] = L
    mask = (torch.arange(L)[None, None, :] >= kv_len[:, None, None]).expand(B, L, L).cuda()
    ref = _torch_core(q, k, v, mask)
    _close(ops.mha_core(q, k, v, kv_len=kv_len.cuda()), ref)       # structured form
    _close(ops.mha_core(q, k, v, mask=mask), ref)                   # dense form of the same mask


@pytest.mark.parametrize("B,L,H", [(2, 151, 8), (1, 400, 2), (2, 1100, 2)])
def test_core_causal(B, L, H):
    ops = pkg("ops")
    q, k, v = _rand_qkv(B, L, L, H, seed=L + 7)
    kv_len = torch.tensor([L] + [L - 9] * (B - 1))
    causal = torch.triu(torch.ones(L, L, dtype=torch.bool), diagonal=1)[None].expand(B, L, L)
    pad = (torch.arange(L)[None, None, :] >= kv_len[:, None, None]).expand(B, L, L)
    mask = (causal | pad).cuda()
    ref = _torch_core(q, k, v, mask)
    _close(ops.mha_core(q, k, v, kv_len=kv_len.cuda(), causal=True), ref)
    _close(ops.mha_core(q, k, v, mask=mask), ref)


def test_fully_masked_row_is_nan_like_reference():
    ops = pkg("ops")
    q, k, v = _rand_qkv(1, 8, 8, 1, seed=3)
    mask = torch.zeros(1, 8, 8, dtype=torch.bool)
    mask[0, 5, :] = True
    out = ops.mha_core(q, k, v, mask=mask.cuda())
    assert torch.isnan(out[0, 5]).all() and not torch.isnan(out[0, 4]).any()


@pytest.mark.parametrize("case", CASES)
def test_module_against_reference_golden(case):
    att = pkg("transformer.attention")
    m = att.MultiheadAttention(32, 2, 64, 64, dropout=0.1, return_attn=True).cuda().eval()
    m.load_state_dict({k[2:]: torch.as_tensor(MHA[k]) for k in MHA.files if k.startswith("w_")})
    q = torch.as_tensor(MHA[case + "_q"]).cuda()
    kv = torch.as_tensor(MHA[case + "_kv"]).cuda()
    mask = MHA[case + "_mask"]
    mask = None if mask.size == 0 else torch.as_tensor(mask.astype(bool)).cuda()
    y, attn = m(q, kv, kv, mask=mask)
    ref_y = MHA[case + "_y"]
    # bf16 core inside an fp32 block: rtol 2e-2 of the output scale
    assert np.abs(to_np(y) - ref_y).max() <= 2e-2 * np.abs(ref_y).max()
    ref_attn = MHA[case + "_attn"]                       # head-major rows: head * B + b
    assert attn.shape == ref_attn.shape
    assert np.abs(to_np(attn) - ref_attn).max() <= 2e-2


def test_ctcmodel_twin_constructor_order():
    catt = pkg("ctcModel.attention")
    m = catt.MultiHeadAttention(2, 32, 64, 64, dropout=0.0)
    assert m.n_head == 2 and m.w_qs.weight.shape == (128, 32)
    assert sorted(k for k, _ in m.named_parameters()) == sorted(
        ["w_qs.weight", "w_qs.bias", "w_ks.weight", "w_ks.bias", "w_vs.weight", "w_vs.bias",
         "fc.weight", "fc.bias", "layer_norm.weight", "layer_norm.bias"])


# ---------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------
def _grads_ref(q, k, v, g, mask=None):
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    out = _torch_core(qf, kf, vf, mask)
    out.backward(g.float())
    return qf.grad, kf.grad, vf.grad


def _grads_ours(q, k, v, g, **kw):
    ops = pkg("ops")
    qo, ko, vo = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = ops.mha_core(qo, ko, vo, **kw)
    out.backward(g)
    return qo.grad, ko.grad, vo.grad


@pytest.mark.parametrize("B,Lq,Lk,H", [(2, 128, 128, 2), (1, 21, 21, 8), (3, 167, 167, 8), (2, 16, 300, 4),
                                       (2, 300, 40, 4), (1, 512, 512, 2), (2, 129, 257, 3)])
def test_core_backward_no_mask(B, Lq, Lk, H):
    q, k, v = _rand_qkv(B, Lq, Lk, H, seed=Lq * 3 + Lk)
    g = torch.randn(B, Lq, H, 64, generator=torch.Generator().manual_seed(2)).cuda().to(torch.bfloat16)
    ours = _grads_ours(q, k, v, g)
    ref = _grads_ref(q, k, v, g)
    for o, r in zip(ours, ref):
        _close(o, r, tol=3e-2)


@pytest.mark.parametrize("B,L,H,causal", [(3, 167, 8, False), (2, 300, 2, True), (2, 151, 4, True), (4, 64, 4, False)])
def test_core_backward_masks(B, L, H, causal):
    q, k, v = _rand_qkv(B, L, L, H, seed=L + 11)
    g = torch.randn(B, L, H, 64, generator=torch.Generator().manual_seed(4)).cuda().to(torch.bfloat16)
    kv_len = torch.randint(L // 2, L + 1, (B,), generator=torch.Generator().manual_seed(1))
    kv_len[0] = L
    mask = (torch.arange(L)[None, None, :] >= kv_len[:, None, None]).expand(B, L, L)
    if causal:
        mask = mask | torch.triu(torch.ones(L, L, dtype=torch.bool), diagonal=1)[None]
    mask = mask.cuda()
    ref = _grads_ref(q, k, v, g, mask)
    for o, r in zip(_grads_ours(q, k, v, g, kv_len=kv_len.cuda(), causal=causal), ref):
        _close(o, r, tol=3e-2)
    for o, r in zip(_grads_ours(q, k, v, g, mask=mask), ref):
        _close(o, r, tol=3e-2)
    # keys beyond kv_len receive exactly zero gradient
    gk = _grads_ours(q, k, v, g, kv_len=kv_len.cuda(), causal=causal)[1]
    for b in range(B):
        assert not gk[b, int(kv_len[b]):].float().abs().any()


@pytest.mark.parametrize("case", CASES)
def test_module_gradients_against_reference_golden(case):
    att = pkg("transformer.attention")
    m = att.MultiheadAttention(32, 2, 64, 64, dropout=0.1).cuda().eval()
    m.load_state_dict({k[2:]: torch.as_tensor(MHA[k]) for k in MHA.files if k.startswith("w_")})
    q = torch.as_tensor(MHA[case + "_q"]).cuda().requires_grad_(True)
    kv = torch.as_tensor(MHA[case + "_kv"]).cuda().requires_grad_(True)
    mask = MHA[case + "_mask"]
    mask = None if mask.size == 0 else torch.as_tensor(mask.astype(bool)).cuda()
    y, _ = m(q, kv, kv, mask=mask)
    y.backward(torch.as_tensor(MHA[case + "_g_y"]).cuda())
    for name, got in (("g_q", q.grad), ("g_kv", kv.grad)):
        ref = MHA[case + "_" + name]
        assert np.abs(to_np(got) - ref).max() <= 3e-2 * np.abs(ref).max(), name
    # d/d(w_ks.bias) is analytically zero (a constant added to every key shifts all scores of a
    # query equally); the reference leaves ~1e-7 of fp32 noise there, bf16 leaves ~1e-3.  Judge it
    # against the scale of the sibling bias gradient instead of against zero.
    bias_scale = np.abs(MHA[case + "_gw_w_qs.bias"]).max()
    for pn, p in m.named_parameters():
        ref = MHA[case + "_gw_" + pn]
        scale = max(np.abs(ref).max(), bias_scale if pn == "w_ks.bias" else 0.0)
        assert np.abs(to_np(p.grad) - ref).max() <= 3e-2 * scale + 1e-6, pn


# ---- dropout on the probabilities (reference attention.py:83) -----------------------------------
def _torch_core_dropout(q, k, v, keep, keep_prob, mask=None, scale=None):
    """fp32 reference with an explicit keep mask [B,H,Lq,Lk]: softmax, drop, rescale, product."""
    B, Lq, H, D = q.shape
    scale = scale or 1.0 / D ** 0.5
    qh, kh, vh = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if mask is not None:
        s = s.masked_fill(mask[:, None].bool(), float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = p * keep.float() / keep_prob
    return torch.matmul(p, vh).permute(0, 2, 1, 3)


def test_dropout_mask_statistics_and_determinism():
    ops = pkg("ops")
    B, H, Lq, Lk = 2, 4, 200, 333
    keep, kp = ops.mha_dropout_keep(B, H, Lq, Lk, 0.1, seed=1234)
    assert abs(kp - 230.0 / 256.0) < 1e-7          # 0.1 is quantised to 26/256
    n = keep.numel()
    frac = keep.float().mean().item()
    assert abs(frac - kp) < 5 * (kp * (1 - kp) / n) ** 0.5
    # per row and per column too (no stripes)
    assert (keep.float().mean(-1) - kp).abs().max().item() < 0.12
    assert (keep.float().mean(-2) - kp).abs().max().item() < 0.12
    keep2, _ = ops.mha_dropout_keep(B, H, Lq, Lk, 0.1, seed=1234)
    keep3, _ = ops.mha_dropout_keep(B, H, Lq, Lk, 0.1, seed=1235)
    assert torch.equal(keep, keep2)
    assert 0.7 < (keep ^ keep3).float().mean().item() / (2 * kp * (1 - kp)) < 1.3   # independent masks
    keep0, kp0 = ops.mha_dropout_keep(1, 1, 8, 40, 0.0, seed=7)
    assert kp0 == 1.0 and keep0.all()


@pytest.mark.parametrize("B,Lq,Lk,H,p,causal", [(2, 128, 128, 2, 0.1, False), (1, 167, 167, 8, 0.1, True),
                                               (2, 70, 300, 4, 0.3, False), (1, 257, 129, 3, 0.5, False)])
def test_dropout_forward_backward_match_explicit_mask(B, Lq, Lk, H, p, causal):
    ops = pkg("ops")
    q, k, v = _rand_qkv(B, Lq, Lk, H, seed=11 + Lq)
    kv_len = torch.tensor([Lk - 7 * i for i in range(B)], dtype=torch.int32, device="cuda")
    seed = 99 + Lk
    keep, kp = ops.mha_dropout_keep(B, H, Lq, Lk, p, seed)
    dense = torch.arange(Lk, device="cuda")[None, None, :] >= kv_len[:, None, None]
    dense = dense.expand(B, Lq, Lk)
    if causal:
        dense = dense | torch.triu(torch.ones(Lq, Lk, dtype=torch.bool, device="cuda"), 1)[None]
    qs, ks, vs = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = ops.mha_core(qs, ks, vs, kv_len=kv_len, causal=causal, dropout_p=p, seed=seed)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = _torch_core_dropout(qr, kr, vr, keep, kp, mask=dense)
    _close(out, ref)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)).cuda()
    out.backward(g.to(out.dtype))
    ref.backward(g)
    _close(qs.grad, qr.grad, 3e-2)
    _close(ks.grad, kr.grad, 3e-2)
    _close(vs.grad, vr.grad, 3e-2)
    # same seed -> same result, other seed -> a different mask
    out2 = ops.mha_core(q, k, v, kv_len=kv_len, causal=causal, dropout_p=p, seed=seed)
    out3 = ops.mha_core(q, k, v, kv_len=kv_len, causal=causal, dropout_p=p, seed=seed + 1)
    assert torch.equal(out.detach(), out2)
    assert not torch.equal(out.detach(), out3)


def test_dropout_zero_is_the_plain_kernel_and_expectation_is_unbiased():
    ops = pkg("ops")
    q, k, v = _rand_qkv(1, 128, 256, 2, seed=5)
    plain = ops.mha_core(q, k, v)
    assert torch.equal(plain, ops.mha_core(q, k, v, dropout_p=0.0, seed=123))
    acc = torch.zeros_like(plain, dtype=torch.float32)
    n = 64
    for s in range(n):
        acc += ops.mha_core(q, k, v, dropout_p=0.25, seed=1000 + s).float()
    # mean over 64 masks approaches the undropped output: error ~ sqrt(p/(1-p)/n) of a row's spread
    _close(acc / n, plain, 0.25, elementwise=False)      # a statistical bound on a mean over random masks


def test_module_applies_attention_dropout_only_in_training():
    MHA = pkg("transformer.attention").MultiheadAttention
    torch.manual_seed(0)
    m = MHA(128, 2, dropout=0.2).cuda()
    x = torch.randn(2, 50, 128, device="cuda")
    m.eval()
    y0, _ = m(x, x, x)
    y1, _ = m(x, x, x)
    assert torch.equal(y0, y1)
    m.train()
    torch.manual_seed(1)
    t0, _ = m(x, x, x)
    torch.manual_seed(1)
    t1, _ = m(x, x, x)
    torch.manual_seed(2)
    t2, _ = m(x, x, x)
    assert torch.equal(t0, t1) and not torch.equal(t0, t2) and not torch.equal(t0, y0)
    t0.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())


@pytest.mark.parametrize("causal", [False, True])
def test_long_sequence_microbench_shape(causal):
    """L = 2048 (the SURVEY 8d microbench length, 16 key blocks, 8 two-tile CTAs per head) with ragged key lengths and
    the causal mask, forward and backward against the fp32 reference."""
    B, L, H = 2, 2048, 2
    ops = pkg("ops")
    q, k, v = _rand_qkv(B, L, L, H, seed=77, std=0.5)
    kv_len = torch.tensor([2048, 1531], dtype=torch.int32).cuda()
    mask = torch.arange(L, device="cuda")[None, None, :] >= kv_len[:, None, None]
    mask = mask.expand(B, L, L)
    if causal:
        mask = mask | torch.triu(torch.ones(L, L, dtype=torch.bool, device="cuda"), diagonal=1)[None]
    qr, kr, vr = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    out = ops.mha_core(qr, kr, vr, kv_len=kv_len, causal=causal)
    ref_in = [t.detach().float().requires_grad_(True) for t in (q, k, v)]
    ref = _torch_core(*ref_in, mask=mask)
    _close(out, ref)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(78)).cuda()
    out.backward(g.to(out.dtype))
    ref.backward(g)
    for got, want in zip((qr, kr, vr), ref_in):
        _close(got.grad, want.grad, tol=3e-2)


@pytest.mark.parametrize("B,L,H,causal", [(16, 700, 8, False), (5, 1200, 8, True), (40, 300, 4, False), (3, 2048, 8, True)])
def test_persistent_kernel_is_bitwise_the_two_tile_kernel(B, L, H, causal):
    """The persistent forward kernel walks several (batch, head, 256-query) items per CTA - Q double-buffered, the next
    item's first product issued under the current item's last block, the epilogue deferred behind the next item's first
    block.  Same arithmetic in the same order as the one-item-per-CTA kernel: outputs and log-sum-exps must be
    bit-identical, on shapes with more items than SMs (2.6, 1.4, 2.2 items per CTA), ragged key lengths, a query tile
    that is half empty (L % 256 <= 128: items with ONE tile) and the causal mask (tiles of one item with different
    numbers of key blocks)."""
    lib = pkg("_lib")
    q, k, v = _rand_qkv(B, L, L, H, seed=91, std=0.7)
    g = torch.Generator().manual_seed(92)
    kv_len = torch.randint(L // 3, L + 1, (B,), generator=g).to(torch.int32).cuda()
    outs = []
    for variant in (21, 40):
        lib.set_option("mha_variant", variant)
        out = torch.full_like(q, float("nan"))
        lse = torch.full((B, H, L), float("nan"), device="cuda")
        lib.check(lib.lib().asr_mha_fwd_bf16(lib.ptr(q), lib.ptr(k), lib.ptr(v), lib.ptr(kv_len), None, int(causal), B, H, L, L, 64,
                                             0.125, lib.ptr(out), lib.ptr(lse), lib.stream_ptr()), "asr_mha_fwd_bf16")
        outs.append((out, lse))
    torch.cuda.synchronize()
    assert torch.isfinite(outs[1][0].float()).all() and torch.isfinite(outs[1][1]).all()
    assert torch.equal(outs[0][0].view(torch.int16), outs[1][0].view(torch.int16))
    assert torch.equal(outs[0][1], outs[1][1])


def test_broadcastable_masks_are_expanded_and_bad_shapes_rejected():
    """The reference's masked_fill accepts masks that broadcast to [B,Lq,Lk]; the kernel takes a raw pointer, so the
    wrapper expands them - and refuses shapes that do not fit."""
    ops = pkg("ops")
    B, Lq, Lk, H = 3, 40, 70, 2
    q, k, v = _rand_qkv(B, Lq, Lk, H, seed=5)
    key_pad = torch.zeros(B, 1, Lk, dtype=torch.bool, device="cuda")
    key_pad[0, 0, 50:] = True
    key_pad[2, 0, 13:] = True
    full = key_pad.expand(B, Lq, Lk).contiguous()
    a = ops.mha_core(q, k, v, mask=key_pad)
    b = ops.mha_core(q, k, v, mask=full)
    assert torch.equal(a, b)
    one = torch.triu(torch.ones(1, Lq, Lk, dtype=torch.bool, device="cuda"), diagonal=1)
    assert torch.equal(ops.mha_core(q, k, v, mask=one), ops.mha_core(q, k, v, mask=one.expand(B, Lq, Lk).contiguous()))
    assert torch.equal(ops.mha_probs(q, k, mask=key_pad), ops.mha_probs(q, k, mask=full))
    with pytest.raises(ValueError):
        ops.mha_core(q, k, v, mask=torch.zeros(B, Lq, Lk + 1, dtype=torch.bool, device="cuda"))
    with pytest.raises(ValueError):
        ops.mha_core(q, k, v, mask=torch.zeros(Lq, Lk, dtype=torch.bool, device="cuda"))
    with pytest.raises(ValueError):
        ops.mha_core(q, k, v, kv_len=torch.tensor([1, 2], device="cuda"))
    with pytest.raises(ValueError):
        ops.mha_core(q, k[:, :, :1], v)


def test_device_side_dropout_seed_matches_the_host_seed_and_advances():
    """Dropout seeds read on the device (CUDA-graph training steps): the mask of (*seed_dev + seed_add) is the mask of
    that number passed as a host seed; advancing the device word changes it; backward regenerates it."""
    ops = pkg("ops")
    B, L, H, p = 2, 167, 4, 0.1
    q, k, v = _rand_qkv(B, L, L, H, seed=31)
    ds = ops.DropoutSeed("cuda", seed=1000)
    with ops.device_dropout_seed(ds):
        qa, ka, va = (t.clone().requires_grad_(True) for t in (q, k, v))
        out_dev = ops.mha_core(qa, ka, va, dropout_p=p)
        add1 = (1 * ops.DropoutSeed._STRIDE) & 0xFFFFFFFFFFFFFFFF
        out_dev.float().sum().backward()
    host_seed = (1000 + add1) & 0xFFFFFFFFFFFFFFFF
    qb, kb, vb = (t.clone().requires_grad_(True) for t in (q, k, v))
    out_host = ops.mha_core(qb, kb, vb, dropout_p=p, seed=host_seed)
    out_host.float().sum().backward()
    assert torch.equal(out_dev, out_host)
    assert torch.equal(qa.grad, qb.grad) and torch.equal(ka.grad, kb.grad) and torch.equal(va.grad, vb.grad)
    ds.advance()
    with ops.device_dropout_seed(ds):
        out_next = ops.mha_core(q, k, v, dropout_p=p)
    assert not torch.equal(out_next, out_dev)
    assert torch.equal(out_next, ops.mha_core(q, k, v, dropout_p=p, seed=(1001 + add1) & 0xFFFFFFFFFFFFFFFF))


def test_attn_is_returned_by_default_like_the_reference():
    att = pkg("transformer.attention")
    m = att.MultiheadAttention(64, 2, dropout=0.0).cuda().eval()
    x = torch.randn(2, 9, 64, device="cuda")
    out, attn = m(x, x, x)
    assert attn is not None and tuple(attn.shape) == (2 * 2, 9, 9)
    torch.testing.assert_close(attn.sum(-1), torch.ones(4, 9, device="cuda"), rtol=1e-3, atol=1e-3)
    m2 = att.MultiheadAttention(64, 2, dropout=0.0, return_attn=False).cuda().eval()
    assert m2(x, x, x)[1] is None
