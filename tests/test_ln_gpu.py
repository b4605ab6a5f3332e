"""dropout + residual + LayerNorm of the training step (csrc/ln.cu; SURVEY.md 8(f3): what torch runs for
/root/reference/src/transformer/module.py:50-52, attention.py:59-60, encoder.py:49) against torch in fp64 on the same inputs.
Tolerances: fp32 arithmetic with another summation order than torch's - 2e-6 of the output scale forward, 1e-5 of each
gradient's scale backward (column sums over up to 15 030 rows)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import pkg

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda().to(dtype)


def _reference(y, res, w, b, eps, keep=None, keep_prob=1.0):
    yd = y.detach().double().requires_grad_(True)
    rd = res.detach().double().requires_grad_(True) if res is not None else None
    wd, bd = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    z = yd if keep is None else yd * keep.double() / keep_prob
    if rd is not None:
        z = z + rd
    return F.layer_norm(z, (y.shape[-1],), wd, bd, eps), (yd, rd, wd, bd)


def _rel(got, want):
    return (got.double() - want).abs().max().item() / max(want.abs().max().item(), 1e-30)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("with_res", [True, False])
@pytest.mark.parametrize("M,D", [(1, 512), (7, 256), (1344, 512), (15030, 512), (1350, 1024), (2500, 256)])
def test_residual_layer_norm_matches_torch_fp64(M, D, with_res, dtype):
    ops = pkg("ops")
    y = _rand((M, D), 1, 2.0, dtype).requires_grad_(True)
    res = _rand((M, D), 2).requires_grad_(True) if with_res else None
    w = (_rand((D,), 3, 0.2) + 1.0).requires_grad_(True)
    b = _rand((D,), 4, 0.1).requires_grad_(True)
    g = _rand((M, D), 5)
    out = ops.residual_layer_norm(y, res, w, b, 1e-5)
    assert out.dtype == torch.float32 and out.shape == y.shape
    out.backward(g)
    ref, (yd, rd, wd, bd) = _reference(y, res, w, b, 1e-5)
    ref.backward(g.double())
    assert _rel(out, ref.detach()) <= 2e-6
    tol_y = 1e-5 if dtype == torch.float32 else 6e-3           # dy is rounded to bf16 when y is
    assert y.grad.dtype == dtype and _rel(y.grad, yd.grad) <= tol_y
    if with_res:
        assert _rel(res.grad, rd.grad) <= 1e-5
    assert _rel(w.grad, wd.grad) <= 1e-5 and _rel(b.grad, bd.grad) <= 1e-5
    # fixed summation orders: a second run gives the same bits
    y2, w2, b2 = y.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    out2 = ops.residual_layer_norm(y2, res.detach() if with_res else None, w2, b2, 1e-5)
    out2.backward(g)
    assert torch.equal(out, out2) and torch.equal(y.grad, y2.grad) and torch.equal(w.grad, w2.grad) and torch.equal(b.grad, b2.grad)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,D,p", [(1344, 512, 0.1), (300, 1024, 0.5), (4000, 256, 0.1)])
def test_dropout_mask_is_regenerated_in_the_backward(M, D, p, dtype):
    ops = pkg("ops")
    seed = 1234567 + M
    keep, keep_prob = ops.ln_dropout_keep(M, D, p, seed)
    assert abs(keep.float().mean().item() - keep_prob) <= 4.0 * (keep_prob * (1 - keep_prob) / (M * D)) ** 0.5 + 1e-3
    assert abs(keep_prob - (1 - p)) <= 1 / 256
    other, _ = ops.ln_dropout_keep(M, D, p, seed + 1)
    assert (other != keep).float().mean().item() > 0.05                  # another seed, another mask
    y = _rand((M, D), 1, 2.0, dtype).requires_grad_(True)
    res = _rand((M, D), 2).requires_grad_(True)
    w = (_rand((D,), 3, 0.2) + 1.0).requires_grad_(True)
    b = _rand((D,), 4, 0.1).requires_grad_(True)
    g = _rand((M, D), 5)
    out = ops.residual_layer_norm(y, res, w, b, 1e-5, dropout_p=p, training=True, seed=seed)
    out.backward(g)
    ref, (yd, rd, wd, bd) = _reference(y, res, w, b, 1e-5, keep, keep_prob)
    ref.backward(g.double())
    assert _rel(out, ref.detach()) <= 2e-6
    assert _rel(y.grad, yd.grad) <= (1e-5 if dtype == torch.float32 else 6e-3)
    assert ((y.grad == 0) | keep).all() and (y.grad[~keep] == 0).all()    # dropped elements get no gradient
    assert _rel(res.grad, rd.grad) <= 1e-5 and _rel(w.grad, wd.grad) <= 1e-5 and _rel(b.grad, bd.grad) <= 1e-5
    # evaluation mode: no dropout whatever p says
    ev = ops.residual_layer_norm(y.detach(), res.detach(), w.detach(), b.detach(), 1e-5, dropout_p=p, training=False)
    assert _rel(ev, _reference(y, res, w, b, 1e-5)[0].detach()) <= 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_row_scale_is_the_non_pad_mask_of_the_layers(dtype):
    """out = LayerNorm(dropout(y) + residual) * mask[row] (encoder.py:76-80, decoder.py:628-634 of the reference): masked rows
    are exact zeros and pass no gradient - not to y, not to the residual, not to gamma / beta."""
    ops = pkg("ops")
    M, D, p, seed = 1350, 512, 0.1, 99
    keep, keep_prob = ops.ln_dropout_keep(M, D, p, seed)
    mask = (torch.arange(M, device="cuda") % 15 < 11).reshape(90, 15, 1)
    y = _rand((90, 15, D), 1, 2.0, dtype).requires_grad_(True)
    res = _rand((90, 15, D), 2).requires_grad_(True)
    w = (_rand((D,), 3, 0.2) + 1.0).requires_grad_(True)
    b = _rand((D,), 4, 0.1).requires_grad_(True)
    g = _rand((90, 15, D), 5)
    out = ops.residual_layer_norm(y, res, w, b, 1e-5, dropout_p=p, seed=seed, row_scale=mask)
    out.backward(g)
    ref, (yd, rd, wd, bd) = _reference(y.reshape(M, D), res.reshape(M, D), w, b, 1e-5, keep, keep_prob)
    ref = ref * mask.reshape(M, 1).double()
    ref.backward(g.reshape(M, D).double())
    assert (out[~mask.expand_as(out)] == 0).all()
    assert _rel(out.reshape(M, D), ref.detach()) <= 2e-6
    assert _rel(y.grad.reshape(M, D), yd.grad) <= (1e-5 if dtype == torch.float32 else 6e-3)
    assert (res.grad[~mask.expand_as(out)] == 0).all()
    assert _rel(res.grad.reshape(M, D), rd.grad) <= 1e-5 and _rel(w.grad, wd.grad) <= 1e-5 and _rel(b.grad, bd.grad) <= 1e-5
    with pytest.raises(ValueError):
        ops.residual_layer_norm(y, res, w, b, row_scale=mask[:, :3])


def test_bf16_copy_and_the_gradient_that_comes_back_through_it():
    """bf16_copy=True: the kernel also writes its output in bf16 (what the next bf16 GEMM reads instead of a conversion
    kernel); the gradient arriving through that copy is added to the fp32 one inside the backward kernel."""
    ops = pkg("ops")
    M, D = 1350, 512
    y = _rand((M, D), 1, 2.0, torch.bfloat16).requires_grad_(True)
    res = _rand((M, D), 2).requires_grad_(True)
    w = (_rand((D,), 3, 0.2) + 1.0).requires_grad_(True)
    b = _rand((D,), 4, 0.1).requires_grad_(True)
    g32, g16 = _rand((M, D), 5), _rand((M, D), 6, dtype=torch.bfloat16)
    out = ops.residual_layer_norm(y, res, w, b, 1e-5, bf16_copy=True)
    out16 = ops.bf16_copy_of(out)
    assert out16 is not None and out16.dtype == torch.bfloat16 and torch.equal(out16, out.detach().bfloat16())
    assert ops.bf16_copy_of(out * 1.0) is None                     # any op makes a new tensor: no stale copies
    torch.autograd.backward([out, out16], [g32, g16])
    ref, (yd, rd, wd, bd) = _reference(y, res, w, b, 1e-5)
    ref.backward(g32.double() + g16.double())
    assert _rel(y.grad, yd.grad) <= 6e-3 and _rel(res.grad, rd.grad) <= 1e-5
    assert _rel(w.grad, wd.grad) <= 1e-5 and _rel(b.grad, bd.grad) <= 1e-5
    # only the copy is used downstream (a layer whose fp32 output feeds nothing else)
    y2 = y.detach().clone().requires_grad_(True)
    out = ops.residual_layer_norm(y2, res.detach(), w.detach(), b.detach(), 1e-5, bf16_copy=True)
    ops.bf16_copy_of(out).backward(g16)
    ref, (yd, _, _, _) = _reference(y, res, w, b, 1e-5)
    ref.backward(g16.double())
    assert _rel(y2.grad, yd.grad) <= 6e-3


def test_device_side_seed_matches_the_host_seed_it_stands_for():
    """Inside device_dropout_seed(...) the kernels read *seed_dev + the call's constant (CUDA-graph replays): the same mask
    as passing that sum from the host, and a new one after advance()."""
    ops = pkg("ops")
    M, D, p = 513, 512, 0.25
    y, res = _rand((M, D), 1), _rand((M, D), 2)
    w, b = _rand((D,), 3, 0.2) + 1.0, _rand((D,), 4, 0.1)
    ds = ops.DropoutSeed(torch.device("cuda"), seed=777)
    outs = []
    for _ in range(2):
        yg = y.clone().requires_grad_(True)
        with ops.device_dropout_seed(ds):
            o = ops.residual_layer_norm(yg, res, w, b, 1e-5, dropout_p=p)
        o.sum().backward()
        host_seed = (int(ds.tensor.item()) + ops.DropoutSeed._STRIDE) & 0xFFFFFFFFFFFFFFFF
        yh = y.clone().requires_grad_(True)
        oh = ops.residual_layer_norm(yh, res, w, b, 1e-5, dropout_p=p, seed=host_seed)
        oh.sum().backward()
        assert torch.equal(o, oh) and torch.equal(yg.grad, yh.grad)
        outs.append(o.detach())
        ds.advance()
    assert not torch.equal(outs[0], outs[1])


def test_rejects_what_the_kernel_does_not_take():
    ops = pkg("ops")
    w, b = torch.ones(384, device="cuda"), torch.zeros(384, device="cuda")
    with pytest.raises(ValueError):
        ops.residual_layer_norm(torch.zeros(4, 384, device="cuda"), None, w, b)                # width
    w, b = torch.ones(512, device="cuda"), torch.zeros(512, device="cuda")
    with pytest.raises(ValueError):
        ops.residual_layer_norm(torch.zeros(4, 512, device="cuda"), torch.zeros(4, 512, device="cuda").bfloat16(), w, b)
    with pytest.raises(ValueError):
        ops.residual_layer_norm(torch.zeros(4, 512, device="cuda"), None, w, b, dropout_p=1.0)


@pytest.mark.parametrize("autocast", [False, True])
def test_sub_layers_train_through_the_fused_layer_norm(monkeypatch, autocast):
    """PositionwiseFeedForward and MultiheadAttention (dropout off, so both routes are deterministic): the fused route and
    torch's nn.Dropout / + / nn.LayerNorm route give the same outputs and gradients."""
    mod, att, lib = pkg("transformer.module"), pkg("transformer.attention"), pkg("_lib")
    torch.manual_seed(0)
    ffn = mod.PositionwiseFeedForward(512, 2048, dropout=0.0).cuda().train()
    mha = att.MultiheadAttention(512, 8, dropout=0.0).cuda().train()
    mha.return_attn = False
    x = _rand((8, 168, 512), 5)
    gy = _rand((8, 168, 512), 6)
    pad = (torch.arange(168, device="cuda")[None, :, None] < torch.tensor([168, 100, 7, 168, 1, 50, 160, 33], device="cuda")[:, None, None]).float()

    def run(fused):
        monkeypatch.setattr(mod, "USE_FUSED_LAYER_NORM", fused)
        res = {}
        for name, layer in (("ffn", ffn), ("mha", mha)):
            layer.zero_grad()
            xi = x.clone().requires_grad_(True)
            n0 = lib.launch_count()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                y = layer(xi, out_scale=pad) if name == "ffn" else layer(xi, xi, xi, out_scale=pad)[0]
            assert y.dtype == torch.float32
            y.backward(gy)
            res[name] = (lib.launch_count() - n0, y.detach(), xi.grad.clone(), {k: p.grad.clone() for k, p in layer.named_parameters()})
        return res

    a, b = run(True), run(False)
    for name in ("ffn", "mha"):
        # fp32 feed-forward block: fp32 rounding only.  The attention core works in bf16 on either route: a 1e-7 change of
        # its output gradient moves bf16 roundings inside the backward kernel (4e-3 of single elements)
        tol = 2e-2 if autocast else (1e-4 if name == "ffn" else 5e-3)
        assert a[name][0] >= b[name][0] + 3                       # forward, backward and its column sum
        scale = b[name][1].abs().max().item()
        assert (a[name][1] - b[name][1]).abs().max().item() <= tol * scale
        assert (a[name][2] - b[name][2]).abs().max().item() <= tol * b[name][2].abs().max().item()
        for k, gb in b[name][3].items():
            if k == "w_ks.bias":      # softmax ignores a shift common to all keys: the true gradient is 0, both routes hold rounding noise
                continue
            assert (a[name][3][k] - gb).abs().max().item() <= tol * max(gb.abs().max().item(), 1e-6), (name, k)
