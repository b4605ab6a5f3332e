"""The general tcgen05 GEMMs (csrc/gemm2.cu: K-major / MN-major operands, fp32 "3xTF32" and bf16, split-K), the
training-mode linear layers built on them (forward, dX and dW of module.py:46-53 /
attention.py:40-45,59-60 without transposed copies) and the vocabulary projection fused with the CTC loss (8(f1):
cif_model.py:38 + loss.py:39-43) - against torch in fp64 / fp32 on the same device."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import pkg, make_ctc_inputs

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda().to(dtype)


def _operands(M, N, K, a_mn, b_mn, seed, dtype=torch.float32):
    a = _rand((K, M) if a_mn else (M, K), seed, dtype=dtype)
    b = _rand((K, N) if b_mn else (N, K), seed + 1, dtype=dtype)
    A = (a.t() if a_mn else a).double()
    B = (b if b_mn else b.t()).double()
    return a, b, A @ B


LAYOUTS = [(False, False), (False, True), (True, False), (True, True)]


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (1344, 512, 512), (300, 2048, 512), (257, 132, 100), (1344, 512, 2048), (64, 4236, 512)])
def test_gemm_f32_layouts_match_fp64(M, N, K, a_mn, b_mn):
    """fp32-level accuracy (three TF32 products): 2e-5 of the result scale at K <= 2048; cuBLAS-with-TF32 would be 1e-3."""
    ops = pkg("ops")
    if (a_mn and M % 4) or (b_mn and N % 4) or (not (a_mn and b_mn) and K % 4):
        pytest.skip("row stride not a multiple of 16 bytes for this layout")
    a, b, ref = _operands(M, N, K, a_mn, b_mn, seed=M + N + K)
    for split in (False, True):
        got = ops.gemm_f32(a, b, a_mn_major=a_mn, b_mn_major=b_mn, split_k=split)
        err = (got.double() - ref).abs().max().item() / ref.abs().max().item()
        assert err <= 2e-5, (err, split)


def test_gemm_f32_forced_split_k_and_bias_and_padded_output():
    ops, lib = pkg("ops"), pkg("_lib")
    M, N, K = 200, 4233, 512          # the vocabulary projection: odd N inside rows padded to a multiple of 4
    a, b, ref = _operands(M, N, K, False, False, seed=3)
    bias = _rand((N,), 9)
    out = torch.full((M, 4236), float("nan"), device="cuda")
    ops.gemm_f32(a, b, bias=bias, out=out, split_k=False)
    ref_b = ref + bias.double()
    assert (out[:, :N].double() - ref_b).abs().max().item() <= 2e-5 * ref_b.abs().max().item()
    assert torch.isfinite(out[:, N:]).all()              # zeros (+ nothing) in the padding, never NaN garbage
    try:
        lib.set_option("gemm_split_k", 4)
        got = ops.gemm_f32(a, b, bias=bias)
    finally:
        lib.set_option("gemm_split_k", 0)
    assert (got.double() - ref_b).abs().max().item() <= 2e-5 * ref_b.abs().max().item()
    # long contraction, few output tiles: dW = g^T h with the gradient rows padded (a strided view), split along K
    rows, V, H = 5000, 4233, 512
    g = torch.zeros(rows, 4236, device="cuda")
    g[:, :V] = _rand((rows, V), 4, 1e-3)
    h = _rand((rows, H), 5)
    got = ops.gemm_f32(g[:, :V], h, a_mn_major=True, b_mn_major=True)
    ref = g[:, :V].double().t() @ h.double()
    assert got.shape == (V, H)
    assert (got.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    got = ops.gemm_f32(g[:, :V], _rand((V, H), 6), b_mn_major=True, split_k=False)      # d h = g W, K = 4233 with a zero-filled tail
    assert got.shape == (rows, H)


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("M,N,K,splits", [(1344, 512, 512, 0), (896, 512, 2048, 8), (264, 132, 1000, 3), (512, 512, 15032, 5), (100, 4236, 512, 2)])
def test_split_k_inside_a_cluster_is_bitwise_the_workspace_reduction(M, N, K, splits, a_mn, b_mn):
    """Split-K partial tiles meet in distributed shared memory (the z CTAs of a tile are one cluster) - the same sums in
    the same order as the older workspace + reduce-kernel flavour (gemm_split_mode = 1), with one launch."""
    ops, lib = pkg("ops"), pkg("_lib")
    if (a_mn and M % 8) or (b_mn and N % 8) or (not (a_mn and b_mn) and K % 8):
        pytest.skip("row stride not a multiple of 16 bytes for this layout")
    bias = _rand((N,), 11)
    a, b, ref = _operands(M, N, K, a_mn, b_mn, seed=M + N + K)
    a16, b16 = a.bfloat16(), b.bfloat16()
    ref_scale = ref.abs().max().item()
    out = {}
    try:
        lib.set_option("gemm_split_k", splits)
        for mode in (0, 1):
            lib.set_option("gemm_split_mode", mode)
            n0 = lib.launch_count()
            f32 = ops.gemm_f32(a, b, a_mn_major=a_mn, b_mn_major=b_mn, bias=bias)
            used = lib.launch_count() - n0
            padded = torch.full((M, N + 4), float("nan"), device="cuda")
            ops.gemm_f32(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out=padded)
            h = ops.gemm_bf16(a16, b16, a_mn_major=a_mn, b_mn_major=b_mn, bias=bias, relu=True)
            hf = ops.gemm_bf16(a16, b16, a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32)
            out[mode] = (f32, padded[:, :N].clone(), h, hf, used)
            assert torch.isnan(padded[:, N:]).all() or N % 4      # columns past the next multiple of 4 are never written
    finally:
        lib.set_option("gemm_split_k", 0)
        lib.set_option("gemm_split_mode", 0)
    for got, want in zip(out[0][:4], out[1][:4]):
        if splits:
            assert torch.equal(got, want)
        else:      # left to themselves the two flavours may cut K differently (their costs differ): the same sums, another order
            assert (got.double() - want.double()).abs().max().item() <= (2e-5 if got is out[0][0] or got is out[0][1] else 1e-2) * ref_scale
    assert out[0][4] == 1 and out[1][4] in (1, 2)                 # one launch; the workspace flavour adds its reduce kernel when it splits
    ref_b = ref + bias.double()
    assert (out[0][0].double() - ref_b).abs().max().item() <= 2e-5 * ref_b.abs().max().item()


def test_staged_epilogue_is_bitwise_the_direct_one():
    """gemm_stage_out = 2: an unsplit tile leaves through shared memory with row-contiguous stores - the same values."""
    ops, lib = pkg("ops"), pkg("_lib")
    for M, N, K in ((1344, 2048, 512), (300, 132, 96), (200, 4236, 512)):
        a, b, _ = _operands(M, N, K, False, False, seed=M + K)
        bias = _rand((N,), 12)
        got = {}
        try:
            lib.set_option("gemm_split_k", 1)
            for stage in (0, 2):
                lib.set_option("gemm_stage_out", stage)
                padded = torch.full((M, N + 4), float("nan"), device="cuda")
                ops.gemm_f32(a, b, bias=bias, out=padded)
                got[stage] = (padded, ops.gemm_bf16(a.bfloat16(), b.bfloat16(), bias=bias, relu=True),
                              ops.gemm_bf16(a.bfloat16(), b.bfloat16(), out_dtype=torch.float32))
        finally:
            lib.set_option("gemm_split_k", 0)
            lib.set_option("gemm_stage_out", 0)
        assert torch.equal(got[0][0][:, :N], got[2][0][:, :N]) and torch.isnan(got[2][0][:, N:]).all()
        assert torch.equal(got[0][1], got[2][1]) and torch.equal(got[0][2], got[2][2])


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1350, 512, 512), (15030, 2048, 512), (264, 136, 200), (512, 512, 15030)])
def test_gemm_bf16_layouts(M, N, K, a_mn, b_mn):
    ops = pkg("ops")
    if (a_mn and M % 8) or (b_mn and N % 8) or (not (a_mn and b_mn) and K % 8):
        pytest.skip("row stride not a multiple of 16 bytes for this layout")
    a, b, ref = _operands(M, N, K, a_mn, b_mn, seed=M + N + K, dtype=torch.bfloat16)
    scale = ref.abs().max().item()
    got = ops.gemm_bf16(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32)
    assert (got.double() - ref).abs().max().item() <= 5e-4 * scale      # exact bf16 products, fp32 accumulation (K up to 15030)
    bias = _rand((N,), 7)
    got16 = ops.gemm_bf16(a, b, a_mn_major=a_mn, b_mn_major=b_mn, bias=bias, relu=True)
    ref16 = torch.relu(ref + bias.double())
    assert got16.dtype == torch.bfloat16
    assert (got16.double() - ref16).abs().max().item() <= 1e-2 * scale


@pytest.mark.parametrize("rows,K,N,bias", [(1344, 512, 2048, True), (1344, 2048, 512, True), (896, 1024, 512, False), (300, 512, 4, True)])
def test_linear_f32_autograd_matches_torch(rows, K, N, bias):
    ops = pkg("ops")
    x = _rand((4, rows // 4, K), 1).requires_grad_(True)
    w = _rand((N, K), 2, K ** -0.5).requires_grad_(True)
    b = _rand((N,), 3).requires_grad_(True) if bias else None
    gy = _rand((4, rows // 4, N), 4)
    y = ops.linear_f32_autograd(x, w, b)
    y.backward(gy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yd = F.linear(xd, wd, bd)
    yd.backward(gy.double())
    rel = lambda got, want: (got.double() - want).abs().max().item() / want.abs().max().item()  # noqa: E731
    assert rel(y, yd) <= 2e-5 and rel(x.grad, xd.grad) <= 2e-5 and rel(w.grad, wd.grad) <= 2e-5
    if bias:
        assert rel(b.grad, bd.grad) <= 1e-5


def test_model_linear_layers_train_through_the_repo_gemms(monkeypatch):
    """PositionwiseFeedForward in training mode: forward and all parameter gradients through csrc/gemm2.cu - fp32, and bf16
    under autocast - against torch's own F.linear path (the switches off).  The bf16 route is held to the error torch's
    own bf16 route makes against the fp32 result: two bf16 implementations differ wherever a pre-activation rounds to
    the other side of zero (the ReLU mask of that element flips), so they are compared with the truth, not each other."""
    mod = pkg("transformer.module")
    lib = pkg("_lib")
    torch.manual_seed(0)
    ffn = mod.PositionwiseFeedForward(512, 2048, dropout=0.0).cuda().train()
    x = _rand((8, 168, 512), 5)
    gy = _rand((8, 168, 512), 6)

    def run(autocast, ours):
        monkeypatch.setattr(mod, "USE_TENSOR_CORE_FP32", ours)
        monkeypatch.setattr(mod, "USE_TENSOR_CORE_BF16", ours)
        monkeypatch.setattr(mod, "USE_FUSED_LAYER_NORM", ours)
        ffn.zero_grad()
        xi = x.clone().requires_grad_(True)
        n0 = lib.launch_count()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = ffn(xi)
        y.float().backward(gy)
        used = lib.launch_count() - n0
        assert (used >= 6) if ours else (used == 0)           # 2 layers x (forward, dX, dW) - or torch only
        out = {"y": y.float().detach(), "x.grad": xi.grad.clone()}
        out.update({k: p.grad.clone() for k, p in ffn.named_parameters()})
        assert all(v.dtype == torch.float32 for v in out.values())
        return out
    rel = lambda a, b: (a.double() - b.double()).abs().max().item() / (b.double().abs().max().item() + 1e-12)  # noqa: E731
    truth = run(False, False)                                  # torch fp32 (cuBLAS sgemm)
    ours32 = run(False, True)
    for k in truth:
        assert rel(ours32[k], truth[k]) <= 3e-5, (k, rel(ours32[k], truth[k]))
    torch16, ours16 = run(True, False), run(True, True)
    for k in truth:
        e_ours, e_torch = rel(ours16[k], truth[k]), rel(torch16[k], truth[k])
        assert e_ours <= max(2e-2, 1.5 * e_torch), (k, e_ours, e_torch)


def test_ctc_strided_rows_in_place_matches_the_contiguous_call():
    """asr_ctc_fwd_bwd_ld_f32 on rows padded to a multiple of 4 floats, gradient written over the logits, against the
    contiguous out-of-place call.  Not bit-identical by construction: with V odd the contiguous rows start at four
    different 16-byte phases, so the element-to-thread assignment of the row sums differs from the aligned padded rows
    (last-bit differences of the log-sum-exp)."""
    lib, ops = pkg("_lib"), pkg("ops")
    L = lib.lib()
    B, T, V, S = 5, 60, 4233, 7
    logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=23)
    x = logits.clone().requires_grad_(True)
    loss, nll_ref = ops.ctc_loss(x, in_len, targets, return_nll=True)
    loss.backward()
    ld = 4236
    buf = torch.full((B * T, ld), 7.0, device="cuda")
    buf[:, :V] = logits.reshape(B * T, V)
    tgt_len = targets.ne(0).sum(1).to(torch.int32)
    nll = torch.empty(B, device="cuda")
    wsb = L.asr_ctc_workspace_bytes(B, T, V, S)
    ws = torch.empty(wsb // 4 + 1, device="cuda")
    lib.check(L.asr_ctc_fwd_bwd_ld_f32(lib.ptr(buf), lib.ptr(targets), lib.ptr(in_len), lib.ptr(tgt_len), B, T, V, ld, S, V - 1,
                                       lib.ptr(nll), lib.ptr(buf), lib.ptr(ws), wsb, lib.stream_ptr()), "ld")
    torch.testing.assert_close(nll, nll_ref, rtol=2e-6, atol=0)
    got = buf[:, :V].reshape(B, T, V)
    assert (got - x.grad).abs().max().item() <= 2e-6 * x.grad.abs().max().item()
    assert (buf[:, V:] == 7.0).all()                          # the padding columns are never touched
    # and the aligned layout out of place (separate gradient buffer) is bit-identical to the in-place run
    buf2 = torch.full((B * T, ld), 7.0, device="cuda")
    buf2[:, :V] = logits.reshape(B * T, V)
    g2 = torch.zeros(B * T, ld, device="cuda")
    nll2 = torch.empty(B, device="cuda")
    lib.check(L.asr_ctc_fwd_bwd_ld_f32(lib.ptr(buf2), lib.ptr(targets), lib.ptr(in_len), lib.ptr(tgt_len), B, T, V, ld, S, V - 1,
                                       lib.ptr(nll2), lib.ptr(g2), lib.ptr(ws), wsb, lib.stream_ptr()), "ld")
    assert torch.equal(nll2, nll) and torch.equal(g2[:, :V], buf[:, :V])


@pytest.mark.parametrize("B,T,S,V,K", [(4, 50, 6, 300, 64), (6, 167, 14, 4233, 512), (32, 200, 10, 4233, 512)])
def test_ctc_fc_loss_matches_projection_plus_ctc(B, T, S, V, K):
    """8(f1): loss, d hidden and d weight of ctc(h W^T) against F.linear + F.log_softmax + F.ctc_loss in fp64 on the device,
    and against torch's own fp32 path (the bar: 1e-5 of the scale, or twice torch-fp32's own error)."""
    ops = pkg("ops")
    _, targets, in_len = make_ctc_inputs(B, T, V, S, seed=31)
    h = _rand((B, T, K), 1)
    w = _rand((V, K), 2, K ** -0.5)
    tgt_len = targets.ne(0).sum(1)

    def torch_path(dtype):
        hh, ww = h.detach().to(dtype).requires_grad_(True), w.detach().to(dtype).requires_grad_(True)
        lp = F.log_softmax(F.linear(hh, ww), dim=-1).transpose(0, 1)
        loss = F.ctc_loss(lp, targets, in_len.long(), tgt_len, blank=V - 1)
        loss.backward()
        return loss.detach(), hh.grad, ww.grad
    l64, gh64, gw64 = torch_path(torch.float64)
    l32, gh32, gw32 = torch_path(torch.float32)
    hh, ww = h.clone().requires_grad_(True), w.clone().requires_grad_(True)
    loss, nll = ops.ctc_fc_loss(hh, ww, in_len, targets, return_nll=True)
    (2.0 * loss).backward()                                   # incoming gradient scale is applied on the device
    rel = lambda a, b: (a.double() - b).abs().max().item() / b.abs().max().item()  # noqa: E731
    assert abs(float(loss) - float(l64)) <= max(1e-5, 2 * abs(float(l32) - float(l64)) / abs(float(l64))) * abs(float(l64))
    assert rel(hh.grad / 2, gh64) <= max(2e-5, 2 * rel(gh32, gh64)), (rel(hh.grad / 2, gh64), rel(gh32, gh64))
    assert rel(ww.grad / 2, gw64) <= max(2e-5, 2 * rel(gw32, gw64)), (rel(ww.grad / 2, gw64), rel(gw32, gw64))
    # and the un-fused route of this package (projection by torch, ops.ctc_loss) agrees on the loss
    plain = ops.ctc_loss(F.linear(h, w), in_len, targets)
    assert abs(float(plain) - float(loss)) <= 2e-5 * abs(float(plain))


def test_cif_model_with_fused_ctc_fc_matches_the_plain_route():
    from test_model_shell import _build, _golden_state, G
    tl = pkg("transformer.loss")
    res = []
    for fused in (False, True):
        model = _build()
        model.load_state_dict(_golden_state(), strict=False)
        model = model.cuda().train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if hasattr(m, "attn_dropout_p"):
                m.attn_dropout_p = 0.0
        model.fused_ctc_fc = fused
        feats, lens, targets = (torch.as_tensor(G[k]).cuda() for k in ("feats", "lens", "targets"))
        torch.manual_seed(int(G["rand_seed"]))
        out = model(feats, lens, targets)
        assert isinstance(out[0], pkg("ops").ProjectedLogits) == fused
        qua, ctc, ce = tl.cal_ctc_qua_ce_loss(out[0], out[1], out[2], out[3], out[4], targets, smoothing=0.1)
        (0.001 * qua + ctc + ce).backward()
        res.append((float(ctc), model.ctc_fc.weight.grad.clone(), model.encoder.linear_in.weight.grad.clone()))
    assert abs(res[0][0] - res[1][0]) <= 2e-5 * abs(res[0][0])
    for a, b in zip(res[0][1:], res[1][1:]):
        assert (a - b).abs().max().item() <= 1e-4 * a.abs().max().item()


@pytest.mark.parametrize("M,N,dtype", [(1344, 512, torch.float32), (15030, 2048, torch.bfloat16), (7, 4233, torch.float32), (300, 33, torch.bfloat16),
                                       (15030, 512, torch.bfloat16), (15030, 512, torch.float32), (2049, 70, torch.bfloat16), (513, 4236, torch.bfloat16),
                                       (100000, 64, torch.float32)])
def test_colsum_is_the_bias_gradient(M, N, dtype):
    ops = pkg("ops")
    x = _rand((M, N), M + N, dtype=dtype)
    got = ops.colsum(x)
    ref = x.double().sum(0)
    assert got.dtype == torch.float32
    assert (got.double() - ref).abs().max().item() <= 2e-6 * x.double().abs().sum(0).max().item()
    assert torch.equal(got, ops.colsum(x))                      # fixed summation order
    view = _rand((M, N + 8), 3, dtype=dtype)[:, :N]            # strided rows
    assert (ops.colsum(view).double() - view.double().sum(0)).abs().max().item() <= 2e-6 * view.double().abs().sum(0).max().item()


def test_gemm_f32_ragged_rows_skip_padding_without_changing_valid_results():
    """asr_gemm_f32_ragged: row tiles / K steps that lie entirely in the padding of an utterance are skipped; valid rows
    are bit-identical to the plain call, dead output tiles are zeros (or left alone), dead K steps contribute nothing."""
    ops = pkg("ops")
    B, T, K, N = 6, 200, 64, 300                     # T not a multiple of the 128-row tiles / 32-row K steps: tiles straddle utterances
    lens = torch.tensor([200, 37, 0, 129, 64, 1], dtype=torch.int32, device="cuda")
    valid = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).reshape(-1)
    x = _rand((B * T, K), 1)
    w = _rand((N, K), 2)
    full = ops.gemm_f32(x, w, split_k=False)
    out = torch.full((B * T, N), 5.0, device="cuda")
    ops.gemm_f32(x, w, out=out, split_k=False, row_len=lens, group_rows=T, skip_dead_output=True)
    assert torch.equal(out[valid], full[valid])
    assert (out[400:512] == 5.0).all()               # the row tile [384, 512) lies entirely in padding: never written
    z = ops.gemm_f32(x, w, split_k=False, row_len=lens, group_rows=T)
    assert torch.equal(z[valid], full[valid]) and (z[400:512] == 0).all()
    # contraction over the rows: zero the padded rows (what the CTC row pass guarantees), then skipping changes nothing
    g = _rand((B * T, N), 3) * valid[:, None]
    h = _rand((B * T, K), 4)
    ref = ops.gemm_f32(g, h, a_mn_major=True, b_mn_major=True, split_k=False)
    for split in (False, True):
        got = ops.gemm_f32(g, h, a_mn_major=True, b_mn_major=True, split_k=split, row_len=lens, group_rows=T)
        # split along K: the same products added in another order (up to eight partial sums), well inside the 2e-5 contract
        assert (got - ref).abs().max().item() <= 4e-6 * ref.abs().max().item()
    assert torch.equal(ops.gemm_f32(g, h, a_mn_major=True, b_mn_major=True, split_k=False, row_len=lens, group_rows=T), ref)


@pytest.mark.parametrize("rows,N", [(1344, 2048), (7, 24), (33, 40)])
def test_relu_backward_of_the_bf16_linear_is_torchs(rows, N):
    """dX / dW of relu(x W^T + b) under bf16: the gradient is masked by the saved output (y > 0) in one kernel - the same
    bits as torch's gy * (y > 0), zeros, negatives and NaN activations included."""
    ops, lib = pkg("ops"), pkg("_lib")
    y = _rand((rows, N), 1, dtype=torch.bfloat16)
    y.view(-1)[::7] = 0.0
    y.view(-1)[3::11] = float("nan")
    gy = _rand((rows, N), 2, dtype=torch.bfloat16)
    out = torch.empty_like(gy)
    lib.check(lib.lib().asr_relu_bwd_bf16(lib.ptr(gy), lib.ptr(y), lib.ptr(out), gy.numel(), lib.stream_ptr()), "relu_bwd")
    assert torch.equal(out, gy * (y > 0).to(gy.dtype))
    x = _rand((rows, 64), 3, dtype=torch.bfloat16).requires_grad_(True)
    w = _rand((N, 64), 4, 0.2).requires_grad_(True)
    b = _rand((N,), 5).requires_grad_(True)
    g = _rand((rows, N), 6, dtype=torch.bfloat16)
    ops.linear_bf16_autograd(x, w, b, relu=True).backward(g)
    xd, wd, bd = x.detach().double().requires_grad_(True), w.detach().bfloat16().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    F.relu(F.linear(xd, wd, bd)).backward(g.double())
    rel = lambda got, want: (got.double() - want).abs().max().item() / want.abs().max().item()  # noqa: E731
    assert rel(x.grad, xd.grad) <= 2e-2 and rel(w.grad, wd.grad) <= 2e-2 and rel(b.grad, bd.grad) <= 2e-2


@pytest.mark.parametrize("relu", [False, True])
def test_bf16_linear_with_an_odd_output_width(relu):
    """The 4233-class vocabulary projection under bf16 autocast (decoder.py:58 `tgt_word_prj`): forward through the
    shared-memory epilogue (rows of 4233 bf16 cannot take 16-byte stores), backward on the gradient padded to 4240 columns."""
    ops, lib = pkg("ops"), pkg("_lib")
    rows, K, N = 1350, 512, 4233
    x = _rand((rows, K), 1, dtype=torch.bfloat16).requires_grad_(True)
    w = _rand((N, K), 2, K ** -0.5).requires_grad_(True)
    b = _rand((N,), 3).requires_grad_(True)
    g = _rand((rows, N), 4)                      # an fp32 gradient, as the loss hands it over
    assert ops.linear_bf16_ok(x, w)
    n0 = lib.launch_count()
    y = ops.linear_bf16_autograd(x, w, b, relu=relu)
    y.backward(g.to(y.dtype))
    assert lib.launch_count() - n0 >= 4 and y.shape == (rows, N) and y.dtype == torch.bfloat16
    xd, wd, bd = x.detach().double().requires_grad_(True), w.detach().bfloat16().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    yd = F.linear(xd, wd, bd)
    if relu:      # the mask of the bf16 output: a pre-activation within a bf16 ulp of zero may round to the other side
        yd = yd * (y.detach() > 0).double()
    yd.backward(g.to(y.dtype).double())
    rel = lambda got, want: (got.double() - want).abs().max().item() / want.abs().max().item()  # noqa: E731
    assert rel(y, yd.detach()) <= 1e-2
    assert rel(x.grad, xd.grad) <= 2e-2 and rel(w.grad, wd.grad) <= 2e-2 and rel(b.grad, bd.grad) <= 2e-2


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(15030, 512, 512), (40000, 128, 64), (1350, 2048, 1024), (264, 136, 200), (128, 128, 64)])
def test_persistent_bf16_gemm_is_bitwise_the_one_tile_kernel(M, N, K, a_mn, b_mn, variant):
    """gemm_persistent = 2: one CTA walks several output tiles with the accumulator double-buffered in tensor memory - the
    same products in the same order as one CTA per tile (gemm_persistent = 1)."""
    ops, lib = pkg("ops"), pkg("_lib")
    if (a_mn and M % 8) or (b_mn and N % 8) or (not (a_mn and b_mn) and K % 8):
        pytest.skip("row stride not a multiple of 16 bytes for this layout")
    a, b, ref = _operands(M, N, K, a_mn, b_mn, seed=M + N + K, dtype=torch.bfloat16)
    bias = _rand((N,), 7)
    got = {}
    try:
        lib.set_option("gemm_split_k", 1)
        lib.set_option("gemm_variant", variant)    # 128- or 256-wide tiles, the same on both routes
        for mode in (1, 2):
            lib.set_option("gemm_persistent", mode)
            got[mode] = (ops.gemm_bf16(a, b, a_mn_major=a_mn, b_mn_major=b_mn, bias=bias, relu=True),
                         ops.gemm_bf16(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32),
                         ops.gemm_bf16(a, b, a_mn_major=a_mn, b_mn_major=b_mn))
    finally:
        lib.set_option("gemm_split_k", 0)
        lib.set_option("gemm_variant", 0)
        lib.set_option("gemm_persistent", 0)
    for x, y in zip(got[1], got[2]):
        assert torch.equal(x, y)
    assert (got[2][1].double() - ref).abs().max().item() <= 5e-4 * ref.abs().max().item()
