"""CTC kernels vs the oracle, the reference-generated golden vectors and torch's
own F.ctc_loss on the same device (the library call the reference makes)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from conftest import load_golden
from helpers import pkg, make_ctc_inputs, to_np

pytestmark = pytest.mark.gpu

CTC = load_golden("ctc")
CASES = sorted({k.split("_")[0] for k in CTC.files})


@pytest.fixture(autouse=True, params=[(0, 0), (0, 1), (1, 0)],
                ids=["bidirectional_lattice", "one_warp_lattice", "fused_apply"])
def apply_mode(request):
    """Lattice kernel variants: bidirectional four-warp lattice + separate K3 pass (default), the
    one-warp lattice, and the one-warp lattice that applies the sparse update itself."""
    lib = pkg("_lib")
    lib.set_option("ctc_fuse_apply", request.param[0])
    lib.set_option("ctc_lattice_variant", request.param[1])
    yield request.param
    lib.set_option("ctc_fuse_apply", 0)
    lib.set_option("ctc_lattice_variant", 0)


def _ours(logits, targets, in_len, need_grad=True):
    ops = pkg("ops")
    lg = logits.clone().requires_grad_(need_grad)
    loss, nll = ops.ctc_loss(lg, in_len, targets, return_nll=True)
    grad = None
    if need_grad:
        loss.backward()
        grad = lg.grad
    return loss.detach(), nll, grad


def _torch_ref(logits, targets, in_len, dtype=torch.float32):
    """Exactly the reference's call sequence (transformer/loss.py:39-43) on this device."""
    lg = logits.to(dtype).clone().requires_grad_(True)
    V = lg.size(-1)
    tl = targets.ne(0).int().sum(1)
    lp = F.log_softmax(lg, dim=-1).transpose(0, 1)
    loss = F.ctc_loss(lp, targets, in_len, tl, blank=V - 1)
    loss.backward()
    nll = F.ctc_loss(lp.detach(), targets, in_len, tl, blank=V - 1, reduction="none")
    return loss.detach(), nll, lg.grad


@pytest.mark.parametrize("case", CASES)
def test_golden(case):
    logits = torch.as_tensor(CTC[case + "_logits"]).cuda()
    targets = torch.as_tensor(CTC[case + "_targets"]).cuda()
    in_len = torch.as_tensor(CTC[case + "_in_len"]).cuda()
    loss, nll, grad = _ours(logits, targets, in_len)
    ref_nll, ref_loss, ref_grad = CTC[case + "_nll"], CTC[case + "_loss"], CTC[case + "_grad"]
    fin = np.isfinite(ref_nll)
    np.testing.assert_array_equal(np.isfinite(to_np(nll)), fin)
    np.testing.assert_allclose(to_np(nll)[fin], ref_nll[fin], rtol=1e-5)
    if np.isfinite(ref_loss):
        np.testing.assert_allclose(float(loss), ref_loss, rtol=1e-5)
    else:
        assert np.isinf(float(loss)) and float(loss) > 0
    g = to_np(grad)
    gscale = np.nanmax(np.abs(ref_grad)) + 1e-30
    # the golden vector is the reference's own fp32 result; its distance from the fp64 oracle is
    # the reference's rounding noise, which no other fp32 implementation can be asked to reproduce
    _, _, o_grad = oracle.ctc_loss_and_grad(CTC[case + "_logits"], CTC[case + "_targets"], CTC[case + "_in_len"])
    for b in range(g.shape[0]):
        if fin[b]:
            ref_noise = np.abs(ref_grad[b] - o_grad[b]).max()
            # fp32 rtol 1e-5 of the gradient scale
            assert np.abs(g[b] - ref_grad[b]).max() <= 1e-5 * gscale + 2 * ref_noise
            assert np.abs(g[b] - o_grad[b]).max() <= 1e-5 * gscale + ref_noise
        else:
            np.testing.assert_array_equal(np.isnan(g[b]), np.isnan(ref_grad[b]))
        assert not g[b, int(CTC[case + "_in_len"][b]):].any()


def test_forward_only_matches_and_writes_no_grad():
    logits = torch.as_tensor(CTC["b_logits"]).cuda()
    targets = torch.as_tensor(CTC["b_targets"]).cuda()
    in_len = torch.as_tensor(CTC["b_in_len"]).cuda()
    with torch.no_grad():
        loss, nll, _ = _ours(logits, targets, in_len, need_grad=False)
    np.testing.assert_allclose(to_np(nll), CTC["b_nll"], rtol=1e-5)
    np.testing.assert_allclose(float(loss), CTC["b_loss"], rtol=1e-5)


def test_reference_entry_points():
    tl = pkg("transformer.loss")
    cl = pkg("ctcModel.loss")
    g = load_golden("qua")
    dev = "cuda"
    qua, ctc, ce = tl.cal_ctc_qua_ce_loss(torch.as_tensor(g["logits"]).to(dev), torch.as_tensor(g["in_len"]).to(dev),
                                          torch.as_tensor(g["_number"]).to(dev), torch.as_tensor(g["number"]).to(dev),
                                          torch.as_tensor(g["ce_logits"]).to(dev), torch.as_tensor(g["targets"]).to(dev),
                                          smoothing=0.1)
    np.testing.assert_allclose(float(qua), g["qua"], rtol=1e-6)
    np.testing.assert_allclose(float(ctc), g["ctc"], rtol=1e-5)
    np.testing.assert_allclose(float(ce), g["ce"], rtol=1e-5)
    loss = cl.cal_loss(torch.as_tensor(g["logits"]).to(dev), torch.as_tensor(g["in_len"]).to(dev),
                       torch.as_tensor(g["targets"]).to(dev))
    np.testing.assert_allclose(float(loss), g["ctc"], rtol=1e-5)


def test_incoming_gradient_scale():
    """loss * 3 must scale the fused gradient (device-side, no host sync)."""
    ops = pkg("ops")
    logits, targets, in_len = make_ctc_inputs(4, 30, 50, 6, seed=3)
    a = logits.clone().requires_grad_(True)
    ops.ctc_loss(a, in_len, targets).backward()
    b = logits.clone().requires_grad_(True)
    (3.0 * ops.ctc_loss(b, in_len, targets)).backward()
    torch.testing.assert_close(b.grad, 3.0 * a.grad, rtol=1e-6, atol=0)


@pytest.mark.parametrize("B,T,V,S", [(4, 50, 100, 10), (3, 33, 4233, 7), (6, 70, 31, 20), (2, 200, 64, 40),
                                     (2, 90, 17, 80), (3, 64, 4233, 1), (2, 40, 10, 100)])
def test_random_vs_oracle_fp64(B, T, V, S):
    logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=100 + T + S)
    loss, nll, grad = _ours(logits, targets, in_len)
    o_loss, o_nll, o_grad = oracle.ctc_loss_and_grad(to_np(logits), to_np(targets), to_np(in_len))
    r_loss, r_nll, r_grad = _torch_ref(logits, targets, in_len)
    fin = np.isfinite(o_nll)
    np.testing.assert_array_equal(np.isfinite(to_np(nll)), fin)
    # ours vs fp64 truth must be no worse than 1e-5, or than torch's own fp32 error (x2)
    err_ours = np.abs(to_np(nll)[fin] - o_nll[fin]) / np.abs(o_nll[fin])
    err_ref = np.abs(to_np(r_nll)[fin] - o_nll[fin]) / np.abs(o_nll[fin])
    assert (err_ours <= np.maximum(1e-5, 2 * err_ref)).all(), (err_ours, err_ref)
    if not fin.any():
        assert np.isnan(to_np(grad)[:, 0]).all()
        return
    gs = np.nanmax(np.abs(o_grad[fin]))
    ge_ours = np.abs(to_np(grad)[fin] - o_grad[fin]).max() / gs
    ge_ref = np.abs(to_np(r_grad)[fin] - o_grad[fin]).max() / gs
    # (T=90, S=80 has |log-likelihood| ~ 350 bits, where one fp32 ulp is already 3e-5: every fp32
    # log-domain lattice, torch's included, sits at ~1e-4 there and the variants differ by rounding path)
    assert ge_ours <= max(1e-5, 3 * ge_ref), (ge_ours, ge_ref)


@pytest.mark.parametrize("B,T,S", [(32, 200, 10), (64, 400, 20)])
def test_config2_sizes_vs_torch_on_device(B, T, S):
    """BASELINE config 2 shapes (V=4233): ours vs the library call the reference
    makes, on the same device, plus size-independent properties."""
    V = 4233
    logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=1236)
    loss, nll, grad = _ours(logits, targets, in_len)
    r_loss, r_nll, r_grad = _torch_ref(logits, targets, in_len)
    d_loss, d_nll, d_grad = _torch_ref(logits, targets, in_len, dtype=torch.float64)
    err_ours = ((nll.double() - d_nll).abs() / d_nll.abs()).max().item()
    err_ref = ((r_nll.double() - d_nll).abs() / d_nll.abs()).max().item()
    assert err_ours <= max(1e-5, 2 * err_ref), (err_ours, err_ref)
    gs = d_grad.abs().max().item()
    ge_ours = (grad.double() - d_grad).abs().max().item() / gs
    ge_ref = (r_grad.double() - d_grad).abs().max().item() / gs
    assert ge_ours <= max(1e-5, 2 * ge_ref), (ge_ours, ge_ref)
    # properties: rows of d nll/d logits sum to zero (softmax and occupancy both sum to 1);
    # frames beyond the input length are exactly zero
    row_sum = grad.double().sum(-1).abs().max().item()
    assert row_sum <= 1e-6 * max(1.0, gs * V)
    for b in (0, B // 2, B - 1):
        assert not grad[b, int(in_len[b]):].any()


def test_determinism():
    logits, targets, in_len = make_ctc_inputs(8, 120, 500, 15, seed=7)
    _, n1, g1 = _ours(logits, targets, in_len)
    _, n2, g2 = _ours(logits, targets, in_len)
    assert torch.equal(n1, n2) and torch.equal(g1, g2)


@pytest.mark.parametrize("chunks", [2, 3, 8])
def test_sliced_pipeline_is_bit_identical(chunks):
    """The batch-sliced multi-stream pipeline (option ctc_chunks) must not change a single bit:
    every utterance's lattice is independent and the mean-loss normaliser stays the full B."""
    lib = pkg("_lib")
    B, T, V, S = 11, 70, 257, 9        # 11 utterances: ragged slices for every chunk count
    logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=321)
    try:
        lib.set_option("ctc_chunks", 1)
        loss1, nll1, g1 = _ours(logits, targets, in_len)
        lib.set_option("ctc_chunks", chunks)
        loss2, nll2, g2 = _ours(logits, targets, in_len)
        # and again, back to back on the same streams (event reuse)
        loss3, nll3, g3 = _ours(logits, targets, in_len)
    finally:
        lib.set_option("ctc_chunks", 0)
    torch.cuda.synchronize()
    assert torch.equal(nll1, nll2) and torch.equal(g1, g2) and torch.equal(loss1, loss2)
    assert torch.equal(nll1, nll3) and torch.equal(g1, g3)


def test_every_length_around_block_and_midpoint_boundaries():
    """Input lengths 1..135 (one utterance each) cross every boundary of the bidirectional lattice:
    one warp owning all frames (len <= 32), a one-frame second half (len 33), partial last blocks,
    both halves with several blocks; target lengths 0..7 include the empty target and infeasible pairs."""
    T, V, S = 135, 23, 7
    B = T
    g = torch.Generator().manual_seed(99)
    logits = torch.randn(B, T, V, generator=g)
    tgt_len = (torch.arange(B) * 5) % (S + 1)
    targets = torch.randint(1, V - 1, (B, S), generator=g)
    targets[::3, 1] = targets[::3, 0]                     # repeats need an extra blank frame
    targets = targets * (torch.arange(S)[None, :] < tgt_len[:, None]).long()
    in_len = torch.arange(1, B + 1, dtype=torch.int32)
    loss, nll, grad = _ours(logits.cuda(), targets.cuda(), in_len.cuda())
    o_loss, o_nll, o_grad = oracle.ctc_loss_and_grad(logits.numpy(), targets.numpy(), in_len.numpy())
    fin = np.isfinite(o_nll)
    assert fin.sum() > 100 and (~fin).sum() >= 2
    np.testing.assert_array_equal(np.isfinite(to_np(nll)), fin)
    np.testing.assert_allclose(to_np(nll)[fin], o_nll[fin], rtol=1e-5, atol=1e-5)
    gn = to_np(grad)
    # per utterance: fp32 rtol 1e-5 of the gradient scale, or 3x the error torch's own fp32 lattice
    # makes on the same utterance (|log-likelihood| reaches 600 bits here, one fp32 ulp = 6e-5)
    _, _, r_grad = _torch_ref(logits.cuda(), targets.cuda(), in_len.cuda())
    r_grad = to_np(r_grad)
    for b in np.nonzero(fin)[0]:
        gs = np.abs(o_grad[b]).max()
        e_ours = np.abs(gn[b] - o_grad[b]).max() / gs
        e_ref = np.abs(r_grad[b] - o_grad[b]).max() / gs
        assert e_ours <= max(1e-5, 3 * e_ref), (b, e_ours, e_ref)
    for b in np.nonzero(~fin)[0]:
        np.testing.assert_array_equal(np.isnan(gn[b]), np.isnan(o_grad[b]))
    for b in range(B):
        assert not gn[b, int(in_len[b]):].any()


def test_begin_finish_with_work_in_between_is_bit_identical():
    """asr_ctc_begin_f32 / asr_ctc_finish_f32 with unrelated kernels queued in between must give
    exactly what the single call gives (two outstanding tickets, interleaved)."""
    import ctypes
    lib = pkg("_lib")
    L, p = lib.lib(), lib.ptr
    B, T, V, S = 9, 80, 131, 6
    outs = []
    try:
        lib.set_option("ctc_chunks", 3)
        data = []
        for seed in (5, 6):
            logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=seed)
            tgt_len = targets.ne(0).sum(1).to(torch.int32)
            wsb = L.asr_ctc_workspace_bytes(B, T, V, S)
            d = dict(logits=logits, targets=targets, in_len=in_len, tgt_len=tgt_len, wsb=wsb)
            for tag in ("one", "two"):
                d["nll_" + tag] = torch.empty(B, device="cuda")
                d["g_" + tag] = torch.empty_like(logits)
                d["ws_" + tag] = torch.empty(wsb // 4 + 1, device="cuda")
            data.append(d)

        def args(d, tag):
            return (p(d["logits"]), p(d["targets"]), p(d["in_len"]), p(d["tgt_len"]), B, T, V, S, V - 1,
                    p(d["nll_" + tag]), p(d["g_" + tag]), p(d["ws_" + tag]), d["wsb"], lib.stream_ptr())
        for d in data:
            lib.check(L.asr_ctc_fwd_bwd_f32(*args(d, "one")), "fwd_bwd")
        tickets = [ctypes.c_int(-1), ctypes.c_int(-1)]
        filler = torch.randn(1 << 20, device="cuda")
        for d, tk in zip(data, tickets):
            lib.check(L.asr_ctc_begin_f32(*args(d, "two"), ctypes.byref(tk)), "begin")
            filler = filler * 1.0001 + 1.0
        assert tickets[0].value != tickets[1].value
        for d, tk in zip(data, tickets):
            lib.check(L.asr_ctc_finish_f32(*args(d, "two"), tk.value), "finish")
        torch.cuda.synchronize()
        for d in data:
            assert torch.equal(d["nll_one"], d["nll_two"])
            assert torch.equal(d["g_one"], d["g_two"])
    finally:
        lib.set_option("ctc_chunks", 0)


@pytest.mark.parametrize("B,T,V,S", [(2, 600, 300, 255), (3, 3000, 64, 30), (2, 700, 40, 150), (1, 1, 5, 1), (4, 37, 9, 0)])
def test_extreme_shapes_vs_torch_on_device(B, T, V, S):
    """Maximum label count (511 lattice states: one-warp fallback), very long inputs (global checkpoints),
    12 states per lane, a single frame, and empty targets, against torch's own fp32 path and the fp64 oracle."""
    if S == 0:
        logits = torch.randn(B, T, V, generator=torch.Generator().manual_seed(1)).cuda()
        targets = torch.zeros(B, 1, dtype=torch.int64).cuda()
        in_len = torch.tensor([T, T - 3, 1, 20][:B], dtype=torch.int32).cuda()
    else:
        logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=7 + T)
    loss, nll, grad = _ours(logits, targets, in_len)
    o_loss, o_nll, o_grad = oracle.ctc_loss_and_grad(to_np(logits), to_np(targets), to_np(in_len))
    r_loss, r_nll, r_grad = _torch_ref(logits, targets, in_len)
    fin = np.isfinite(o_nll)
    np.testing.assert_array_equal(np.isfinite(to_np(nll)), fin)
    err_ours = np.abs(to_np(nll)[fin] - o_nll[fin]) / np.maximum(np.abs(o_nll[fin]), 1.0)
    err_ref = np.abs(to_np(r_nll)[fin] - o_nll[fin]) / np.maximum(np.abs(o_nll[fin]), 1.0)
    assert (err_ours <= np.maximum(1e-5, 2 * err_ref)).all(), (err_ours, err_ref)
    if fin.any():
        gs = np.nanmax(np.abs(o_grad[fin]))
        ge_ours = np.abs(to_np(grad)[fin] - o_grad[fin]).max() / gs
        ge_ref = np.abs(to_np(r_grad)[fin] - o_grad[fin]).max() / gs
        assert ge_ours <= max(1e-5, 3 * ge_ref), (ge_ours, ge_ref)
    for b in range(B):
        assert not to_np(grad)[b, int(in_len[b]):].any()


def test_bad_labels_give_nan_not_a_silently_wrong_gradient():
    """A label equal to the blank or outside [0, V) has no lattice state of its own (its gradient entry would collide
    with the blank's): the utterance reports NaN loss and NaN gradient rows, the other utterances are untouched."""
    ops = pkg("ops")
    B, T, V, S = 4, 40, 50, 6
    logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=17, full_len=True)
    good = logits.clone().requires_grad_(True)
    loss_g, nll_g = ops.ctc_loss(good, in_len, targets, return_nll=True)
    loss_g.backward()
    for bad_value in (V - 1, V + 3):
        t2 = targets.clone()
        t2[1, 2] = bad_value
        x = logits.clone().requires_grad_(True)
        loss, nll = ops.ctc_loss(x, in_len, t2, return_nll=True)
        loss.backward()
        assert torch.isnan(nll[1]) and torch.isnan(loss)
        assert torch.isnan(x.grad[1, : int(in_len[1])]).all()
        keep = [0, 2, 3]
        assert torch.equal(nll[keep], nll_g[keep]) and torch.equal(x.grad[keep], good.grad[keep])


def test_ticket_ring_fails_loudly():
    """At most 4 begins may be open per device: the 5th returns an error instead of aliasing an open ticket's events, and
    finish on a ticket that is not open is refused."""
    import ctypes
    lib = pkg("_lib")
    L = lib.lib()
    B, T, V, S = 3, 20, 30, 4
    logits, targets, in_len = make_ctc_inputs(B, T, V, S, seed=18)
    tgt_len = targets.ne(0).sum(1).to(torch.int32)
    nll, g = torch.empty(B, device="cuda"), torch.empty_like(logits)
    wsb = L.asr_ctc_workspace_bytes(B, T, V, S)
    wss = [torch.empty(wsb // 4 + 1, device="cuda") for _ in range(5)]
    args = lambda ws: (lib.ptr(logits), lib.ptr(targets), lib.ptr(in_len), lib.ptr(tgt_len), B, T, V, S, V - 1, lib.ptr(nll),  # noqa: E731
                       lib.ptr(g), lib.ptr(ws), wsb, lib.stream_ptr())
    tickets = []
    try:
        for i in range(4):
            tk = ctypes.c_int(-1)
            assert L.asr_ctc_begin_f32(*args(wss[i]), ctypes.byref(tk)) == 0
            tickets.append(tk.value)
        assert sorted(tickets) == [0, 1, 2, 3]
        tk = ctypes.c_int(-1)
        assert L.asr_ctc_begin_f32(*args(wss[4]), ctypes.byref(tk)) != 0
        assert "between begin and finish" in lib.last_error()
    finally:
        for i, t in enumerate(tickets):
            assert L.asr_ctc_finish_f32(*args(wss[i]), t) == 0
    assert L.asr_ctc_finish_f32(*args(wss[0]), tickets[0]) != 0 and "not open" in lib.last_error()
    torch.cuda.synchronize()
    tk = ctypes.c_int(-1)                        # and the ring is usable again
    assert L.asr_ctc_begin_f32(*args(wss[0]), ctypes.byref(tk)) == 0
    assert L.asr_ctc_finish_f32(*args(wss[0]), tk.value) == 0
    torch.cuda.synchronize()
