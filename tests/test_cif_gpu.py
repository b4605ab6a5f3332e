"""CIF kernels vs the oracle and the reference-generated golden vectors (B200)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden
from helpers import pkg, make_cif_inputs, to_np, bits

pytestmark = pytest.mark.gpu

CIF = load_golden("cif")
CASES = sorted({k.split("_")[0] for k in CIF.files})
# (variant, width): 1 = plain loads, 2 = one-warp TMA pipeline, 3 = warp-specialised TMA pipeline,
# 4 = schedule kernel + segment-parallel rows
VARIANTS = [(1, 32), (1, 64), (1, 128), (2, 32), (2, 64), (2, 128), (3, 32), (3, 64), (3, 128), (4, 0), (4, 32), (0, 0)]


@pytest.fixture(autouse=True)
def _reset_options():
    lib = pkg("_lib")
    yield
    for k in ("cif_fwd_variant", "cif_fwd_width", "cif_fwd_stages", "cif_fwd_rows"):
        lib.set_option(k, 0)


def _set(variant, width, stages=0, nw=0):
    lib = pkg("_lib")
    lib.set_option("cif_fwd_variant", variant)
    lib.set_option("cif_fwd_width", width)
    lib.set_option("cif_fwd_stages", stages)
    lib.set_option("cif_fwd_rows", nw)


def _run(hidden, alphas, thr, L, g_out=None):
    ops = pkg("ops")
    h = torch.as_tensor(hidden).cuda().requires_grad_(True)
    a = torch.as_tensor(alphas).cuda().requires_grad_(True)
    out, aux = ops.cif(h, a, thr, L=L, return_aux=True)
    res = {"out": to_np(out), "fire_t": to_np(aux["fire_t"]), "n_fired": to_np(aux["n_fired"]),
           "alpha_sum": to_np(aux["alpha_sum"])}
    if g_out is not None and L > 0:
        out.backward(torch.as_tensor(g_out).cuda())
        res["g_hidden"], res["g_alpha"] = to_np(h.grad), to_np(a.grad)
    return res


@pytest.mark.parametrize("variant,width", VARIANTS)
@pytest.mark.parametrize("case", CASES)
def test_golden_forward_bit_exact(case, variant, width):
    _set(variant, width)
    ref_out = CIF[case + "_out"]
    L = ref_out.shape[1]
    r = _run(CIF[case + "_hidden"], CIF[case + "_alphas"], float(CIF[case + "_thr"]), L)
    np.testing.assert_array_equal(r["n_fired"], CIF[case + "_n_fired"])          # integer work: exact
    if L > 0:
        np.testing.assert_array_equal(r["fire_t"], CIF[case + "_fire_t"][:, :L])  # fire positions: exact
    assert r["out"].shape == ref_out.shape
    np.testing.assert_array_equal(bits(r["out"]), bits(ref_out))                  # fp32, reference op order: exact bits


@pytest.mark.parametrize("case", CASES)
def test_golden_backward(case):
    ref_out = CIF[case + "_out"]
    L = ref_out.shape[1]
    if L == 0:
        pytest.skip("no fires -> no gradient path")
    r = _run(CIF[case + "_hidden"], CIF[case + "_alphas"], float(CIF[case + "_thr"]), L, CIF[case + "_g_out"])
    # g_hidden = cur*gpre + rem*G has no reduction: exact (up to the sign of zero)
    np.testing.assert_allclose(r["g_hidden"], CIF[case + "_g_hidden"], rtol=0, atol=0)
    # g_alpha holds H-long dot products (order differs from torch): fp32 rtol 1e-5 of the gradient scale
    ref = CIF[case + "_g_alpha"]
    scale = np.abs(ref).max() + 1e-30
    assert np.abs(r["g_alpha"] - ref).max() <= 1e-5 * scale * max(1.0, np.sqrt(ref_out.shape[2] / 8.0))


def test_default_L_and_module_api():
    cif_model = pkg("transformer.cif_model")
    hidden, alphas = CIF["b_hidden"], CIF["b_alphas"]
    y = cif_model.CIF_Model.cif(None, torch.as_tensor(hidden).cuda(), torch.as_tensor(alphas).cuda(), 0.95)
    np.testing.assert_array_equal(bits(to_np(y)), bits(CIF["b_out"]))


def test_overflow_raises_like_reference():
    ops = pkg("ops")
    with pytest.raises(RuntimeError):
        ops.cif(torch.as_tensor(CIF["a_hidden"]).cuda(), torch.as_tensor(CIF["a_alphas"]).cuda(), 0.95, L=3)


def test_cpu_tensor_is_rejected():
    ops = pkg("ops")
    with pytest.raises(RuntimeError):
        ops.cif(torch.zeros(1, 4, 8), torch.zeros(1, 4), 0.95)


@pytest.mark.parametrize("B,T,H,n", [(8, 21, 512, 14), (5, 167, 512, 20), (3, 301, 320, 40), (2, 97, 100, 9),
                                     (2, 64, 37, 9), (1, 700, 256, 60)])
@pytest.mark.parametrize("variant,width", [(1, 0), (2, 0), (2, 128), (3, 0), (3, 64), (3, 128)])
def test_random_vs_oracle(B, T, H, n, variant, width):
    if variant >= 2 and H % 4:
        pytest.skip("TMA path needs 16-byte rows")
    _set(variant, width)
    hidden, alphas = make_cif_inputs(B, T, H, n, seed=1234 + T)
    L = oracle.cif_oracle.cif_label_len(to_np(alphas))
    ref_out, ref_fire, ref_n = oracle.cif_forward(to_np(hidden), to_np(alphas), 0.95, L=L)
    g_out = torch.randn(B, L, H, generator=torch.Generator().manual_seed(5)).numpy()
    r = _run(to_np(hidden), to_np(alphas), 0.95, L, g_out)
    np.testing.assert_array_equal(r["n_fired"], ref_n)
    np.testing.assert_array_equal(r["fire_t"], ref_fire[:, :L])
    np.testing.assert_array_equal(bits(r["out"]), bits(ref_out))
    gh, ga = oracle.cif_backward(to_np(hidden), to_np(alphas), 0.95, g_out, dtype=np.float64)
    np.testing.assert_allclose(r["g_hidden"], gh, rtol=1e-5, atol=1e-6)
    scale = np.abs(ga).max()
    assert np.abs(r["g_alpha"] - ga).max() <= 1e-5 * scale * max(1.0, np.sqrt(H / 8.0))


def test_stage_counts_agree():
    hidden, alphas = make_cif_inputs(4, 500, 256, 50, seed=3)
    outs = []
    for stages in (1, 2, 3, 6, 12):
        _set(2, 64, stages)
        outs.append(_run(to_np(hidden), to_np(alphas), 0.95, 60)["out"])
    for stages, nw in ((1, 1), (2, 4), (3, 2), (8, 4), (12, 1)):
        _set(3, 32, stages, nw)
        outs.append(_run(to_np(hidden), to_np(alphas), 0.95, 60)["out"])
    for o in outs[1:]:
        np.testing.assert_array_equal(bits(o), bits(outs[0]))


def test_cfg4_long_utterance_stress():
    """BASELINE config 4: T=3000, H=512, B=64 (forward + backward).
    Size-independent checks: variant agreement (bit-exact), a 3-utterance slice
    against the oracle, and the adjoint identity <g_out, cif(dh)> = <g_hidden, dh>
    (cif is linear in hidden for a fixed fire schedule)."""
    ops = pkg("ops")
    B, T, H = 64, 3000, 512
    hidden, alphas = make_cif_inputs(B, T, H, 300, seed=1238)
    L = ops.cif_label_len(alphas)
    _set(2, 0)
    h = hidden.clone().requires_grad_(True)
    a = alphas.clone().requires_grad_(True)
    out, aux = ops.cif(h, a, 0.95, L=L, return_aux=True)
    g_out = torch.randn(out.shape, generator=torch.Generator().manual_seed(9)).cuda()
    out.backward(g_out)
    _set(1, 0)
    out_plain = ops.cif(hidden, alphas, 0.95, L=L)
    assert torch.equal(out.detach(), out_plain)
    _set(3, 0)
    out_ws, aux_ws = ops.cif(hidden, alphas, 0.95, L=L, return_aux=True)
    assert torch.equal(out.detach(), out_ws)
    assert torch.equal(aux["fire_t"], aux_ws["fire_t"]) and torch.equal(aux["n_fired"], aux_ws["n_fired"])
    # slice vs oracle
    sel = [0, 31, 63]
    ref_out, ref_fire, ref_n = oracle.cif_forward(to_np(hidden[sel]), to_np(alphas[sel]), 0.95, L=L)
    np.testing.assert_array_equal(to_np(aux["n_fired"])[sel], ref_n)
    np.testing.assert_array_equal(to_np(aux["fire_t"])[sel], ref_fire[:, :L])
    np.testing.assert_array_equal(bits(to_np(out.detach()[sel])), bits(ref_out))
    gh, ga = oracle.cif_backward(to_np(hidden[sel]), to_np(alphas[sel]), 0.95, to_np(g_out[sel]), dtype=np.float64)
    np.testing.assert_allclose(to_np(h.grad[sel]), gh, rtol=1e-5, atol=1e-6)
    assert np.abs(to_np(a.grad[sel]) - ga).max() <= 1e-4 * np.abs(ga).max()
    # adjoint identity over the whole batch
    dh = torch.randn(hidden.shape, generator=torch.Generator().manual_seed(11)).cuda()
    lhs = (ops.cif(dh, alphas, 0.95, L=L).double() * g_out.double()).sum()
    rhs = (h.grad.double() * dh.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * (abs(float(lhs)) + abs(float(rhs)) + 1.0) * 10
    # zero rows beyond the fire count
    n_fired = aux["n_fired"]
    for b in (0, 17, 63):
        assert not out[b, int(n_fired[b]):].any()


def test_alpha_sum_and_quantity_term():
    ops = pkg("ops")
    hidden, alphas = make_cif_inputs(6, 120, 64, 12, seed=21)
    num = torch.randint(5, 15, (6,)).float().cuda()
    out, aux = ops.cif(hidden, alphas, 0.95, target_num=num, return_aux=True)
    ref = to_np(alphas).astype(np.float64).sum(-1)
    np.testing.assert_allclose(to_np(aux["alpha_sum"]), ref, rtol=1e-5)
    np.testing.assert_allclose(to_np(aux["qua_term"]), (ref - to_np(num)) ** 2, rtol=1e-4, atol=1e-6)


def test_per_call_kernel_hint_matches_the_option_and_is_bit_identical():
    """asr_cif_fwd_hint_f32: the kernel choice as a per-call argument (bench.py queues the warp-specialised kernel between
    the CTC phases) instead of the process-wide option table; every choice produces the same bits."""
    lib = pkg("_lib")
    L_ = lib.lib()
    hidden, alphas = make_cif_inputs(6, 300, 256, 20, seed=11)
    B, T, H = hidden.shape
    Lout = 22
    ref = None
    for hint in (0, 1, 2, 3, 4):
        out = torch.empty(B, Lout, H, device="cuda")
        fire_t = torch.empty(B, Lout, dtype=torch.int32, device="cuda")
        n_fired = torch.empty(B, dtype=torch.int32, device="cuda")
        cur, rem = torch.empty(B, T, device="cuda"), torch.empty(B, T, device="cuda")
        sched = torch.empty(B, T, dtype=torch.int32, device="cuda")
        asum = torch.empty(B, device="cuda")
        lib.check(L_.asr_cif_fwd_hint_f32(lib.ptr(hidden), lib.ptr(alphas), 0.95, B, T, H, Lout, lib.ptr(out), lib.ptr(fire_t),
                                          lib.ptr(n_fired), lib.ptr(cur), lib.ptr(rem), lib.ptr(sched), lib.ptr(asum), None, None,
                                          hint, lib.stream_ptr()), "asr_cif_fwd_hint_f32")
        got = (out, fire_t, n_fired, cur, rem, sched)
        if ref is None:
            ref = got
        else:
            for a, b in zip(ref, got):
                assert torch.equal(a, b), hint
    assert lib.get_option("cif_fwd_variant") == 0            # the hint did not touch the process-wide option
