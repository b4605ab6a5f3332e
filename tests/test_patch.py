"""`patch.install()` - the "train.py unchanged" route of INTEGRATION.md: the drop-ins are installed into the
reference's own imported modules, a model is built by the REFERENCE's create_model and trained the
way transformer/solver.py:141-157 does.

The reference modules come from /root/reference/src (build container) or the staged copy baseline/_ref/src (GPU
box; oracle/build_ref.py).  Each test runs in its own interpreter: install() rebinds names inside sys.modules."""
import os
import subprocess
import sys
import textwrap

import pytest

from helpers import ROOT

sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present (run oracle/build_ref.py)")

PRELUDE = """
import argparse, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, %r)
from oracle import ref_loader
ns = ref_loader.load()
PKG = "end-to-end_asr_pytorch_b200"
patch = importlib.import_module(PKG + ".patch")
ours_cif = importlib.import_module(PKG + ".transformer.cif_model")
ours_loss = importlib.import_module(PKG + ".transformer.loss")
ours_closs = importlib.import_module(PKG + ".ctcModel.loss")
ours_att = importlib.import_module(PKG + ".transformer.attention")
ours_catt = importlib.import_module(PKG + ".ctcModel.attention")
def small_args(**kw):
    a = dict(d_input=80, LFR_m=4, n_conv_layers=3, d_model=64, n_layers_enc=1, n_head=2, d_inner=128, dropout=0.1,
             d_assigner_hidden=64, w_context=3, n_assigner_layers=3, sos_id=2, eos_id=3, vocab_size=100, n_layers_dec=1,
             spec_aug_cfg=None)
    a.update(kw)
    return argparse.Namespace(**a)
""" % ROOT


def _run(body, timeout=600):
    code = PRELUDE + textwrap.dedent(body)
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, "child failed:\nSTDOUT:\n%s\nSTDERR:\n%s" % (r.stdout[-3000:], r.stderr[-6000:])
    return r.stdout


@needs_ref
def test_install_rebinds_every_hot_path_symbol_and_uninstall_restores():
    _run("""
    ref_cif_fn = ns.cif_model.CIF_Model.cif
    ref_fwd = ns.cif_model.CIF_Model.forward
    ref_mha, ref_cmha = ns.attention.MultiheadAttention, ns.cattention.MultiHeadAttention
    ref_ctc_ce, ref_qua, ref_cal = ns.loss.cal_ctc_ce_loss, ns.loss.cal_ctc_qua_ce_loss, ns.closs.cal_loss
    import ctcModel.solver as csolver
    before = ns.cif_model.CIF_Model.create_model(small_args())
    keys_before = {k: tuple(v.shape) for k, v in before.state_dict().items()}

    done = patch.install()
    # CIF: only the kernel entry is replaced; forward stays the reference's (it calls self.cif)
    assert ns.cif_model.CIF_Model.cif is ours_cif.CIF_Model.cif
    assert ns.cif_model.CIF_Model.forward is ref_fwd
    # the losses, in their home modules and wherever `from transformer.loss import ...` already copied them
    assert ns.loss.cal_ctc_ce_loss is ours_loss.cal_ctc_ce_loss and ns.loss.cal_ctc_qua_ce_loss is ours_loss.cal_ctc_qua_ce_loss
    assert ns.solver.cal_ctc_ce_loss is ours_loss.cal_ctc_ce_loss and ns.solver.cal_ctc_qua_ce_loss is ours_loss.cal_ctc_qua_ce_loss
    assert ns.closs.cal_loss is ours_closs.cal_loss and csolver.cal_loss is ours_closs.cal_loss
    # attention: the class objects the encoder / decoder modules instantiate
    assert ns.attention.MultiheadAttention is ours_att.MultiheadAttention
    assert ns.encoder.MultiheadAttention is ours_att.MultiheadAttention and ns.decoder.MultiheadAttention is ours_att.MultiheadAttention
    assert ns.cattention.MultiHeadAttention is ours_catt.MultiHeadAttention
    import ctcModel.encoder as cenc                      # importable thanks to the get_non_pad_mask shim
    assert cenc.MultiHeadAttention is ours_catt.MultiHeadAttention
    for name in ("CIF_Model.cif", "cal_ctc_ce_loss", "cal_ctc_qua_ce_loss", "cal_loss", "MultiheadAttention", "MultiHeadAttention"):
        assert done.get(name, 0) >= 1, (name, done)
    # the module names train.py:139-157 imports now resolve
    from transformer.Transformer import Transformer, CTC_Transformer, Conv_CTC_Transformer
    from transformer.CIF_Model import CIF_Model
    assert CIF_Model is ns.cif_model.CIF_Model and Conv_CTC_Transformer is ns.transformer.Conv_CTC_Transformer
    # models built by the reference's own create_model now hold the drop-in attention, with unchanged state_dict keys
    for cls, args in ((CIF_Model, small_args()), (Conv_CTC_Transformer, small_args()), (Transformer, small_args())):
        m = cls.create_model(args)
        att = [x for x in m.modules() if type(x).__name__ in ("MultiheadAttention", "MultiHeadAttention")]
        assert att and all(isinstance(x, ours_att.MultiheadAttention) for x in att), cls
        assert all(x.return_attn is None and not ours_att.MultiheadAttention.RETURN_ATTN_DEFAULT for x in att)
    after = CIF_Model.create_model(small_args())
    assert {k: tuple(v.shape) for k, v in after.state_dict().items()} == keys_before
    before.load_state_dict(after.state_dict())           # checkpoints interchange both ways
    # install twice = same state, uninstall = everything back
    patch.install()
    patch.uninstall()
    assert ns.cif_model.CIF_Model.cif is ref_cif_fn and ns.attention.MultiheadAttention is ref_mha
    assert ns.encoder.MultiheadAttention is ref_mha and ns.cattention.MultiHeadAttention is ref_cmha
    assert ns.solver.cal_ctc_ce_loss is ref_ctc_ce and ns.solver.cal_ctc_qua_ce_loss is ref_qua and ns.closs.cal_loss is ref_cal
    assert ours_att.MultiheadAttention.RETURN_ATTN_DEFAULT is True
    assert "transformer.Transformer" not in sys.modules and "transformer.CIF_Model" not in sys.modules
    print("ok")
    """)


@needs_ref
@pytest.mark.gpu
def test_reference_built_cif_model_trains_through_the_drop_ins():
    """A CIF_Model constructed by the reference's classes after install(), the golden weights, one solver-style step
    (solver.py:141-157) on the GPU: against the reference's own CPU run of the same step (tests/golden/cif_model.npz)."""
    _run("""
    G = np.load(os.path.join(%r, "tests", "golden", "cif_model.npz"))
    patch.install()
    launches0 = importlib.import_module(PKG + "._lib").launch_count()
    from transformer.CIF_Model import CIF_Model
    from transformer.solver import cal_ctc_qua_ce_loss          # the name the solver calls: now ours
    assert cal_ctc_qua_ce_loss is ours_loss.cal_ctc_qua_ce_loss
    model = CIF_Model.create_model(small_args())
    sd = {k[3:]: torch.as_tensor(G[k]) for k in G.files if k.startswith("sd:")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(".pe") for k in missing)
    model.cuda().eval()                                           # dropout off: the golden run is eval mode
    feats, lens, targets = (torch.as_tensor(G[k]).cuda() for k in ("feats", "lens", "targets"))
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.98), eps=1e-9)
    torch.manual_seed(int(G["rand_seed"]))                        # the torch.rand(B) of cif_model.py:47
    ctc_logits, len_ctc, _num, num, logits = model(feats, lens, targets)
    qua, ctc, ce = cal_ctc_qua_ce_loss(ctc_logits, len_ctc, _num, num, logits, targets, smoothing=0.1)
    loss = 0.001 * qua + ctc + ce
    opt.zero_grad()
    loss.backward()
    opt.step()
    used = importlib.import_module(PKG + "._lib").launch_count() - launches0
    assert used >= 8, "the drop-in kernels did not run (%%d launches)" %% used
    to_np = lambda t: t.detach().float().cpu().numpy()
    np.testing.assert_array_equal(to_np(len_ctc), G["len_ctc"])
    np.testing.assert_array_equal(to_np(num), G["num"])
    assert logits.shape == G["logits"].shape                     # same number of fired rows
    # bf16 attention core inside an fp32 model: 2e-2 of the activation scale
    assert np.abs(to_np(ctc_logits) - G["ctc_logits"]).max() <= 2e-2 * np.abs(G["ctc_logits"]).max()
    np.testing.assert_allclose(float(qua), G["qua"], rtol=2e-2)
    np.testing.assert_allclose(float(ctc), G["ctc"], rtol=2e-2)
    np.testing.assert_allclose(float(ce), G["ce"], rtol=2e-2)
    params = dict(model.named_parameters())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params.values())
    for k in [k for k in G.files if k.startswith("grad:")]:
        ref = G[k]
        got = to_np(params[k[5:]].grad)
        assert np.abs(got - ref).max() <= 6e-2 * np.abs(ref).max() + 1e-7, (k, np.abs(got - ref).max(), np.abs(ref).max())
    # training mode (attention dropout inside the kernels) runs too
    model.train()
    for _ in range(2):
        out = model(feats, lens, targets)
        l3 = cal_ctc_qua_ce_loss(out[0], out[1], out[2], out[3], out[4], targets, smoothing=0.1)
        opt.zero_grad(); (0.001 * l3[0] + l3[1] + l3[2]).backward(); opt.step()
    assert torch.isfinite(l3[1]) and torch.isfinite(l3[2])
    print("ok", float(qua), float(ctc), float(ce), used)
    """ % ROOT)


@needs_ref
@pytest.mark.gpu
def test_reference_built_conv_ctc_transformer_trains_through_the_drop_ins():
    """Config-2's model (Conv_CTC_Transformer, transformer.py:130-222) built by the reference after install():
    Transformer_CTC_Solver's step (solver.py:83-95) - decoder self / cross attention with the reference's DENSE masks."""
    _run("""
    patch.install()
    from transformer.Transformer import Conv_CTC_Transformer
    from transformer.solver import cal_ctc_ce_loss
    torch.manual_seed(5)
    model = Conv_CTC_Transformer.create_model(small_args()).cuda().train()
    g = torch.Generator().manual_seed(6)
    feats = torch.randn(4, 167, 320, generator=g).cuda()
    lens = torch.tensor([167, 150, 141, 120]).cuda()
    targets = torch.randint(4, 99, (4, 9), generator=g)
    targets[1, 7:] = 0; targets[3, 5:] = 0
    targets = targets.cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    losses = []
    for _ in range(3):
        logits_ctc, len_logits_ctc, logits_ce, targets_eos = model(feats, lens, targets)
        ctc, ce = cal_ctc_ce_loss(logits_ctc, len_logits_ctc, logits_ce, targets_eos, smoothing=0.1)
        opt.zero_grad(); (ctc + ce).backward(); opt.step()
        losses.append(float(ctc + ce))
    assert all(np.isfinite(losses)), losses
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    print("ok", losses)
    """)
