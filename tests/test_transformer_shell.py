"""The encoder-decoder shells of BASELINE configs 2 and 3 (`Transformer`, `Conv_CTC_Transformer`, `Decoder`,
`DecoderLayer`: callers of the attention / CTC hot path) against small models run through the REFERENCE's own
classes (tests/golden/transformer_models.npz; transformer.py:21-35,135-153, decoder.py:41-96,617-636)."""
import argparse

import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import pkg, to_np
from test_model_shell import _torch_fp32_core

G = load_golden("transformer_models")


def _state(prefix):
    tag = prefix + "sd:"
    return {k[len(tag):]: torch.as_tensor(G[k]) for k in G.files if k.startswith(tag)}


def _build_transformer():
    tr, en, de = pkg("transformer.transformer"), pkg("transformer.encoder"), pkg("transformer.decoder")
    return tr.Transformer(en.Encoder(d_input=320, n_layers=2, n_head=2, d_model=64, d_inner=128, dropout=0.1),
                          de.Decoder(sos_id=2, eos_id=3, n_tgt_vocab=100, n_layers=2, n_head=2, d_model=64, d_inner=128,
                                     dropout=0.1))


def _build_conv_ctc():
    tr, en, de, ce = (pkg("transformer.transformer"), pkg("transformer.encoder"), pkg("transformer.decoder"),
                      pkg("transformer.conv_encoder"))
    return tr.Conv_CTC_Transformer(ce.Conv2dSubsample(d_input=320, d_model=64, n_layers=3),
                                   en.Encoder(d_input=64, n_layers=1, n_head=2, d_model=64, d_inner=128, dropout=0.1),
                                   de.Decoder(sos_id=2, eos_id=3, n_tgt_vocab=100, n_layers=1, n_head=2, d_model=64,
                                              d_inner=128, dropout=0.1))


@pytest.mark.parametrize("build,prefix", [(_build_transformer, "t:"), (_build_conv_ctc, "c:")])
def test_state_dict_keys_match_the_reference(build, prefix):
    model = build()
    ours = {k: tuple(v.shape) for k, v in model.state_dict().items() if not k.endswith(".pe")}
    ref = {k: tuple(v.shape) for k, v in _state(prefix).items()}
    assert ours == ref
    missing, unexpected = model.load_state_dict(_state(prefix), strict=False)
    assert not unexpected and all(k.endswith(".pe") for k in missing)


@pytest.mark.parametrize("prefix", ["t:", "c:"])
def test_decoder_preprocess_matches_the_reference(prefix):
    """<sos>/<eos> framing and padding (decoder.py:41-58), incl. the batch whose longest row is shorter than To."""
    dec = pkg("transformer.decoder").Decoder(sos_id=2, eos_id=3, n_tgt_vocab=100, n_layers=1, n_head=2, d_model=64, d_inner=128)
    targets = torch.as_tensor(G[prefix + "targets"])
    ys_in, ys_out = dec.preprocess(targets)
    np.testing.assert_array_equal(ys_out.numpy(), G[prefix + "targets_eos"])
    assert ys_in.shape == ys_out.shape and (ys_in[:, 0] == 2).all()
    np.testing.assert_array_equal(ys_in[:, 1:].numpy(), np.where(G[prefix + "targets_eos"][:, :-1] == 3, 0, G[prefix + "targets_eos"][:, :-1]) *
                                  (np.arange(1, ys_in.shape[1])[None, :] <= (targets.numpy() != 0).sum(1)[:, None]))
    short = targets[3:]                                   # no full-length row: width = max_b(len_b) + 1 < To + 1
    a, b = dec.preprocess(short)
    assert a.shape[1] == int((short != 0).sum(1).max()) + 1
    dec.assume_full_width = True                          # capture-friendly sizing: same content, padded to To + 1
    a2, b2 = dec.preprocess(short)
    assert a2.shape[1] == short.shape[1] + 1
    np.testing.assert_array_equal(a2[:, :a.shape[1]].numpy(), a.numpy())
    np.testing.assert_array_equal(b2[:, :b.shape[1]].numpy(), b.numpy())
    assert (a2[:, a.shape[1]:] == 0).all() and (b2[:, b.shape[1]:] == 0).all()


def test_create_model_and_alias_modules():
    args = argparse.Namespace(d_input=80, LFR_m=4, n_conv_layers=3, d_model=64, n_layers_enc=1, n_head=2, d_inner=128,
                              dropout=0.1, sos_id=2, eos_id=3, vocab_size=50, n_layers_dec=1, spec_aug_cfg=None)
    alias = pkg("transformer.Transformer")               # the module name train.py:139-151 imports
    assert alias.Transformer is pkg("transformer.transformer").Transformer
    assert pkg("transformer.CIF_Model").CIF_Model is pkg("transformer.cif_model").CIF_Model
    m = alias.Transformer.create_model(args)              # the reference's version of this recurses (transformer.py:72)
    assert m.encoder.linear_in.in_features == 320 and m.decoder.tgt_word_prj.out_features == 50
    m = alias.CTC_Transformer.create_model(args)
    assert m.ctc_fc.weight.shape == (50, 64)
    m = alias.Conv_CTC_Transformer.create_model(args)
    assert m.ctc_fc.weight.shape == (50, 64) and m.conv_encoder.affine.in_features == 32 * 160


def _run_transformer(model):
    tl = pkg("transformer.loss")
    feats, lens, targets = (torch.as_tensor(G["t:" + k]).cuda() for k in ("feats", "lens", "targets"))
    logits, targets_eos = model(feats, lens, targets)
    ce = tl.cal_ce_loss(logits, targets_eos, smoothing=0.1)
    ce.backward()
    return logits, targets_eos, ce


def _run_conv_ctc(model):
    tl = pkg("transformer.loss")
    feats, lens, targets = (torch.as_tensor(G["c:" + k]).cuda() for k in ("feats", "lens", "targets"))
    ctc_logits, len_ctc, logits, targets_eos = model(feats, lens, targets)
    ctc, ce = tl.cal_ctc_ce_loss(ctc_logits, len_ctc, logits, targets_eos, smoothing=0.1)
    (ctc + ce).backward()
    return ctc_logits, len_ctc, logits, targets_eos, ctc, ce


def _check_grads(model, prefix, tol):
    params = dict(model.named_parameters())
    tag = prefix + "grad:"
    for k in [k for k in G.files if k.startswith(tag)]:
        ref, got = G[k], to_np(params[k[len(tag):]].grad)
        assert np.abs(got - ref).max() <= tol * np.abs(ref).max() + 1e-7, (k, np.abs(got - ref).max(), np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("tc_fp32", [False, True], ids=["torch_sgemm", "tensor_core_3xtf32"])
def test_transformer_matches_reference_with_fp32_attention(monkeypatch, tc_fp32):
    monkeypatch.setattr(pkg("transformer.attention"), "mha_core", _torch_fp32_core)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    # the shell's Linear layers: torch's SIMT sgemm (7e-7 per product) or this package's 3xTF32 tensor-core GEMMs (~3e-6);
    # the network amplifies either (softmax / LayerNorm backward are differences of near-equal numbers)
    monkeypatch.setattr(pkg("transformer.module"), "USE_TENSOR_CORE_FP32", tc_fp32)
    model = _build_transformer()
    model.load_state_dict(_state("t:"), strict=False)
    model = model.cuda().eval()
    logits, targets_eos, ce = _run_transformer(model)
    np.testing.assert_array_equal(to_np(targets_eos), G["t:targets_eos"])
    np.testing.assert_allclose(to_np(logits), G["t:logits"], rtol=1e-3, atol=5e-4)
    np.testing.assert_allclose(float(ce), G["t:ce"], rtol=1e-4)
    _check_grads(model, "t:", 6e-3 if tc_fp32 else 2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("tc_fp32", [False, True], ids=["torch_sgemm", "tensor_core_3xtf32"])
def test_conv_ctc_transformer_matches_reference_with_fp32_attention(monkeypatch, tc_fp32):
    monkeypatch.setattr(pkg("transformer.attention"), "mha_core", _torch_fp32_core)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    # the shell's Linear layers: torch's SIMT sgemm (7e-7 per product) or this package's 3xTF32 tensor-core GEMMs (~3e-6);
    # the network amplifies either (softmax / LayerNorm backward are differences of near-equal numbers)
    monkeypatch.setattr(pkg("transformer.module"), "USE_TENSOR_CORE_FP32", tc_fp32)
    model = _build_conv_ctc()
    model.load_state_dict(_state("c:"), strict=False)
    model = model.cuda().eval()
    ctc_logits, len_ctc, logits, targets_eos, ctc, ce = _run_conv_ctc(model)
    np.testing.assert_array_equal(to_np(len_ctc), G["c:len_ctc"])
    np.testing.assert_array_equal(to_np(targets_eos), G["c:targets_eos"])
    np.testing.assert_allclose(to_np(ctc_logits), G["c:ctc_logits"], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(to_np(logits), G["c:logits"], rtol=1e-3, atol=5e-4)
    np.testing.assert_allclose(float(ctc), G["c:ctc"], rtol=1e-4)      # the fused sm_100a CTC with <eos>-extended targets
    np.testing.assert_allclose(float(ce), G["c:ce"], rtol=1e-4)
    _check_grads(model, "c:", 6e-3 if tc_fp32 else 2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["t:", "c:"])
def test_shells_with_tcgen05_attention_are_close(which):
    """Self-attention (causal + key padding as kv_len), cross attention (key padding) and encoder attention on the
    tcgen05 core, bf16: 2e-2 of the activation scale; gradients finite and close."""
    model = (_build_transformer if which == "t:" else _build_conv_ctc)()
    model.load_state_dict(_state(which), strict=False)
    model = model.cuda().eval()
    if which == "t:":
        logits, _, ce = _run_transformer(model)
    else:
        _, _, logits, _, ctc, ce = _run_conv_ctc(model)
        np.testing.assert_allclose(float(ctc), G["c:ctc"], rtol=2e-2)
    assert np.abs(to_np(logits) - G[which + "logits"]).max() <= 2e-2 * np.abs(G[which + "logits"]).max()
    np.testing.assert_allclose(float(ce), G[which + "ce"], rtol=2e-2)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    _check_grads(model, which, 8e-2)


@pytest.mark.gpu
def test_bf16_model_runs_the_fused_evaluation_path(monkeypatch):
    """A `.bfloat16()` CIF_Model end to end in evaluation: the non-pad mask follows the activations' dtype,
    so the fused tcgen05 linear layers are reachable from the model shell (d_model = 512), and the CIF boundary casts
    to fp32."""
    cm, ce, en, aa, de = (pkg("transformer.cif_model"), pkg("transformer.conv_encoder"), pkg("transformer.encoder"),
                          pkg("transformer.attentionAssigner"), pkg("transformer.decoder"))
    att, ops = pkg("transformer.attention"), pkg("ops")
    calls = {"ln": 0, "act": 0}
    real_ln, real_act = ops.linear_residual_layernorm, ops.linear_act

    def ln(*a, **k):
        calls["ln"] += 1
        return real_ln(*a, **k)

    def act(*a, **k):
        calls["act"] += 1
        return real_act(*a, **k)
    for mod in (att, ops):
        monkeypatch.setattr(mod, "linear_residual_layernorm", ln)
        monkeypatch.setattr(mod, "linear_act", act)

    def build():
        torch.manual_seed(11)
        return cm.CIF_Model(ce.Conv2dSubsample(d_input=320, d_model=512, n_layers=3),
                            en.Encoder(d_input=512, n_layers=2, n_head=8, d_model=512, d_inner=1024, dropout=0.1),
                            aa.Attention_Assigner(d_input=512, d_hidden=128, w_context=3, n_layers=2),
                            de.Decoder_CIF(sos_id=2, n_tgt_vocab=100, n_layers=1, n_head=8, d_model=512, d_inner=1024,
                                           dropout=0.1)).cuda().eval()
    ref, low = build(), build().bfloat16()
    g = torch.Generator().manual_seed(12)
    feats = torch.randn(4, 167, 320, generator=g).cuda()
    lens = torch.tensor([167, 160, 149, 121]).cuda()
    targets = torch.randint(4, 99, (4, 8), generator=g)
    targets[2, 6:] = 0
    targets = targets.cuda()
    with torch.no_grad():
        torch.manual_seed(3)
        a = ref(feats, lens, targets)
        assert calls["ln"] == 0                                   # the fp32 model does not take the bf16 route
        torch.manual_seed(3)
        b = low(feats.bfloat16(), lens, targets)
    assert calls["ln"] == 2 * (2 + 1) and calls["act"] == 4 * (2 + 1), calls     # per layer: fc+LN, w_2+LN; q, k, v, w_1
    assert b[0].dtype == torch.bfloat16 and b[4].dtype == torch.bfloat16
    assert torch.isfinite(b[0].float()).all() and torch.isfinite(b[4].float()).all()
    # encoder side (before any fire decision): bf16 model against the fp32 model, 8e-2 of the logit scale
    assert (b[0].float() - a[0]).abs().max() <= 8e-2 * a[0].abs().max()
