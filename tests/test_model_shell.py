"""The caller-side model shell (conv front end, encoder, assigner, Decoder_CIF) around the
drop-in CIF / CTC / attention modules, against a small CIF_Model run through the REFERENCE
classes (tests/golden/cif_model.npz: BASELINE config-1-shaped inputs, 8 x 167 x 320)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import pkg, to_np

G = load_golden("cif_model")


def _build():
    cm = pkg("transformer.cif_model")
    ce = pkg("transformer.conv_encoder")
    en = pkg("transformer.encoder")
    aa = pkg("transformer.attentionAssigner")
    de = pkg("transformer.decoder")
    model = cm.CIF_Model(ce.Conv2dSubsample(d_input=320, d_model=64, n_layers=3),
                         en.Encoder(d_input=64, n_layers=1, n_head=2, d_model=64, d_inner=128, dropout=0.1),
                         aa.Attention_Assigner(d_input=64, d_hidden=64, w_context=3, n_layers=3),
                         de.Decoder_CIF(sos_id=2, n_tgt_vocab=100, n_layers=1, n_head=2, d_model=64, d_inner=128,
                                        dropout=0.1))
    return model


def _golden_state():
    return {k[3:]: torch.as_tensor(G[k]) for k in G.files if k.startswith("sd:")}


def test_state_dict_keys_match_the_reference():
    model = _build()
    ours = {k: tuple(v.shape) for k, v in model.state_dict().items() if not k.endswith(".pe")}
    ref = {k: tuple(v.shape) for k, v in _golden_state().items()}
    assert ours == ref
    missing, unexpected = model.load_state_dict(_golden_state(), strict=False)
    assert not unexpected and all(k.endswith(".pe") for k in missing)


def test_create_model_builds_the_recipe_default():
    import argparse
    cm = pkg("transformer.cif_model")
    args = argparse.Namespace(d_input=80, LFR_m=4, n_conv_layers=3, d_model=64, n_layers_enc=1, n_head=2, d_inner=128,
                              dropout=0.1, d_assigner_hidden=64, w_context=3, n_assigner_layers=3, sos_id=2,
                              vocab_size=50, n_layers_dec=1, spec_aug_cfg=None)
    m = cm.CIF_Model.create_model(args)
    assert m.ctc_fc.weight.shape == (50, 64) and m.conv_encoder.affine.in_features == 32 * 160


def _torch_fp32_core(q, k, v, kv_len=None, mask=None, causal=False, scale=None, dropout_p=0.0, seed=None):
    """fp32 stand-in for ops.mha_core with the same signature (test only: isolates the CIF / CTC
    drop-ins from bf16 rounding in the attention, which can move a fire by one frame)."""
    assert dropout_p == 0.0, "the stand-in is for eval-mode parity runs"
    B, Lq, H, D = q.shape
    Lk = k.shape[1]
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * (scale or 1.0 / D ** 0.5)
    dead = torch.zeros(B, 1, Lq, Lk, dtype=torch.bool, device=q.device)
    if kv_len is not None:
        dead = dead | (torch.arange(Lk, device=q.device)[None, None, None, :] >= kv_len.view(B, 1, 1, 1))
    if causal:
        dead = dead | torch.triu(torch.ones(Lq, Lk, dtype=torch.bool, device=q.device), diagonal=1)[None, None]
    if mask is not None:
        dead = dead | mask.bool()[:, None]
    p = torch.softmax(s.masked_fill(dead, float("-inf")), dim=-1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v.float())


def _run(model):
    tl = pkg("transformer.loss")
    dev = "cuda"
    feats = torch.as_tensor(G["feats"]).to(dev)
    lens = torch.as_tensor(G["lens"]).to(dev)
    targets = torch.as_tensor(G["targets"]).to(dev)
    torch.manual_seed(int(G["rand_seed"]))
    ctc_logits, len_ctc, _num, num, logits = model(feats, lens, targets)
    qua, ctc, ce = tl.cal_ctc_qua_ce_loss(ctc_logits, len_ctc, _num, num, logits, targets, smoothing=0.1)
    (0.001 * qua + ctc + ce).backward()
    return ctc_logits, len_ctc, _num, num, logits, qua, ctc, ce


@pytest.mark.gpu
@pytest.mark.parametrize("tc_fp32", [False, True], ids=["torch_sgemm", "tensor_core_3xtf32"])
def test_full_forward_backward_matches_reference_with_fp32_attention(monkeypatch, tc_fp32):
    att = pkg("transformer.attention")
    monkeypatch.setattr(att, "mha_core", _torch_fp32_core)
    # the caller-side convolutions / GEMMs must run in true fp32 for a 1e-3 comparison (cuDNN uses TF32 by default)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    # the shell's Linear layers: torch's SIMT sgemm (7e-7 per product) or this package's 3xTF32 tensor-core GEMMs (~3e-6);
    # the network amplifies either (softmax / LayerNorm backward are differences of near-equal numbers)
    monkeypatch.setattr(pkg("transformer.module"), "USE_TENSOR_CORE_FP32", tc_fp32)
    model = _build()
    model.load_state_dict(_golden_state(), strict=False)
    model = model.cuda().eval()
    ctc_logits, len_ctc, _num, num, logits, qua, ctc, ce = _run(model)
    np.testing.assert_array_equal(to_np(len_ctc), G["len_ctc"])
    np.testing.assert_array_equal(to_np(num), G["num"])
    np.testing.assert_allclose(to_np(_num), G["_num"], rtol=1e-4)
    np.testing.assert_allclose(to_np(ctc_logits), G["ctc_logits"], rtol=1e-3, atol=2e-4)
    assert logits.shape == G["logits"].shape                 # same L: same number of fired rows
    np.testing.assert_allclose(to_np(logits), G["logits"], rtol=1e-3, atol=5e-4)
    np.testing.assert_allclose(float(qua), G["qua"], rtol=1e-4)
    np.testing.assert_allclose(float(ctc), G["ctc"], rtol=1e-4)
    np.testing.assert_allclose(float(ce), G["ce"], rtol=1e-4)
    params = dict(model.named_parameters())
    for k in [k for k in G.files if k.startswith("grad:")]:
        ref = G[k]
        got = to_np(params[k[5:]].grad)
        assert np.abs(got - ref).max() <= (6e-3 if tc_fp32 else 2e-3) * np.abs(ref).max() + 1e-7, k


@pytest.mark.gpu
def test_full_forward_with_tcgen05_attention_is_close():
    model = _build()
    model.load_state_dict(_golden_state(), strict=False)
    model = model.cuda().eval()
    ctc_logits, len_ctc, _num, num, logits, qua, ctc, ce = _run(model)
    # bf16 attention inside the encoder: 2e-2 of the activation scale
    assert np.abs(to_np(ctc_logits) - G["ctc_logits"]).max() <= 2e-2 * np.abs(G["ctc_logits"]).max()
    np.testing.assert_allclose(float(ctc), G["ctc"], rtol=2e-2)
    np.testing.assert_allclose(float(qua), G["qua"], rtol=2e-2)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
