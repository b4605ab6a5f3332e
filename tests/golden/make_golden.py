#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE.

Run in the build container only (it needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

The reference has no test fixtures for the hot path, so
the oracle is pinned against outputs of the reference's own modules, imported
read-only from /root/reference/src with the documented shims:

  1. torch.Tensor.cuda -> identity (hard-coded .cuda() in cif_model.py:47,61,
     62,69,76,100; this container has no GPU).
  2. nothing else is needed for CIF_Model.cif, transformer.loss,
     ctcModel.loss and transformer.attention.

Every array is produced by reference code:
  * CIF      : transformer.cif_model.CIF_Model.cif  (cif_model.py:57-106) and
               torch autograd through it (backward goldens).
  * CIF glue : the scaling lines cif_model.py:43-48 executed verbatim through
               CIF_Model.forward is not possible without the whole model, so
               the glue golden stores the inputs/outputs of those four lines
               re-executed here with the torch.rand(B) draw stored alongside.
  * CIF_Model: a small model built from the reference's own classes (conv front
               end, encoder, assigner, Decoder_CIF) run through CIF_Model.forward,
               cal_ctc_qua_ce_loss and autograd on config-1-shaped inputs.
  * CTC      : transformer.loss.cal_ctc_ce_loss (loss.py:34-48) and
               ctcModel.loss.cal_loss (ctcModel/loss.py:4-13); per-utterance
               nll from F.ctc_loss(reduction='none') on the same log-probs.
  * MHA      : transformer.attention.MultiheadAttention (attention.py:6-86),
               eval mode (dropout off), with the masks from utils/utils.py.
  * Transformer / Conv_CTC_Transformer (transformer.py): small models of BASELINE
               configs 3 and 2 through forward, cal_ce_loss / cal_ctc_ce_loss and autograd.
Outputs are small .npz files committed next to this script.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference not present; goldens can only be regenerated in the build container")
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self  # shim 1
    import transformer.cif_model as cif_model
    import transformer.loss as tloss
    import ctcModel.loss as closs
    import transformer.attention as attention
    import utils.utils as uutils
    return cif_model, tloss, closs, attention, uutils


def _ref_cif(cif_model, hidden, alphas, thr):
    """Call the unbound reference method (it never touches self)."""
    return cif_model.CIF_Model.cif(None, hidden, alphas, thr)


def _fire_positions(alphas_row, thr):
    """Fire positions re-derived from the reference's intermediate definition
    (`fires` = pre-reset integrate, cif_model.py:71-72,97-99) by running the
    same scalar torch ops the reference runs, one utterance at a time."""
    integrate = torch.zeros([1])
    pos = []
    for t in range(alphas_row.numel()):
        integrate = integrate + alphas_row[t]
        if bool(integrate > thr):
            pos.append(t)
            integrate = integrate - torch.ones([1])
    return pos


def make_cif(cif_model):
    out = {}
    cases = {
        # name: (B, T, H, alpha_kind)
        "a": (3, 37, 8, "sigmoid"),
        "b": (8, 21, 32, "scaled14"),      # cfg-1-like: 21 encoder frames, 14 labels
        "c": (2, 64, 16, "big"),           # alphas > 1 -> back-to-back fires / negative remainders
        "d": (4, 50, 12, "ragged"),        # zero alphas beyond a per-utterance length
        "e": (2, 5, 4, "zero"),            # no fires at all -> L = 0
        "f": (1, 1, 4, "one"),             # T = 1
        "g": (2, 130, 128, "scaled30"),    # one full TMA slice wide
    }
    g = torch.Generator().manual_seed(20261017)
    for name, (B, T, H, kind) in cases.items():
        hidden = torch.randn(B, T, H, generator=g)
        if kind == "sigmoid":
            alphas = torch.sigmoid(torch.randn(B, T, generator=g))
        elif kind.startswith("scaled"):
            n = float(kind[6:])
            alphas = torch.sigmoid(torch.randn(B, T, generator=g))
            noise = torch.rand(B, generator=g) - 0.5
            alphas = alphas * ((n + noise) / alphas.sum(-1))[:, None]
        elif kind == "big":
            alphas = torch.rand(B, T, generator=g) * 1.7
        elif kind == "ragged":
            alphas = torch.sigmoid(torch.randn(B, T, generator=g))
            lens = torch.tensor([50, 41, 33, 7])
            alphas = alphas * (torch.arange(T)[None, :] < lens[:, None]).float()
        elif kind == "zero":
            alphas = torch.zeros(B, T)
        elif kind == "one":
            alphas = torch.full((B, T), 0.97)
        thr = 0.95
        hid = hidden.clone().requires_grad_(True)
        alp = alphas.clone().requires_grad_(True)
        y = _ref_cif(cif_model, hid, alp, thr)
        L = y.size(1)
        g_out = torch.randn(B, L, H, generator=g)
        if L > 0:
            (y * g_out).sum().backward()
            g_hidden, g_alpha = hid.grad, alp.grad
        else:
            g_hidden, g_alpha = torch.zeros_like(hidden), torch.zeros_like(alphas)
        fire_t = np.full((B, max(L, 1)), -1, dtype=np.int32)
        n_fired = np.zeros((B,), dtype=np.int32)
        for b in range(B):
            pos = _fire_positions(alphas[b], thr)
            n_fired[b] = len(pos)
            fire_t[b, :len(pos)] = pos
        out.update({
            f"{name}_hidden": hidden.numpy(), f"{name}_alphas": alphas.numpy(),
            f"{name}_thr": np.float32(thr), f"{name}_out": y.detach().numpy(),
            f"{name}_g_out": g_out.numpy(), f"{name}_g_hidden": g_hidden.numpy(),
            f"{name}_g_alpha": g_alpha.numpy(), f"{name}_fire_t": fire_t,
            f"{name}_n_fired": n_fired,
        })
        print(f"cif {name}: B{B} T{T} H{H} {kind}: L={L} fired={n_fired.tolist()}")
    np.savez_compressed(os.path.join(HERE, "cif.npz"), **out)


def make_cif_glue():
    """cif_model.py:43-48 with the rand() term fixed (noise = rand - 0.5)."""
    g = torch.Generator().manual_seed(77)
    B, T = 5, 29
    alpha = torch.sigmoid(torch.randn(B, T, generator=g))
    lens = torch.tensor([29, 29, 20, 11, 3])
    alpha = alpha * (torch.arange(T)[None, :] < lens[:, None]).float()
    targets = torch.randint(1, 40, (B, 9), generator=g)
    targets[2, 6:] = 0
    targets[4, 2:] = 0
    rnd = torch.rand(B, generator=g)
    a = alpha.clone()
    _num = a.sum(-1)                                           # :44
    num = (targets > 0).float().sum(-1)                        # :46
    num_noise = num + rnd - 0.5                                # :47
    a *= (num_noise / _num)[:, None].repeat(1, a.size(1))      # :48
    np.savez_compressed(os.path.join(HERE, "cif_glue.npz"),
                        alpha=alpha.numpy(), targets=targets.numpy(), rand=rnd.numpy(),
                        _num=_num.numpy(), num=num.numpy(), scaled=a.numpy())
    print("cif glue: done")


def make_assigner_tail(uutils):
    """attentionAssigner.py:36-40 (linear -> sigmoid -> sequence_mask, the reference's own mask builder)
    followed by cif_model.py:43-48, with autograd gradients for a fixed upstream gradient."""
    g = torch.Generator().manual_seed(78)
    B, T, D = 4, 23, 36
    x = torch.randn(B, T, D, generator=g, requires_grad=True)
    lin = torch.nn.Linear(D, 1)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(1, D, generator=g) * 0.3)
        lin.bias.copy_(torch.randn(1, generator=g) * 0.1)
    lens = torch.tensor([23, 17, 9, 1])
    targets = torch.randint(1, 40, (B, 7), generator=g)
    targets[1, 5:] = 0
    targets[3, 1:] = 0
    rnd = torch.rand(B, generator=g)
    alphas = lin(x).squeeze(-1)                                   # attentionAssigner.py:37
    alphas = torch.sigmoid(alphas)                                # :38
    pad_mask = uutils.sequence_mask(lens)                         # :39
    alpha = alphas * pad_mask                                     # :41
    _num = alpha.sum(-1)                                          # cif_model.py:44
    num = (targets > 0).float().sum(-1)                           # :46
    num_noise = num + rnd - 0.5                                   # :47
    scaled = alpha * (num_noise / _num)[:, None].repeat(1, alpha.size(1))   # :48 (out of place for autograd)
    g_alpha = torch.randn(B, T, generator=g)
    g_num = torch.randn(B, generator=g)
    ((scaled * g_alpha).sum() + (_num * g_num).sum()).backward()
    np.savez_compressed(os.path.join(HERE, "assigner_tail.npz"),
                        x=x.detach().numpy(), w=lin.weight.detach().numpy(), b=lin.bias.detach().numpy(),
                        lens=lens.numpy(), num_noise=num_noise.detach().numpy(), alpha_raw=alpha.detach().numpy(),
                        _num=_num.detach().numpy(), scaled=scaled.detach().numpy(), g_alpha=g_alpha.numpy(),
                        g_num=g_num.numpy(), g_x=x.grad.numpy(), g_w=lin.weight.grad.numpy(), g_b=lin.bias.grad.numpy())
    print("assigner tail: done")


def make_spec_aug(uutils):
    """utils/utils.py:168-194 executed on CPU; the torch.rand draws are recorded so that the device kernels and
    the oracle can be fed the very same bands / spans, and the generator seed so that the host mirror's draws
    (utils.spec_aug_draw) can be checked against them."""
    out = {}
    real_rand = torch.rand
    for name, (B, T, V, cfg, seed) in {"a": (4, 60, 80, "2-27-2-40", 5), "b": (2, 90, 320, "2-27-2-40", 6),
                                       "c": (2, 33, 40, "1-10-3-8", 7), "d": (5, 100, 33, "2-5-1-30", 8)}.items():
        g = torch.Generator().manual_seed(100 + seed)
        x = torch.randn(B, T, V, generator=g)
        lens = torch.randint(T // 2 + 1, T + 1, (B,), generator=g)
        lens[0] = T
        for b in range(B):
            x[b, int(lens[b]):] = 0.0          # "features are padded with zeros" (utils.py:173)
        draws = []

        def recording_rand(*a, **k):
            r = real_rand(*a, **k)
            draws.append(r.clone())
            return r
        torch.manual_seed(seed)
        torch.rand = recording_rand
        try:
            y, _ = uutils.spec_aug(x.clone(), lens, cfg)
        finally:
            torch.rand = real_rand
        out[name + "_x"] = x.numpy()
        out[name + "_lens"] = lens.numpy()
        out[name + "_cfg"] = np.array([int(i) for i in cfg.split("-")])
        out[name + "_seed"] = np.array([seed])
        out[name + "_draws"] = torch.stack(draws).numpy()      # [4 * time_mask_num, B] in call order
        out[name + "_y"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "spec_aug.npz"), **out)
    print("spec_aug: done")


def make_lfr():
    """utils/data.py:191-218 executed on a few utterances (the function is pure numpy)."""
    import importlib
    import types
    # utils/data.py imports kaldi_io (not installed here) for its ark readers; the function under
    # test is pure numpy, so an empty stand-in module is enough to execute the reference file
    sys.modules.setdefault("kaldi_io", types.ModuleType("kaldi_io"))
    udata = importlib.import_module("utils.data")
    rng = np.random.default_rng(79)
    out = {}
    for name, (T, D, m, n) in {"a": (17, 8, 4, 3), "b": (16, 5, 4, 3), "c": (3, 8, 4, 3), "d": (9, 4, 1, 1),
                               "e": (10, 6, 1, 2), "f": (7, 3, 3, 1), "g": (1, 4, 4, 3)}.items():
        x = rng.normal(size=(T, D)).astype(np.float32)
        out[name + "_x"] = x
        out[name + "_mn"] = np.array([m, n])
        out[name + "_y"] = udata.build_LFR_features(x, m, n).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "lfr.npz"), **out)
    print("lfr: done")


def make_ctc(tloss, closs):
    out = {}
    g = torch.Generator().manual_seed(4233)
    cases = {
        # name: (B, T, V, S, kind)
        "a": (3, 12, 7, 4, "repeats_ragged"),
        "b": (4, 30, 50, 8, "ragged"),
        "c": (2, 6, 5, 4, "infeasible"),    # T < S + repeats for utt 0 -> inf
        "d": (3, 10, 6, 3, "empty_target"),  # target_length 0 for utt 1
        "e": (2, 14, 4233, 5, "aishell_vocab"),
        "f": (5, 25, 33, 12, "full_len"),
    }
    for name, (B, T, V, S, kind) in cases.items():
        logits = torch.randn(B, T, V, generator=g)
        targets = torch.randint(1, V - 1, (B, S), generator=g)
        in_len = torch.full((B,), T, dtype=torch.int32)
        if kind == "repeats_ragged":
            targets[0] = torch.tensor([2, 2, 3, 3])
            targets[1, 3:] = 0
            targets[2, 2:] = 0
            in_len = torch.tensor([12, 9, 7], dtype=torch.int32)
        elif kind == "ragged":
            targets[1, 5:] = 0
            targets[3, 1:] = 0
            targets[2, 2] = targets[2, 1]
            in_len = torch.tensor([30, 22, 17, 30], dtype=torch.int32)
        elif kind == "infeasible":
            targets[0] = torch.tensor([1, 1, 1, 1])   # needs 4 + 3 = 7 > 6 frames
            targets[1, 2:] = 0
        elif kind == "empty_target":
            targets[1, :] = 0
            in_len = torch.tensor([10, 8, 10], dtype=torch.int32)
        elif kind == "aishell_vocab":
            targets[1, 3:] = 0
            in_len = torch.tensor([14, 11], dtype=torch.int32)
        lg = logits.clone().requires_grad_(True)
        # reference call convention A: transformer/loss.py:34-48 (CE half gets a dummy)
        dummy_ce = torch.zeros(B, S, V)
        ctc_loss, _ = tloss.cal_ctc_ce_loss(lg, in_len, dummy_ce, targets, smoothing=0.0)
        if torch.isfinite(ctc_loss):
            ctc_loss.backward()
            grad = lg.grad.clone()
        else:
            ctc_loss.backward()
            grad = lg.grad.clone()
        # reference call convention B: ctcModel/loss.py:4-13 must agree with A
        loss_b = closs.cal_loss(logits, in_len, targets)
        assert torch.equal(loss_b, ctc_loss.detach()) or (not torch.isfinite(loss_b) and not torch.isfinite(ctc_loss))
        # per-utterance nll at the same boundary the reference calls (torch F.ctc_loss)
        tl = targets.ne(0).int().sum(1)
        lp = F.log_softmax(logits, dim=-1).transpose(0, 1)
        nll = F.ctc_loss(lp, targets, in_len, tl, blank=V - 1, reduction="none")
        out.update({
            f"{name}_logits": logits.numpy(), f"{name}_targets": targets.numpy(),
            f"{name}_in_len": in_len.numpy(), f"{name}_tgt_len": tl.numpy(),
            f"{name}_loss": ctc_loss.detach().numpy(), f"{name}_nll": nll.numpy(),
            f"{name}_grad": grad.numpy(),
        })
        print(f"ctc {name}: B{B} T{T} V{V} S{S} {kind}: loss={float(ctc_loss.detach()):.6f} nll={nll.tolist()}")
    np.savez_compressed(os.path.join(HERE, "ctc.npz"), **out)


def make_qua(tloss):
    g = torch.Generator().manual_seed(9)
    _number = torch.rand(6, generator=g) * 20
    number = torch.randint(1, 20, (6,), generator=g).float()
    B, T, V, S = 6, 9, 11, 3
    logits = torch.randn(B, T, V, generator=g)
    targets = torch.randint(1, V - 1, (B, S), generator=g)
    in_len = torch.full((B,), T, dtype=torch.int32)
    ce_logits = torch.randn(B, S, V, generator=g)
    qua, ctc, ce = tloss.cal_ctc_qua_ce_loss(logits, in_len, _number, number, ce_logits, targets, smoothing=0.1)
    np.savez_compressed(os.path.join(HERE, "qua.npz"), _number=_number.numpy(), number=number.numpy(),
                        logits=logits.numpy(), targets=targets.numpy(), in_len=in_len.numpy(),
                        ce_logits=ce_logits.numpy(), qua=qua.numpy(), ctc=ctc.numpy(), ce=ce.numpy())
    print(f"qua: {float(qua):.6f} ctc {float(ctc):.6f} ce {float(ce):.6f}")


def make_mha(attention, uutils):
    out = {}
    torch.manual_seed(512)
    g = torch.Generator().manual_seed(512)
    d_model, n_head, d_k = 32, 2, 64
    mha = attention.MultiheadAttention(d_model, n_head, d_k, d_k, dropout=0.1).eval()
    # make biases / layer-norm affine non-trivial
    with torch.no_grad():
        for p in mha.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    for k, v in mha.state_dict().items():
        out["w_" + k] = v.numpy()
    cases = {
        "self_pad": (3, 13, 13, "pad"),       # encoder self-attention, key padding (utils.py:157-165)
        "self_causal": (2, 9, 9, "causal"),   # decoder self-attention (utils.py:136-144 | key pad)
        "cross": (2, 5, 17, "pad"),           # decoder cross-attention
        "nomask": (2, 7, 7, "none"),
    }
    for name, (B, Lq, Lk, kind) in cases.items():
        q = torch.randn(B, Lq, d_model, generator=g)
        kv = q if Lq == Lk and name != "cross" else torch.randn(B, Lk, d_model, generator=g)
        kv_len = torch.randint(max(1, Lk // 2), Lk + 1, (B,), generator=g)
        kv_len[0] = Lk
        if kind == "pad":
            mask = uutils.get_attn_pad_mask(kv_len, Lq)
        elif kind == "causal":
            seq = (torch.arange(Lk)[None, :] < kv_len[:, None]).long()
            mask = (uutils.get_attn_key_pad_mask(seq, seq, 0).to(torch.uint8)
                    + uutils.get_subsequent_mask(seq)).gt(0)
        else:
            mask = None
        qq = q.clone().requires_grad_(True)
        kk = kv.clone().requires_grad_(True)
        y, attn = mha(qq, kk, kk, mask=mask)
        g_y = torch.randn(y.shape, generator=g)
        grads = torch.autograd.grad((y * g_y).sum(), [qq, kk] + [p for p in mha.parameters()])
        out.update({f"{name}_q": q.numpy(), f"{name}_kv": kv.numpy(), f"{name}_kv_len": kv_len.numpy(),
                    f"{name}_mask": (mask.numpy().astype(np.uint8) if mask is not None else np.zeros((0,), np.uint8)),
                    f"{name}_y": y.detach().numpy(), f"{name}_attn": attn.detach().numpy(),
                    f"{name}_g_y": g_y.numpy(), f"{name}_g_q": grads[0].numpy(), f"{name}_g_kv": grads[1].numpy()})
        for (pn, _), gp in zip(mha.named_parameters(), grads[2:]):
            out[f"{name}_gw_{pn}"] = gp.numpy()
        print(f"mha {name}: B{B} Lq{Lq} Lk{Lk} {kind}")
    np.savez_compressed(os.path.join(HERE, "mha.npz"), **out)


def make_masks(uutils):
    lens = torch.tensor([5, 3, 1, 4])
    seq = torch.tensor([[3, 4, 5, 0, 0], [1, 0, 0, 0, 0], [9, 9, 9, 9, 9]])
    np.savez_compressed(
        os.path.join(HERE, "masks.npz"),
        lens=lens.numpy(), seq=seq.numpy(),
        sequence_mask=uutils.sequence_mask(lens).numpy(),
        sequence_mask_7=uutils.sequence_mask(lens, 7).numpy(),
        attn_pad_mask=uutils.get_attn_pad_mask(lens, 3).numpy(),
        subsequent_mask=uutils.get_subsequent_mask(seq).numpy(),
        key_pad_mask=uutils.get_attn_key_pad_mask(seq, seq, 0).numpy())
    print("masks: done")


def make_cif_model(cif_model, tloss):
    """A small CIF_Model (reference classes, reference forward, reference losses, autograd)
    on config-1-shaped inputs: 8 utterances x 167 LFR frames x 320 (500 raw frames, m=4 n=3)."""
    from transformer.conv_encoder import Conv2dSubsample
    from transformer.encoder import Encoder
    from transformer.attentionAssigner import Attention_Assigner
    from transformer.decoder import Decoder_CIF
    torch.manual_seed(2026)
    d_model, vocab = 64, 100
    model = cif_model.CIF_Model(Conv2dSubsample(d_input=320, d_model=d_model, n_layers=3),
                                Encoder(d_input=d_model, n_layers=1, n_head=2, d_model=d_model, d_inner=128, dropout=0.1),
                                Attention_Assigner(d_input=d_model, d_hidden=d_model, w_context=3, n_layers=3),
                                Decoder_CIF(sos_id=2, n_tgt_vocab=vocab, n_layers=1, n_head=2, d_model=d_model,
                                            d_inner=128, dropout=0.1)).eval()
    g = torch.Generator().manual_seed(1235)
    B, T, S = 8, 167, 14
    feats = torch.randn(B, T, 320, generator=g)
    lens = torch.tensor([167, 167, 160, 151, 140, 133, 120, 101])
    feats = feats * (torch.arange(T)[None, :, None] < lens[:, None, None]).float()
    targets = torch.randint(4, vocab - 1, (B, S), generator=g)
    tl = torch.tensor([14, 13, 14, 12, 11, 12, 10, 9])
    targets = targets * (torch.arange(S)[None, :] < tl[:, None]).long()
    torch.manual_seed(77)                      # fixes torch.rand(B) inside forward (cif_model.py:47)
    ctc_logits, len_ctc, _num, num, logits = model(feats, lens, targets)
    qua, ctc, ce = tloss.cal_ctc_qua_ce_loss(ctc_logits, len_ctc, _num, num, logits, targets, smoothing=0.1)
    loss = 0.001 * qua + ctc + ce
    loss.backward()
    out = {"feats": feats.numpy(), "lens": lens.numpy(), "targets": targets.numpy(), "rand_seed": np.int64(77),
           "ctc_logits": ctc_logits.detach().numpy(), "len_ctc": len_ctc.numpy(), "_num": _num.detach().numpy(),
           "num": num.numpy(), "logits": logits.detach().numpy(), "qua": qua.detach().numpy(),
           "ctc": ctc.detach().numpy(), "ce": ce.detach().numpy()}
    for k, v in model.state_dict().items():
        if not k.endswith(".pe"):              # the sinusoid tables are deterministic, not stored
            out["sd:" + k] = v.numpy()
    for k, p in model.named_parameters():
        if k in ("ctc_fc.weight", "assigner.linear.weight", "decoder.tgt_word_prj.weight", "encoder.linear_in.weight",
                 "conv_encoder.affine.bias", "encoder.layer_stack.0.slf_attn.w_qs.weight"):
            out["grad:" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "cif_model.npz"), **out)
    print("cif_model: logits", tuple(logits.shape), "ctc_logits", tuple(ctc_logits.shape),
          "losses", float(qua), float(ctc), float(ce), "params", sum(p.numel() for p in model.parameters()))


def make_transformer_models(tloss):
    """The encoder-decoder shells of BASELINE configs 2 and 3 built from the reference's own classes and run
    through their forward, the solver's loss call and autograd (transformer.py:21-35,135-153; solver.py:23-32,83-88).
    Needs two of the documented shims: `decoder.pad_list` -> the padded tensor (decoder.py:54-56 takes the
    tuple utils.pad_list returns), and plain `Transformer` is built as Transformer(encoder, decoder) because
    `create_model` recurses (transformer.py:72)."""
    import transformer.decoder as rdec
    import transformer.transformer as rtr
    import utils.utils as uu
    from transformer.conv_encoder import Conv2dSubsample
    from transformer.encoder import Encoder
    rdec.pad_list = lambda xs, pad_value, max_len=None: uu.pad_list(xs, pad_value, max_len)[0]      # shim 2
    d_model, vocab = 64, 100
    out = {}

    def pack(prefix, model, named_grads):
        for k, v in model.state_dict().items():
            if not k.endswith(".pe"):
                out[prefix + "sd:" + k] = v.numpy()
        for k, p_ in model.named_parameters():
            if k in named_grads:
                out[prefix + "grad:" + k] = p_.grad.numpy()

    # ---- config 3: Transformer(Encoder(320, ...), Decoder(...)), CE loss --------------------------------------
    torch.manual_seed(2027)
    enc = Encoder(d_input=320, n_layers=2, n_head=2, d_model=d_model, d_inner=128, dropout=0.1)
    dec = rdec.Decoder(sos_id=2, eos_id=3, n_tgt_vocab=vocab, n_layers=2, n_head=2, d_model=d_model, d_inner=128, dropout=0.1)
    model = rtr.Transformer(enc, dec).eval()                       # shim 4
    g = torch.Generator().manual_seed(1237)
    B, T, S = 5, 50, 9
    feats = torch.randn(B, T, 320, generator=g)
    lens = torch.tensor([50, 50, 44, 37, 29])
    targets = torch.randint(4, vocab - 1, (B, S), generator=g)
    tl = torch.tensor([9, 7, 8, 5, 3])
    targets = targets * (torch.arange(S)[None, :] < tl[:, None]).long()
    logits, targets_eos = model(feats, lens, targets)
    ce = tloss.cal_ce_loss(logits, targets_eos, smoothing=0.1)
    ce.backward()
    out.update({"t:feats": feats.numpy(), "t:lens": lens.numpy(), "t:targets": targets.numpy(), "t:logits": logits.detach().numpy(),
                "t:targets_eos": targets_eos.numpy(), "t:ce": ce.detach().numpy()})
    pack("t:", model, ("encoder.linear_in.weight", "decoder.tgt_word_prj.weight", "decoder.tgt_word_emb.weight",
                       "decoder.layer_stack.0.enc_attn.w_ks.weight", "decoder.layer_stack.1.slf_attn.fc.weight",
                       "encoder.layer_stack.0.pos_ffn.w_1.bias"))
    print("transformer: logits", tuple(logits.shape), "ce", float(ce))

    # ---- config 2's model: Conv_CTC_Transformer, CTC (targets + <eos>) + CE -----------------------------------
    torch.manual_seed(2028)
    conv = Conv2dSubsample(d_input=320, d_model=d_model, n_layers=3)
    enc = Encoder(d_input=d_model, n_layers=1, n_head=2, d_model=d_model, d_inner=128, dropout=0.1)
    dec = rdec.Decoder(sos_id=2, eos_id=3, n_tgt_vocab=vocab, n_layers=1, n_head=2, d_model=d_model, d_inner=128, dropout=0.1)
    model = rtr.Conv_CTC_Transformer(conv, enc, dec).eval()
    g = torch.Generator().manual_seed(1238)
    B, T, S = 6, 167, 12
    feats = torch.randn(B, T, 320, generator=g)
    lens = torch.tensor([167, 167, 158, 149, 131, 120])
    feats = feats * (torch.arange(T)[None, :, None] < lens[:, None, None]).float()
    targets = torch.randint(4, vocab - 1, (B, S), generator=g)
    tl = torch.tensor([12, 11, 12, 9, 8, 6])
    targets = targets * (torch.arange(S)[None, :] < tl[:, None]).long()
    ctc_logits, len_ctc, logits, targets_eos = model(feats, lens, targets)
    ctc, ce = tloss.cal_ctc_ce_loss(ctc_logits, len_ctc, logits, targets_eos, smoothing=0.1)
    (ctc + ce).backward()
    out.update({"c:feats": feats.numpy(), "c:lens": lens.numpy(), "c:targets": targets.numpy(),
                "c:ctc_logits": ctc_logits.detach().numpy(), "c:len_ctc": len_ctc.numpy(), "c:logits": logits.detach().numpy(),
                "c:targets_eos": targets_eos.numpy(), "c:ctc": ctc.detach().numpy(), "c:ce": ce.detach().numpy()})
    pack("c:", model, ("ctc_fc.weight", "conv_encoder.affine.weight", "decoder.tgt_word_prj.weight",
                       "decoder.layer_stack.0.enc_attn.w_qs.weight", "encoder.layer_stack.0.slf_attn.w_vs.bias"))
    np.savez_compressed(os.path.join(HERE, "transformer_models.npz"), **out)
    print("conv_ctc_transformer: logits", tuple(logits.shape), "ctc_logits", tuple(ctc_logits.shape), "losses", float(ctc), float(ce))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "transformer_models":      # only the round-2 addition
        cif_model, tloss, closs, attention, uutils = _import_reference()
        torch.set_num_threads(1)
        make_transformer_models(tloss)
        return
    cif_model, tloss, closs, attention, uutils = _import_reference()
    torch.set_num_threads(1)
    make_cif(cif_model)
    make_cif_glue()
    make_assigner_tail(uutils)
    make_lfr()
    make_spec_aug(uutils)
    make_ctc(tloss, closs)
    make_qua(tloss)
    make_mha(attention, uutils)
    make_masks(uutils)
    make_cif_model(cif_model, tloss)
    make_transformer_models(tloss)


if __name__ == "__main__":
    main()
