"""Import the reference's own modules (TEST INFRASTRUCTURE - see oracle/__init__.py and oracle/build_ref.py).

`load()` puts the reference's `src` directory on sys.path - /root/reference/src in the build container, the staged
byte-for-byte copy baseline/_ref/src on the GPU box - applies the documented shims and returns the
modules of the hot path.  Nothing under end-to-end_asr_pytorch_b200/ may import this file.

Shims (never edits of the reference):
  1. no CUDA device: `torch.Tensor.cuda` -> identity (hard-coded .cuda() at cif_model.py:47,61,62,69,76,100)
  2. `transformer.decoder.pad_list` -> the padded tensor of utils.pad_list's (tensor, lengths) tuple (decoder.py:54-56)
  3. `utils.utils.get_non_pad_mask` defined before ctcModel.encoder is imported (ctcModel/encoder.py:5)
  4. plain `Transformer` is built as Transformer(encoder, decoder) (create_model recurses, transformer.py:72)
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = ("/root/reference/src", os.path.join(os.path.dirname(HERE), "baseline", "_ref", "src"))


def ref_src():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "transformer", "cif_model.py")):
            return c
    return None


def available():
    return ref_src() is not None


_loaded = None


def load(cpu_shim=None):
    """-> namespace with cif_model, loss, closs, attention, cattention, utils, encoder, decoder, conv_encoder,
    assigner, transformer, solver, src (the directory).  cpu_shim: None = apply shim 1 only when torch sees no GPU."""
    global _loaded
    import torch
    if cpu_shim is None:
        cpu_shim = not torch.cuda.is_available()
    if cpu_shim and getattr(torch.Tensor.cuda, "__name__", "") != "_identity_cuda":
        def _identity_cuda(self, *a, **k):
            return self
        torch.Tensor.cuda = _identity_cuda                          # shim 1
    if _loaded is not None:
        return _loaded
    src = ref_src()
    if src is None:
        raise ImportError("reference sources not found (neither /root/reference/src nor baseline/_ref/src; run oracle/build_ref.py "
                          "in the build container)")
    if src not in sys.path:
        sys.path.insert(0, src)
    uutils = importlib.import_module("utils.utils")
    if not hasattr(uutils, "get_non_pad_mask"):                     # shim 3
        uutils.get_non_pad_mask = lambda x, input_lengths=None, pad_idx=None: \
            uutils.sequence_mask(input_lengths, x.size(1)).unsqueeze(-1)
    ns = types.SimpleNamespace(src=src, utils=uutils)
    ns.cif_model = importlib.import_module("transformer.cif_model")
    ns.loss = importlib.import_module("transformer.loss")
    ns.closs = importlib.import_module("ctcModel.loss")
    ns.attention = importlib.import_module("transformer.attention")
    ns.cattention = importlib.import_module("ctcModel.attention")
    ns.encoder = importlib.import_module("transformer.encoder")
    ns.decoder = importlib.import_module("transformer.decoder")
    ns.conv_encoder = importlib.import_module("transformer.conv_encoder")
    ns.assigner = importlib.import_module("transformer.attentionAssigner")
    ns.transformer = importlib.import_module("transformer.transformer")
    ns.solver = importlib.import_module("transformer.solver")
    if ns.decoder.pad_list is uutils.pad_list:                      # shim 2
        ns.decoder.pad_list = lambda xs, pad_value, max_len=None: uutils.pad_list(xs, pad_value, max_len)[0]
    _loaded = ns
    return ns


def joint_hot_path_step(ns, hidden, alphas_raw, logits, len_logits, targets, noise, g_fired=None, threshold=0.95):
    """One pass of the hot path through the REFERENCE's own functions, on whatever device the tensors live:
    the scaling lines of CIF_Model.forward (cif_model.py:43-48, re-executed verbatim with `noise` standing for
    torch.rand(B)), CIF_Model.cif (:57-106), the quantity term and cal_ctc_ce_loss's CTC half (loss.py:39-43,55),
    then autograd backward of (ctc + 0.001 qua + <g_fired, fired>) - the work solver.py:146-157 does around the encoder."""
    import torch
    hidden = hidden.detach().clone().requires_grad_(True)
    alphas_raw = alphas_raw.detach().clone().requires_grad_(True)
    logits = logits.detach().clone().requires_grad_(True)
    _num = alphas_raw.sum(-1)
    num = (targets > 0).float().sum(-1)
    num_noise = num + noise - 0.5
    alphas = alphas_raw * (num_noise / _num)[:, None].repeat(1, alphas_raw.size(1))
    fired = ns.cif_model.CIF_Model.cif(None, hidden, alphas, threshold)
    qua = torch.pow(_num - num, 2).mean()
    # cal_ctc_ce_loss also wants CE logits; its CTC half is these three lines (loss.py:39-43), called through the
    # reference's own ctcModel.loss.cal_loss, which is exactly that half
    ctc = ns.closs.cal_loss(logits, len_logits, targets)
    if g_fired is None:
        g_fired = torch.ones_like(fired)
    total = ctc + 0.001 * qua + (fired * g_fired[:, :fired.size(1)]).sum()
    total.backward()
    return {"ctc": ctc.detach(), "qua": qua.detach(), "fired": fired.detach(),
            "g_hidden": hidden.grad, "g_alphas": alphas_raw.grad, "g_logits": logits.grad}
