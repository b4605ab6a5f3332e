"""Drop-in for /root/reference/src/transformer/cif_model.py (class CIF_Model).

Same constructor, `forward`, `cif`, `recognize`, `create_model`, `load_model`
and `serialize` signatures and the same state_dict keys (`ctc_fc.weight`, and
whatever the conv_encoder / encoder / assigner / decoder sub-modules register),
so checkpoints interchange with the reference.

What changes: `cif()` is one call into the sm_100a integrate-and-fire kernel
pair (forward + analytic backward) instead of a Python loop over frames, and the
hard-coded `.cuda()` calls follow the input's device.  The alpha scaling glue
(reference :43-48) keeps the reference's exact torch ops - including drawing the
noise from the CPU generator like `torch.rand(B).cuda()` does - so the scaled
alphas, and with them the fire positions, are bit-identical for the same seed.
"""
import importlib

import torch
import torch.nn as nn

from ..ops import cif as _cif_op, ProjectedLogits
from .module import Linear


def _sibling(name):
    """A caller-side module (conv_encoder, encoder, attentionAssigner, decoder,
    utils.utils): ours if the package has one, else the reference's, which is on
    sys.path when this file is used as a drop-in inside the reference tree."""
    pkg_root = __name__.rsplit(".", 2)[0]
    for cand in (pkg_root + "." + name, name):
        try:
            return importlib.import_module(cand)
        except ImportError:
            continue
    raise ImportError("cannot find '%s' in this package or in the reference tree" % name)


class CIF_Model(nn.Module):
    """Conv front-end -> Transformer encoder -> (ctc_fc | assigner -> CIF -> decoder)."""

    def __init__(self, conv_encoder, encoder, assigner, decoder, spec_aug_cfg=None, fused_alpha=False):
        super().__init__()
        self.conv_encoder = conv_encoder
        self.encoder = encoder
        self.assigner = assigner
        # True: the assigner tail and the scaling below run as one kernel pass (ops.cif_alpha).  Same
        # math, different fp32 summation order than torch's Linear/sum, so a weight that lands within
        # one ulp of the threshold may fire a frame earlier or later than in the reference.
        self.fused_alpha = fused_alpha
        # True: no host synchronisation inside forward, so a whole training step can be captured in a CUDA graph:
        # the noise of :47 is drawn by the device generator (the reference draws it on the host and copies it), the
        # fired tensor is sized L = targets.size(1) instead of reading max_b round(sum alphas) back (equal whenever
        # one utterance of the batch has no padding - alphas are scaled to sum to #labels +- 0.5), and the
        # more-fires-than-rows check (which the reference dies on, :100) is skipped.
        self.static_shapes = False
        # True (training mode only): forward returns an ops.ProjectedLogits in place of `ctc_logits`, and the loss
        # functions of this package run the ctc_fc projection fused with the CTC loss: the
        # [B,T,V] logits are produced, turned into their gradient in place and consumed by the backward GEMMs inside
        # one call, instead of living in HBM from forward to backward.
        self.fused_ctc_fc = False
        self.decoder = decoder
        self.spec_aug_cfg = spec_aug_cfg
        self.ctc_fc = Linear(encoder.d_output, decoder.d_output, bias=False)

        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, features, len_features, targets, threshold=0.95):
        """features N x T x D, len_features N, targets N x To (0-padded).
        Returns (ctc_logits, len_ctc_logits, _num, num, logits) like the reference."""
        if self.spec_aug_cfg:
            spec_aug = _sibling("utils.utils").spec_aug
            features, len_features = spec_aug(features, len_features, self.spec_aug_cfg)

        conv_outputs, len_sequence = self.conv_encoder(features, len_features)
        encoder_outputs = self.encoder(conv_outputs, len_sequence)

        if (getattr(self, "fused_ctc_fc", False) and self.training and encoder_outputs.is_cuda
                and encoder_outputs.dtype == torch.float32 and not torch.is_autocast_enabled("cuda")):
            ctc_logits = ProjectedLogits(encoder_outputs, self.ctc_fc.weight)
        else:
            ctc_logits = self.ctc_fc(encoder_outputs)
        len_ctc_logits = len_sequence

        # quantity (before scaling) and target-length scaling, reference :43-48
        num = (targets > 0).float().sum(-1)
        static = getattr(self, "static_shapes", False)
        noise = torch.rand(targets.size(0), device=num.device) if static else torch.rand(targets.size(0)).to(num.device)
        num_noise = num + noise - 0.5
        if getattr(self, "fused_alpha", False) and hasattr(self.assigner, "forward_scaled"):
            alpha, _num = self.assigner.forward_scaled(encoder_outputs, len_sequence, num_noise)
        else:
            alpha = self.assigner(encoder_outputs, len_sequence)
            _num = alpha.sum(-1)
            alpha = alpha * (num_noise / _num)[:, None]

        if static:
            fired = _cif_op(encoder_outputs.float(), alpha.float(), threshold, L=targets.size(1),
                            check_overflow=False).to(encoder_outputs.dtype)
        else:
            fired = self.cif(encoder_outputs, alpha, threshold=threshold)

        logits = self.decoder(fired, targets)

        return ctc_logits, len_ctc_logits, _num, num, logits

    def cif(self, hidden, alphas, threshold, log=False):
        """Integrate-and-fire (reference :57-106): [B,T,H], [B,T] -> [B,L,H] with
        L = max_b round(sum_t alphas).  `log` is accepted for signature parity.
        The kernels work in fp32 like the reference; a reduced-precision model (bf16 evaluation) is
        cast at this boundary and the fired frames go back to the activations' dtype."""
        if hidden.dtype != torch.float32 or alphas.dtype != torch.float32:
            return _cif_op(hidden.float(), alphas.float(), threshold).to(hidden.dtype)
        return _cif_op(hidden, alphas, threshold)

    def recognize(self, input, input_length, char_list, args, threshold=0.95, target_num=None):
        """Beam-search decode of one utterance (reference :108-131); only the
        CIF step runs our kernel, decoding itself is the decoder's business."""
        conv_padded_outputs, input_length = self.conv_encoder(input.unsqueeze(0), input_length)
        encoder_outputs = self.encoder(conv_padded_outputs, input_length)

        alpha = self.assigner(encoder_outputs, input_length)
        if target_num:
            _num = alpha.sum(-1)
            alpha = alpha * (target_num / _num)[:, None]

        fired = self.cif(encoder_outputs, alpha, threshold=threshold)
        return self.decoder.recognize_beam(fired, char_list, args)

    @classmethod
    def create_model(cls, args):
        Conv2dSubsample = _sibling("transformer.conv_encoder").Conv2dSubsample
        Encoder = _sibling("transformer.encoder").Encoder
        Attention_Assigner = _sibling("transformer.attentionAssigner").Attention_Assigner
        Decoder = _sibling("transformer.decoder").Decoder_CIF

        conv_encoder = Conv2dSubsample(d_input=args.d_input * args.LFR_m, d_model=args.d_model,
                                       n_layers=args.n_conv_layers)
        encoder = Encoder(d_input=args.d_model, n_layers=args.n_layers_enc, n_head=args.n_head,
                          d_model=args.d_model, d_inner=args.d_inner, dropout=args.dropout)
        assigner = Attention_Assigner(d_input=args.d_model, d_hidden=args.d_assigner_hidden,
                                      w_context=args.w_context, n_layers=args.n_assigner_layers)
        decoder = Decoder(sos_id=args.sos_id, n_tgt_vocab=args.vocab_size, n_layers=args.n_layers_dec,
                          n_head=args.n_head, d_model=args.d_model, d_inner=args.d_inner, dropout=args.dropout)
        return cls(conv_encoder, encoder, assigner, decoder, args.spec_aug_cfg)

    @classmethod
    def load_model(cls, path, args):
        model = cls.create_model(args)
        package = torch.load(path, map_location=lambda storage, loc: storage)
        model.load_state_dict(package['state_dict'])
        return model

    @staticmethod
    def serialize(model, optimizer, epoch, tr_loss=None, cv_loss=None):
        package = {'state_dict': model.state_dict(), 'optim_dict': optimizer.state_dict(), 'epoch': epoch}
        if tr_loss is not None:
            package['tr_loss'] = tr_loss
            package['cv_loss'] = cv_loss
        return package
