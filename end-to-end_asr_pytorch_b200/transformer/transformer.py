"""Drop-in for /root/reference/src/transformer/transformer.py: the `Transformer`, `CTC_Transformer` and
`Conv_CTC_Transformer` model shells (callers of the attention / CTC hot path; BASELINE configs 2 and 3).

Same constructors, `forward` signatures and return values, and - through the encoder / decoder / conv
front end of this package - the same state_dict keys (`encoder.*`, `decoder.*`, `conv_encoder.*`,
`ctc_fc.weight`), so checkpoints interchange.  Training forwards only: `recognize` / `batch_recognize`
delegate to decoder methods that are out of scope here (beam search) and raise if the
decoder does not provide them.

Two defects of the reference as checked in are not reproduced: `Transformer.create_model` recursing
into itself with the wrong arguments (transformer.py:72) - it builds `cls(encoder, decoder)` here - and
`Conv_CTC_Transformer.load_model_from_package` passing a keyword `create_model` does not take (:217).
"""
import importlib

import torch
import torch.nn as nn

from .module import Linear


def _sibling(name):
    pkg_root = __name__.rsplit(".", 2)[0]
    return importlib.import_module(pkg_root + "." + name)


class Transformer(nn.Module):
    """Encoder-decoder with attention only (reference transformer.py:7-98; config 3)."""

    def __init__(self, encoder, decoder, spec_aug_cfg=None):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.spec_aug_cfg = spec_aug_cfg
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def _augment(self, features, len_features):
        if self.spec_aug_cfg:
            features, len_features = _sibling("utils.utils").spec_aug(features, len_features, self.spec_aug_cfg)
        return features, len_features

    def forward(self, features, len_features, padded_target):
        """features N x Ti x D, len_features N, padded_target N x To -> (logits, targets_eos)."""
        features, len_features = self._augment(features, len_features)
        encoder_padded_outputs = self.encoder(features, len_features)
        logits, targets_eos = self.decoder(padded_target, encoder_padded_outputs, len_features)
        return logits, targets_eos

    def recognize(self, input, input_length, char_list, args):
        encoder_outputs = self.encoder(input.unsqueeze(0), input_length)
        return self.decoder.recognize_beam(encoder_outputs[0], char_list, args)

    @classmethod
    def _encoder_decoder(cls, args, d_input):
        Encoder = _sibling("transformer.encoder").Encoder
        Decoder = _sibling("transformer.decoder").Decoder
        encoder = Encoder(d_input=d_input, n_layers=args.n_layers_enc, n_head=args.n_head, d_model=args.d_model,
                          d_inner=args.d_inner, dropout=args.dropout)
        decoder = Decoder(sos_id=args.sos_id, eos_id=args.eos_id, n_tgt_vocab=args.vocab_size, n_layers=args.n_layers_dec,
                          n_head=args.n_head, d_model=args.d_model, d_inner=args.d_inner, dropout=args.dropout)
        return encoder, decoder

    @classmethod
    def create_model(cls, args):
        encoder, decoder = cls._encoder_decoder(args, args.d_input * args.LFR_m)
        return cls(encoder, decoder)

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


class CTC_Transformer(Transformer):
    """+ a CTC head on the encoder (reference transformer.py:101-127)."""

    def __init__(self, encoder, decoder, spec_aug_cfg=None):
        super().__init__(encoder, decoder, spec_aug_cfg)
        self.ctc_fc = Linear(encoder.d_output, decoder.d_output, bias=False)

    def forward(self, features, len_features, padded_target):
        """-> (ctc_pred_len, ctc_pred, pred), pred being the decoder's (logits, targets_eos) pair as in the reference."""
        features, len_features = self._augment(features, len_features)
        encoder_padded_outputs = self.encoder(features, len_features)
        ctc_pred = self.ctc_fc(encoder_padded_outputs)
        pred = self.decoder(padded_target, encoder_padded_outputs, len_features)
        return len_features, ctc_pred, pred


class Conv_CTC_Transformer(CTC_Transformer):
    """Conv2d sub-sampling front end + encoder + CTC head + decoder (reference transformer.py:130-222; the model of
    BASELINE config 2)."""

    def __init__(self, conv_encoder, encoder, decoder, spec_aug_cfg=None):
        super().__init__(encoder, decoder, spec_aug_cfg)
        self.conv_encoder = conv_encoder

    def forward(self, features, len_features, targets, spec_aug_cfg=False):
        """-> (ctc_logits, len_ctc_logits, logits, targets_eos) for cal_ctc_ce_loss (solver.py:83-88)."""
        features, len_features = self._augment(features, len_features)
        conv_outputs, len_sequence = self.conv_encoder(features, len_features)
        encoder_outputs = self.encoder(conv_outputs, len_sequence)
        ctc_logits = self.ctc_fc(encoder_outputs)
        logits, targets_eos = self.decoder(targets, encoder_outputs, len_sequence)
        return ctc_logits, len_sequence, logits, targets_eos

    def recognize(self, feature, len_feature, char_list, args):
        conv_outputs, len_sequence = self.conv_encoder(feature, len_feature)
        encoder_outputs = self.encoder(conv_outputs, len_sequence)
        return self.decoder.recognize(encoder_outputs[0], char_list, args)

    def batch_recognize(self, features, len_features, beam_size):
        conv_outputs, len_sequences = self.conv_encoder(features, len_features)
        encoder_outputs = self.encoder(conv_outputs, len_sequences)
        return self.decoder.batch_decode(encoder_outputs, len_sequences, beam_size)

    @classmethod
    def create_model(cls, args):
        Conv2dSubsample = _sibling("transformer.conv_encoder").Conv2dSubsample
        conv_encoder = Conv2dSubsample(d_input=args.d_input * args.LFR_m, d_model=args.d_model, n_layers=args.n_conv_layers)
        encoder, decoder = cls._encoder_decoder(args, args.d_model)
        return cls(conv_encoder, encoder, decoder, spec_aug_cfg=args.spec_aug_cfg)

    @classmethod
    def load_model_from_package(cls, package, args):
        model = cls.create_model(args)
        model.load_state_dict(package['state_dict'])
        return model
