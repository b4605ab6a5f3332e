"""Callers of the attention hot path: the teacher-forced decoders, training forward only
(decoding / beam search are out of scope).  Names follow
/root/reference/src/transformer/decoder.py: `Decoder` (:13-96: `tgt_word_emb`, `layer_stack.N.
{slf_attn,enc_attn,pos_ffn}`, `tgt_word_prj`), `DecoderLayer` (:617-636) and `Decoder_CIF`
(:327-396: `tgt_word_emb`, `layer_stack`, `input_affine`, `tgt_word_prj`)."""
import torch
import torch.nn as nn

from .module import Linear

from .attention import MultiheadAttention
from .encoder import EncoderLayer
from .module import PositionalEncoding, PositionwiseFeedForward
from ..utils.utils import get_attn_key_pad_mask, get_subsequent_mask


class DecoderLayer(nn.Module):
    """Self-attention, encoder-decoder attention and position-wise FFN, each followed by the
    non-pad mask (reference decoder.py:617-636)."""

    def __init__(self, d_model, d_inner, n_head, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiheadAttention(d_model, n_head, dropout=dropout, return_attn=False)
        self.enc_attn = MultiheadAttention(d_model, n_head, dropout=dropout, return_attn=False)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, dropout=dropout)

    def forward(self, dec_input, enc_output, non_pad_mask=None, slf_attn_mask=None, dec_enc_attn_mask=None,
                slf_kv_len=None, slf_causal=False, enc_kv_len=None):
        """The masks are the reference's dense [B,Lq,Lk] tensors and / or their structured forms
        (`slf_kv_len` + `slf_causal` for decoder.py:74-78, `enc_kv_len` for get_attn_pad_mask)."""
        out, _ = self.slf_attn(dec_input, dec_input, dec_input, mask=slf_attn_mask, kv_len=slf_kv_len, causal=slf_causal,
                               out_scale=non_pad_mask)
        out, _ = self.enc_attn(out, enc_output, enc_output, mask=dec_enc_attn_mask, kv_len=enc_kv_len, out_scale=non_pad_mask)
        return self.pos_ffn(out, out_scale=non_pad_mask)


class Decoder(nn.Module):
    """Teacher-forced Transformer decoder (reference decoder.py:13-96)."""

    def __init__(self, sos_id, eos_id, n_tgt_vocab, n_layers, n_head, d_model, d_inner, dropout=0.1):
        super().__init__()
        self.sos_id, self.eos_id, self.n_tgt_vocab = sos_id, eos_id, n_tgt_vocab
        self.d_word_vec = self.d_model = d_model
        self.n_layers, self.n_head, self.d_inner = n_layers, n_head, d_inner
        self.d_output = n_tgt_vocab
        self.tgt_word_emb = nn.Embedding(n_tgt_vocab, d_model)
        self.positional_encoding = PositionalEncoding(d_model)
        self.dropout = nn.Dropout(dropout)
        self.layer_stack = nn.ModuleList([DecoderLayer(d_model, d_inner, n_head, dropout=dropout)
                                          for _ in range(n_layers)])
        self.tgt_word_prj = Linear(d_model, n_tgt_vocab, bias=False)
        nn.init.xavier_normal_(self.tgt_word_prj.weight)
        # True: size the decoder input as targets.size(1) + 1 without reading max_b(len_b) back from the device
        # (no host sync: the step can be captured in a CUDA graph).  Identical to the reference whenever one row
        # of the batch has no padding; otherwise the extra columns are padding (no loss, no gradient).
        self.assume_full_width = False

    def preprocess(self, targets):
        """(<sos> + labels, labels + <eos>), both 0-padded to max_b(len_b) + 1 (reference decoder.py:41-58, with
        pad_list's padded tensor - the reference as checked in takes the (tensor, lengths) tuple, SURVEY.md shim 2).
        Labels are the non-zero entries of each row in order, so this is a stable left-compaction - done on the
        device without a per-utterance Python loop; the output width needs one host read, like the reference's."""
        B, S = targets.shape
        keep = targets != 0
        n = keep.sum(1)
        width = S + 1 if self.assume_full_width else (int(n.max().item()) + 1 if B > 0 else 1)
        dest = torch.cumsum(keep.long(), 1) - 1                     # slot of every kept label inside its row
        dest = torch.where(keep, dest, torch.full_like(dest, width))  # dropped entries go to a spill column
        packed = targets.new_zeros((B, width + 1))
        packed.scatter_(1, dest, targets)
        packed = packed[:, :width]                                  # labels, left-aligned, 0-padded
        ys_out = packed.scatter(1, n.unsqueeze(1), torch.full_like(n, self.eos_id).unsqueeze(1))
        sos = torch.full((B, 1), self.sos_id, dtype=targets.dtype, device=targets.device)
        ys_in = torch.cat([sos, packed[:, :-1]], 1)
        return ys_in, ys_out

    def forward(self, targets, encoder_padded_outputs, encoder_input_lengths):
        """targets N x To (0-padded), encoder outputs N x Ti x H, their lengths N
        -> (logits N x (To'+1) x vocab, targets_eos N x (To'+1))."""
        targets_sos, targets_eos = self.preprocess(targets)
        non_pad_mask = (targets_sos > 0).unsqueeze(-1)
        if self.sos_id > 0:
            # key padding on <sos>+labels OR strictly-upper triangle (decoder.py:74-78) = (valid prefix length, causal);
            # get_attn_pad_mask(encoder lengths) (decoder.py:80) = key lengths
            slf = dict(slf_kv_len=(targets_sos > 0).sum(-1), slf_causal=True)
            slf_mask = None
        else:      # a zero <sos> id is itself masked as a key by the reference: keep its dense mask
            slf = {}
            slf_mask = (get_attn_key_pad_mask(seq_k=targets_sos, seq_q=targets_sos, pad_idx=0).to(torch.uint8)
                        + get_subsequent_mask(targets_sos)).gt(0)
        x = self.dropout(self.tgt_word_emb(targets_sos) + self.positional_encoding(targets_sos))
        for layer in self.layer_stack:
            x = layer(x, encoder_padded_outputs, non_pad_mask=non_pad_mask, slf_attn_mask=slf_mask,
                      enc_kv_len=encoder_input_lengths, **slf)
        return self.tgt_word_prj(x), targets_eos


class Decoder_CIF(nn.Module):
    def __init__(self, sos_id, n_tgt_vocab, n_layers, n_head, d_model, d_inner, dropout=0.1):
        super().__init__()
        self.sos_id, self.n_tgt_vocab = sos_id, n_tgt_vocab
        self.d_word_vec = self.d_model = d_model
        self.n_layers, self.n_head, self.d_inner = n_layers, n_head, d_inner
        self.d_output = n_tgt_vocab
        self.tgt_word_emb = nn.Embedding(n_tgt_vocab, d_model)
        self.positional_encoding = PositionalEncoding(d_model)
        self.dropout = nn.Dropout(dropout)
        self.layer_stack = nn.ModuleList([EncoderLayer(d_model, d_inner, n_head, dropout=dropout)
                                          for _ in range(n_layers)])
        self.input_affine = Linear(2 * d_model, d_model, bias=False)
        self.tgt_word_prj = Linear(2 * d_model, n_tgt_vocab, bias=False)
        nn.init.xavier_normal_(self.tgt_word_prj.weight)

    def preprocess(self, target):
        """<sos> + target shifted right, zeroed where the target is padding."""
        sos = torch.full((target.size(0), 1), self.sos_id, dtype=torch.long, device=target.device)
        return torch.cat([sos, target[:, :-1]], 1) * (target > 0).long()

    def forward(self, encoded_attentioned, target):
        """fired frames N x To x d_model, targets N x To -> logits N x To x vocab.
        The reference's self-attention mask (key padding on ys_in OR strictly-upper
        triangle, decoder.py:374-378) is handed to the kernel as (kv_len, causal)."""
        ys_in = self.preprocess(target)
        non_pad_mask = (target > 0).unsqueeze(-1)
        kv_len = (ys_in > 0).sum(-1)
        x = self.dropout(self.tgt_word_emb(ys_in) + self.positional_encoding(ys_in))
        x = self.input_affine(torch.cat([encoded_attentioned, x], -1))
        for layer in self.layer_stack:
            x = layer(x, non_pad_mask=non_pad_mask, kv_len=kv_len, causal=True)
        return self.tgt_word_prj(torch.cat([encoded_attentioned, x], -1))
