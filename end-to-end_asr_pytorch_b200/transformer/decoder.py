"""Caller: the CIF decoder, training forward only (decoding / beam search are out of
scope).  Names follow Decoder_CIF in
/root/reference/src/transformer/decoder.py:327-396 (`tgt_word_emb`, `layer_stack`,
`input_affine`, `tgt_word_prj`)."""
import torch
import torch.nn as nn

from .module import Linear

from .encoder import EncoderLayer
from .module import PositionalEncoding


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
