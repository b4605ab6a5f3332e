"""Drop-in for /root/reference/src/ctcModel/attention.py: the same block as
transformer/attention.py with the reference's other constructor order
`MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=0.1)` (ctcModel/attention.py:9)."""
from ..transformer.attention import MultiheadAttention as _Base


class MultiHeadAttention(_Base):
    ''' Multi-Head Attention module '''

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1, **kw):
        super().__init__(d_model, n_head, d_k, d_v, dropout=dropout, **kw)
