"""Drop-in for /root/reference/src/ctcModel/loss.py (cal_loss, :4-13)."""
from ..ops import ctc_loss as _ctc_loss


def cal_loss(logits, len_logits, gold, smoothing=0.0):
    """Mean CTC loss of `logits` [B,T,V] against 0-padded `gold` [B,S]; blank is
    the last class.  `smoothing` is accepted and ignored, as in the reference."""
    return _ctc_loss(logits, len_logits, gold, blank=logits.size(-1) - 1)
