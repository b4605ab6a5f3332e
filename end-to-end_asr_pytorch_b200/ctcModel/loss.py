"""Drop-in for /root/reference/src/ctcModel/loss.py (cal_loss, :4-13)."""
from ..ops import ctc_loss as _ctc_loss, ctc_fc_loss as _ctc_fc_loss, ProjectedLogits


def cal_loss(logits, len_logits, gold, smoothing=0.0):
    """Mean CTC loss of `logits` [B,T,V] against 0-padded `gold` [B,S]; blank is
    the last class.  `smoothing` is accepted and ignored, as in the reference.
    `logits` may be an ops.ProjectedLogits (the vocabulary projection fused with the loss)."""
    if isinstance(logits, ProjectedLogits):
        return _ctc_fc_loss(logits.hidden, logits.weight, len_logits, gold, blank=logits.size(-1) - 1)
    return _ctc_loss(logits, len_logits, gold, blank=logits.size(-1) - 1)
