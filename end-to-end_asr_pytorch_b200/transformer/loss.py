"""Drop-in for /root/reference/src/transformer/loss.py.

Same three functions, same signatures and return values.  The CTC term runs the
fused sm_100a kernels (log-softmax folded in, gradient produced in the same
pass); the label-smoothed cross entropy is not on the hot path and
stays plain PyTorch, restated here so the module is self-contained.
"""
import torch
import torch.nn.functional as F

from ..ops import ctc_loss as _ctc_loss, ctc_fc_loss as _ctc_fc_loss, ProjectedLogits


def cal_ce_loss(logits, targets, smoothing=0.0):
    """Label-smoothed cross entropy with pad id 0 (reference loss.py:5-31)."""
    n_class = logits.size(-1)
    flat_logits = logits.reshape(-1, n_class)
    flat_targets = targets.contiguous().view(-1)
    if smoothing > 0.0:
        log_prb = F.log_softmax(flat_logits, dim=1)
        true_dist = torch.full_like(log_prb, smoothing / n_class)
        true_dist.scatter_(1, flat_targets.long().unsqueeze(1), 1.0 - smoothing)
        keep = flat_targets.ne(0)
        per_token = -(true_dist * log_prb).sum(dim=1)
        # the reference's masked_select(keep).sum() without its dynamic-size result (a host sync): same addends, zeros
        # in place of the dropped ones
        return (per_token * keep.to(per_token.dtype)).sum() / keep.long().sum()
    return F.cross_entropy(flat_logits, flat_targets, ignore_index=0, reduction='mean')


def cal_ctc_ce_loss(logits_ctc, len_logits_ctc, logits_ce, targets, smoothing=0.0):
    """(ctc_loss, ce_loss) - reference loss.py:34-48.  blank = V-1, target
    lengths = number of non-zero labels, reduction 'mean', zero_infinity off."""
    if isinstance(logits_ctc, ProjectedLogits):      # projection fused with the loss: the logits never reach HBM twice
        ctc = _ctc_fc_loss(logits_ctc.hidden, logits_ctc.weight, len_logits_ctc, targets, blank=logits_ctc.size(-1) - 1)
    else:
        ctc = _ctc_loss(logits_ctc, len_logits_ctc, targets, blank=logits_ctc.size(-1) - 1)
    ce = cal_ce_loss(logits_ce, targets, smoothing)
    return ctc, ce


def cal_ctc_qua_ce_loss(logits_ctc, len_logits_ctc, _number, number, logits_ce, targets, smoothing=0.0):
    """(qua_loss, ctc_loss, ce_loss) - reference loss.py:51-61."""
    qua = torch.pow(_number - number, 2).mean()
    ctc, ce = cal_ctc_ce_loss(logits_ctc, len_logits_ctc, logits_ce, targets, smoothing)
    return qua, ctc, ce
