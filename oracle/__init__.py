"""CPU oracle for the acoustic-model training hot path (TEST INFRASTRUCTURE).

This package is a CPU restatement of the reference algorithms on the hot path
(SURVEY.md section 8a): CIF integrate-and-fire, CTC loss + gradient, and the
multi-head attention block.  It exists only to CHECK the CUDA path:

  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
    `--impl reference` legs may import it;
  * nothing under `end-to-end_asr_pytorch_b200/` imports it, and the product
    path raises if the CUDA extension is missing (no CPU fallback).

Parity status: PINNED.  The reference ships no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, executed in the build container by `tests/golden/make_golden.py`
(reference modules imported read-only from /root/reference/src) and committed
as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function
here against those vectors (bit-exact for fire positions and fp32 CIF outputs).

The CTC arithmetic itself lives in a third-party dependency of the reference,
`torch.nn.functional.ctc_loss` (ATen LossCTC.cpp; the reference pins only
"PyTorch 1.5" in README.md:8, torch 2.11.0 is what is installed here).  The
oracle restates the published alpha-beta algorithm in float64 and is anchored
on the reference's call sites (`src/transformer/loss.py:34-48`,
`src/ctcModel/loss.py:4-13`: blank = V-1, 0-padded 2-D int64 targets, mean
reduction, zero_infinity=False) through the golden vectors above.
"""
from .cif_oracle import cif_forward, cif_backward, cif_schedule, cif_scale_alphas  # noqa: F401
from .cif_oracle import assigner_tail_forward, assigner_tail_backward, build_lfr_features, spec_aug_apply  # noqa: F401
from .ctc_oracle import ctc_loss_and_grad  # noqa: F401
from .mha_oracle import mha_core_forward, mha_core_backward, mha_module_forward  # noqa: F401
from .mask_oracle import (sequence_mask, get_attn_pad_mask, get_subsequent_mask,  # noqa: F401
                          get_attn_key_pad_mask)
