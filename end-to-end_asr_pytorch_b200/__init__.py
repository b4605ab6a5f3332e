"""B200-native (sm_100a) acoustic-model training hot path of
eastonYi/end-to-end_asr_pytorch: CIF integrate-and-fire, CTC loss and
multi-head attention behind the reference's own Python API.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/asr_sm100.h)
  _lib.py          ctypes binding (loads csrc/libasr_sm100.so; no CPU fallback)
  ops.py           torch.autograd.Function wrappers (cif, ctc_loss, mha core)
  transformer/     drop-in modules under the reference's import names
  ctcModel/        (attention.py, loss.py, cif_model.py)
  utils/utils.py   attention-mask builders (reference: src/utils/utils.py:125-165)
  patch.py         installs the drop-ins into an imported reference tree
  dp.py            data-parallel gradient all-reduce (one process per GPU; csrc/allreduce.cu over NVLink peer memory, or NCCL)

The directory name is the one the build contract asks for and is not a valid
Python identifier; import it with importlib.import_module(
"end-to-end_asr_pytorch_b200") or through the alias module `asr_b200` at the
repository root.
"""
from . import _lib  # noqa: F401
from .ops import cif, cif_label_len, ctc_loss, ctc_fc_loss  # noqa: F401

__all__ = ["cif", "cif_label_len", "ctc_loss", "ctc_fc_loss"]
