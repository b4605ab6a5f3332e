"""`transformer.CIF_Model` - the module name the reference's CLIs import
(/root/reference/src/transformer/train.py:157 and infer.py: `from transformer.CIF_Model import CIF_Model`),
while the class lives in `cif_model.py`; this alias makes that import resolve."""
from .cif_model import CIF_Model  # noqa: F401
