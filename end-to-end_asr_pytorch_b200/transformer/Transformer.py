"""`transformer.Transformer` - the module name the reference's CLIs import
(/root/reference/src/transformer/train.py:139,145,151 and infer.py:85-91: `from transformer.Transformer import
Transformer | CTC_Transformer | Conv_CTC_Transformer`), while the classes live in `transformer.py`.
On a case-sensitive file system the reference fails there with ModuleNotFoundError (headline
facts); this alias makes those imports resolve."""
from .transformer import Transformer, CTC_Transformer, Conv_CTC_Transformer  # noqa: F401
