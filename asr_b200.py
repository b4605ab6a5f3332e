"""Importable alias of the package directory `end-to-end_asr_pytorch_b200`
(whose name is not a Python identifier): `import asr_b200`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("end-to-end_asr_pytorch_b200")
sys.modules[__name__] = _pkg
