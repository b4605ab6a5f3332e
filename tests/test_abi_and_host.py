"""CPU-side checks: the C-ABI library loads and exports every symbol declared in
include/asr_sm100.h, the ctypes table mirrors the header, host-side helpers work
without a GPU, and the product path refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import pkg, ROOT

HEADER = os.path.join(ROOT, "include", "asr_sm100.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = _declared_symbols()
    for name in ("asr_cif_fwd_f32", "asr_cif_bwd_f32", "asr_ctc_fwd_bwd_f32", "asr_mha_fwd_bf16",
                 "asr_mha_bwd_bf16", "asr_last_error"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    lib = pkg("_lib")
    handle = ctypes.CDLL(lib.LIB_PATH) if os.path.exists(lib.LIB_PATH) else lib.lib()
    for name in _declared_symbols():
        assert hasattr(handle, name), "libasr_sm100.so does not export %s" % name
    assert set(lib.SIGNATURES) == set(_declared_symbols())
    assert lib.lib().asr_abi_version() == 1


def test_workspace_queries_and_options_run_without_gpu():
    lib = pkg("_lib")
    L = lib.lib()
    assert L.asr_cif_bwd_workspace_bytes(4, 100) == 4 * 100 * 4
    assert L.asr_ctc_workspace_bytes(2, 10, 50, 3) >= 2 * 10 * 8 * 4
    lib.set_option("cif_fwd_variant", 2)
    assert lib.get_option("cif_fwd_variant") == 2
    lib.set_option("cif_fwd_variant", 0)
    with pytest.raises(RuntimeError):
        lib.set_option("no_such_option", 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    ops = pkg("ops")
    with pytest.raises(RuntimeError):
        ops.cif(torch.zeros(1, 4, 8), torch.zeros(1, 4), 0.95)
    with pytest.raises(RuntimeError):
        ops.ctc_loss(torch.zeros(1, 4, 8), torch.tensor([4]), torch.ones(1, 2, dtype=torch.long))
    L = pkg("_lib").lib()
    assert L.asr_device_ok() != 0          # no device here: every compute entry point refuses to run
    assert pkg("_lib").last_error()


def test_mask_builders_match_reference_goldens():
    u = pkg("utils.utils")
    g = load_golden("masks")
    lens, seq = torch.as_tensor(g["lens"]), torch.as_tensor(g["seq"])
    np.testing.assert_array_equal(u.sequence_mask(lens).numpy(), g["sequence_mask"])
    np.testing.assert_array_equal(u.sequence_mask(lens, 7).numpy(), g["sequence_mask_7"])
    np.testing.assert_array_equal(u.get_attn_pad_mask(lens, 3).numpy(), g["attn_pad_mask"])
    np.testing.assert_array_equal(u.get_subsequent_mask(seq).numpy(), g["subsequent_mask"])
    np.testing.assert_array_equal(u.get_attn_key_pad_mask(seq, seq, 0).numpy(), g["key_pad_mask"])


def test_ce_loss_matches_reference_golden():
    tl = pkg("transformer.loss")
    g = load_golden("qua")
    ce = tl.cal_ce_loss(torch.as_tensor(g["ce_logits"]), torch.as_tensor(g["targets"]), smoothing=0.1)
    np.testing.assert_allclose(ce.numpy(), g["ce"], rtol=1e-6)


def test_product_code_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "end-to-end_asr_pytorch_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M):
                    bad.append(f)
    assert not bad, bad
