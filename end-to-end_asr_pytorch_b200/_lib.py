"""ctypes binding of csrc/libasr_sm100.so (C ABI: include/asr_sm100.h).

There is no CPU fallback: `lib()` raises if the shared library is missing and
cannot be built, and every compute entry point returns an error on a machine
without an sm_100 device (surfaced here as RuntimeError).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASR_SM100_LIB: an alternative build of the same library (e.g. the -DASR_MHA_TRACE instrumented one of tools/mha_trace.py)
LIB_PATH = os.environ.get("ASR_SM100_LIB") or os.path.join(_HERE, "csrc", "libasr_sm100.so")

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_c_size_t = ctypes.c_size_t
_vp = ctypes.c_void_p
_c_uint64 = ctypes.c_uint64

# name -> (restype, argtypes); mirrors include/asr_sm100.h one to one
SIGNATURES = {
    "asr_abi_version": (_c_int, []),
    "asr_last_error": (ctypes.c_char_p, []),
    "asr_device_ok": (_c_int, []),
    "asr_set_option": (_c_int, [ctypes.c_char_p, _c_int]),
    "asr_get_option": (_c_int, [ctypes.c_char_p, ctypes.POINTER(_c_int)]),
    "asr_launch_count": (ctypes.c_uint64, []),
    "asr_cif_fwd_f32": (_c_int, [_vp, _vp, _c_float, _c_int, _c_int, _c_int, _c_int,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "asr_cif_fwd_hint_f32": (_c_int, [_vp, _vp, _c_float, _c_int, _c_int, _c_int, _c_int,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _vp]),
    "asr_cif_bwd_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "asr_cif_bwd_f32": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int,
                                 _vp, _vp, _vp, _c_size_t, _vp]),
    "asr_cif_alpha_fwd_f32": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp]),
    "asr_cif_alpha_bwd_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "asr_cif_alpha_bwd_f32": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int,
                                       _vp, _vp, _vp, _vp, _c_size_t, _vp]),
    "asr_lfr_f32": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "asr_spec_aug_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "asr_spec_aug_f32": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _c_size_t, _vp]),
    "asr_linear_act_bf16": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "asr_linear_residual_layernorm_bf16": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c_float, _c_int, _c_int, _c_int, _vp, _vp]),
    "asr_linear_f32": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _vp, _vp]),
    "asr_colsum": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "asr_ln_dropout_keep_prob": (_c_float, [_c_float]),
    "asr_ln_fwd": (_c_int, [_vp, _c_int, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_float, _c_float, _c_uint64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "asr_ln_bwd_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "asr_ln_bwd": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_float, _c_uint64, _vp, _vp, _vp, _c_int, _vp, _vp, _c_size_t, _vp]),
    "asr_ln_eval_bf16": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_float, _vp, _vp]),
    "asr_relu_bwd_bf16": (_c_int, [_vp, _vp, _vp, _c_size_t, _vp]),
    "asr_ln_dropout_keep": (_c_int, [_vp, _c_int, _c_int, _c_float, _c_uint64, _vp]),
    "asr_allreduce_signal_bytes": (_c_size_t, [_c_int, _c_int]),
    "asr_allreduce_mean_f32": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_size_t, _c_size_t, _c_int, _vp]),
    "asr_gemm_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "asr_gemm_f32": (_c_int, [_vp, _c_int, _c_int, _vp, _c_int, _c_int, _vp, _c_int, _c_int, _c_int, _vp, _c_int,
                              _vp, _c_size_t, _vp]),
    "asr_gemm_f32_ragged": (_c_int, [_vp, _c_int, _c_int, _vp, _c_int, _c_int, _vp, _c_int, _c_int, _c_int, _vp, _c_int,
                                     _vp, _c_int, _c_int, _vp, _c_size_t, _vp]),
    "asr_gemm_bf16": (_c_int, [_vp, _c_int, _c_int, _vp, _c_int, _c_int, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _c_int,
                               _c_int, _vp, _c_size_t, _vp]),
    "asr_ctc_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "asr_ctc_fwd_bwd_ld_f32": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                        _vp, _vp, _vp, _c_size_t, _vp]),
    "asr_ctc_fwd_bwd_f32": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int,
                                     _vp, _vp, _vp, _c_size_t, _vp]),
    "asr_ctc_begin_f32": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int,
                                   _vp, _vp, _vp, _c_size_t, _vp, ctypes.POINTER(_c_int)]),
    "asr_ctc_finish_f32": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int,
                                    _vp, _vp, _vp, _c_size_t, _vp, _c_int]),
    "asr_ctc_stages_f32": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int,
                                    _vp, _vp, _vp, _c_size_t, _c_int, _vp]),
    "asr_scale_inplace_f32": (_c_int, [_vp, _c_size_t, _vp, _vp]),
    "asr_mha_fwd_bf16": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                  _c_float, _vp, _vp, _vp]),
    "asr_mha_bwd_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int, _c_int]),
    "asr_mha_bwd_bf16": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int,
                                  _c_int, _c_int, _c_int, _c_int, _c_int, _c_float,
                                  _vp, _vp, _vp, _vp, _c_size_t, _vp]),
    "asr_mha_fwd_dropout_bf16": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                          _c_float, _c_float, ctypes.c_uint64, _vp, _vp, _vp]),
    "asr_mha_bwd_dropout_bf16": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int,
                                          _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, ctypes.c_uint64,
                                          _vp, _vp, _vp, _vp, _c_size_t, _vp]),
    "asr_mha_fwd_dropout_dev_bf16": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                              _c_float, _c_float, _vp, ctypes.c_uint64, _vp, _vp, _vp]),
    "asr_mha_bwd_dropout_dev_bf16": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int,
                                              _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, _vp, ctypes.c_uint64,
                                              _vp, _vp, _vp, _vp, _c_size_t, _vp]),
    "asr_mha_dropout_keep_prob": (_c_float, [_c_float]),
    "asr_mha_dropout_keep_u8": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_float, ctypes.c_uint64, _vp, _vp]),
    "asr_mha_probs_f32": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                   _c_float, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


def _load():
    if not os.path.exists(LIB_PATH):
        # built in-tree by __graft_entry__.build(); try once here so a fresh checkout works
        from . import build_ext
        build_ext.build()
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)   # AttributeError if the library is stale
        fn.restype = res
        fn.argtypes = args
    if handle.asr_abi_version() != 1:
        raise RuntimeError("libasr_sm100.so ABI version mismatch; rebuild with build_ext.py --force")
    return handle


def lib():
    """The loaded shared library (loads / builds on first use; raises on failure)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                _lib = _load()
    return _lib


def last_error():
    msg = lib().asr_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, last_error()))


def set_option(key, value):
    check(lib().asr_set_option(key.encode(), int(value)), "asr_set_option(%s)" % key)


def get_option(key):
    v = _c_int(0)
    check(lib().asr_get_option(key.encode(), ctypes.byref(v)), "asr_get_option(%s)" % key)
    return v.value


def launch_count():
    return int(lib().asr_launch_count())


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
