"""ctypes binding of ``libsovits_b200.so`` (C ABI declared in ``include/sovits_b200.h``).

There is deliberately no fallback: if the shared library is missing this module raises, and every
compute entry point lives in the library (hand-written sm_100a CUDA).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsovits_b200.so")

SVB_OK = 0
PREC_FP32 = 0
PREC_TC = 1


class SvbError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        super().__init__(f"libsovits_b200: {what} failed with status {code}: {detail}")
        self.code = code


class svb_tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4)]


class svb_model_cfg(C.Structure):
    _fields_ = [
        ("inter_channels", C.c_int32), ("hidden_channels", C.c_int32), ("gin_channels", C.c_int32),
        ("n_flows", C.c_int32), ("flow_wn_layers", C.c_int32), ("flow_kernel_size", C.c_int32),
        ("upsample_initial_channel", C.c_int32), ("n_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * 8), ("upsample_kernel_sizes", C.c_int32 * 8),
        ("n_resblock_kernels", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * 4),
        ("resblock_dilations", (C.c_int32 * 3) * 4),
        ("sampling_rate", C.c_int32), ("n_harmonics", C.c_int32), ("snake", C.c_int32), ("num_mels", C.c_int32),
        ("ssl_dim", C.c_int32), ("enc_layers", C.c_int32), ("enc_heads", C.c_int32), ("enc_filter", C.c_int32),
        ("enc_kernel", C.c_int32), ("enc_window", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/sovits_b200.h declares
SIGNATURES = {
    "svb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "svb_destroy": (None, [C.c_void_p]),
    "svb_load_weights": (C.c_int, [C.c_void_p, C.POINTER(svb_tensor), C.c_int, C.POINTER(svb_model_cfg)]),
    "svb_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "svb_get_precision": (C.c_int, [C.c_void_p]),
    "svb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "svb_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "svb_flow_reverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "svb_nsf_source": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "svb_generator": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "svb_vocoder": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_size_t, C.c_void_p]),
    "svb_infer_tail": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "svb_pre_conv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "svb_enc_p": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                            C.c_void_p]),
    "svb_infer_tail_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "svb_strerror": (C.c_char_p, [C.c_int]),
    "svb_last_error": (C.c_char_p, [C.c_void_p]),
    "svb_launch_count": (C.c_int64, [C.c_void_p]),
    "svb_fallback_count": (C.c_int64, [C.c_void_p]),
    "svb_debug_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "svb_debug_fetch": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "svb_debug_pair": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "svb_debug_resblock": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_float, C.c_void_p]),
    "svb_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "svb_profile_read": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "svb_prefix_add_ln_im2col": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "svb_prefix_ffn_tail": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "svb_prefix_rel_softmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "svb_prefix_attn_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]),
    "svb_version": (C.c_char_p, []),
}

_lib: Optional[C.CDLL] = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(path):
        raise ImportError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(lib: C.CDLL, ctx, code: int, what: str) -> None:
    if code != SVB_OK:
        detail = lib.svb_last_error(ctx).decode() if ctx else ""
        raise SvbError(code, what, f"{lib.svb_strerror(code).decode()} — {detail}")
