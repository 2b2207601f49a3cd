"""ctypes binding of the C ABI in include/clipa_b200.h.

This is the ONLY route from Python to the kernels.  There is deliberately no fallback: if the
shared library is missing or an entry point fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libclipa_b200.so"

# enums (mirror include/clipa_b200.h)
BF16, F32 = 0, 1
MAJOR_K, MAJOR_MN = 0, 1
ACT_GELU_ERF, ACT_GELU_TANH, ACT_QUICK_GELU = 0, 1, 2
EPI_STORE, EPI_BIAS_ACT, EPI_DACT, EPI_ATOMIC_F32 = 0, 1, 2, 3

ABI_VERSION = 2


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_major", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_major", C.c_int32),
        ("C", C.c_void_p), ("ldc", C.c_int64), ("c_dtype", C.c_int32),
        ("epilogue", C.c_int32),
        ("alpha", C.c_float),
        ("bias", C.c_void_p), ("bias_dtype", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("act", C.c_int32),
        ("split_k", C.c_int32),
        ("max_ctas", C.c_int32),
        ("aux_is_derivative", C.c_int32),
    ]


class ClipaError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "clipa_abi_version": (C.c_int, []),
    "clipa_last_error": (C.c_char_p, []),
    "clipa_launch_count": (C.c_int64, []),
    "clipa_gemm": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "clipa_set_gemm_mode": (C.c_int, [C.c_int]),
    "clipa_layernorm_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p]),
    "clipa_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_int32, C.c_void_p]),
    "clipa_attention_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "clipa_attention_bwd_workspace": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "clipa_attention_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_void_p]),
    "clipa_set_attention_mode": (C.c_int, [C.c_int]),
    "clipa_colsum_accum": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_void_p]),
    "clipa_adamw_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                   C.c_float, C.c_void_p, C.c_int32, C.c_void_p]),
    "clipa_clip_lse_workspace": (C.c_int64, [C.c_int32, C.c_int32]),
    "clipa_clip_lse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                 C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "clipa_clip_softmax_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_float, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "clipa_preprocess_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]),
    "clipa_patchify": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_void_p]),
    "clipa_assemble_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                        C.c_int32, C.c_void_p]),
    "clipa_assemble_tokens_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "clipa_embed_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_void_p]),
    "clipa_embed_tokens_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_void_p]),
    "clipa_pool_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p]),
    "clipa_pool_tokens_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_void_p]),
    "clipa_l2_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_void_p]),
    "clipa_l2_normalize_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_int32, C.c_void_p]),
    "clipa_gemm_f32": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_int64, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                 C.c_void_p]),
    "clipa_act_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "clipa_layernorm_f32_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int64, C.c_int32, C.c_float, C.c_void_p]),
    "clipa_layernorm_f32_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "clipa_attention_f32_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_void_p]),
    "clipa_attention_f32_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "clipa_colsum_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "clipa_gemv_f32_accum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "clipa_row_lse_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "clipa_softmax_grad_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

POOL_FIRST, POOL_LAST, POOL_ARGMAX_ID, POOL_MEAN_ALL, POOL_MEAN_SKIP_FIRST = 0, 1, 2, 3, 4

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it is absent or ABI-mismatched."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("CLIPA_B200_LIB", LIB_PATH))
    if not path.exists():
        raise ClipaError(
            f"{path} not found: build it with `python -m clipa_b200.build` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback path.")
    handle = C.CDLL(str(path))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is missing -> loud failure
        fn.restype = res
        fn.argtypes = args
    got = handle.clipa_abi_version()
    if got != ABI_VERSION:
        raise ClipaError(f"ABI mismatch: library {got}, binding {ABI_VERSION}")
    _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().clipa_last_error().decode("utf-8", "replace")
        raise ClipaError(f"{what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(lib().clipa_launch_count())
