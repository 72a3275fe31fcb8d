"""ctypes loader for libmudg_sm100.so.  Fails loudly: there is no CPU or PyTorch fallback for the product path."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MUDG_LIB_PATH") or os.path.join(_HERE, "libmudg_sm100.so")   # override: A/B builds of the same ABI
TEST_LIB_PATH = os.environ.get("MUDG_TEST_LIB_PATH") or os.path.join(_HERE, "libmudg_sm100_test.so")   # tests/ only: product objects + csrc/test/ (include/mudg_test.h)
_lib = None
_test_lib = None


class MudgError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MudgError(f"{LIB_PATH} is missing: run `python -m mudg_b200.build` (or __graft_entry__.build()) first; "
                            "the CUDA extension is mandatory, there is no fallback path")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mudg_last_error.restype = ctypes.c_char_p
    return _lib


def test_lib() -> ctypes.CDLL:
    """The TEST library (single-kernel hooks, CUDA-core checkers, knobs).  It carries its own copy of the product
    objects, so its state (knobs, last error) is separate from `lib()`.  Never used by the product modules."""
    global _test_lib
    if _test_lib is None:
        if not os.path.exists(TEST_LIB_PATH):
            raise MudgError(f"{TEST_LIB_PATH} is missing: run `python -m mudg_b200.build`")
        _test_lib = ctypes.CDLL(TEST_LIB_PATH)
        _test_lib.mudg_last_error.restype = ctypes.c_char_p
    return _test_lib


def check(rc: int) -> None:
    if rc != 0:
        # last-error strings are per library (thread local): report whichever loaded library holds one
        msgs = [L.mudg_last_error().decode("utf-8", "replace") for L in (_lib, _test_lib) if L is not None]
        raise MudgError(" | ".join(m for m in msgs if m) or f"libmudg call failed (rc={rc})")


def ptr(t):
    """torch tensor (or None) -> void*"""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
