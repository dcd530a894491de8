"""ctypes binding of libe4t_b200.so — the thin C-ABI the module mirror (e4t.*) calls into.

There is deliberately NO fallback: if the CUDA library is missing, or a tensor is not on a CUDA device,
the call raises.  (The CPU oracle lives under /oracle and is test infrastructure only.)
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libe4t_b200.so")
_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float


class E4TError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise E4TError(
                f"{_LIB_PATH} not found: build it with `python e4t-diffusion_b200/csrc/build.py` "
                "(or __graft_entry__.build()); there is no CPU fallback on the product path")
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.e4t_last_error.restype = ctypes.c_char_p
        _lib.e4t_version.restype = c_int
        _lib.e4t_launch_count.restype = ctypes.c_ulonglong
        _lib.e4t_reset_launch_count.restype = None
    return _lib


def check(status):
    if status != 0:
        raise E4TError(load().e4t_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL).  Tensors must live on a CUDA device."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise E4TError("e4t_b200 kernels need CUDA tensors (no CPU fallback)")
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(load().e4t_launch_count())


def reset_launch_count():
    load().e4t_reset_launch_count()


def call(name, *args):
    fn = getattr(load(), name)
    check(fn(*args))
