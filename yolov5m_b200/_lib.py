"""ctypes binding of libyolov5m_b200.so (the C ABI in include/yolov5m_b200.h).

There is no fallback: if the shared library is missing or a call fails the
caller gets an exception -- nothing here ever routes to a CPU/PyTorch path.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libyolov5m_b200.so")
_lib = None

c_int, c_i64, c_vp, c_f = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float
c_fp = ctypes.POINTER(ctypes.c_float)
c_ip = ctypes.POINTER(ctypes.c_int)


class YBError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise YBError(f"{_LIB_PATH} not built: run `python -m yolov5m_b200.build` "
                          "(there is no CPU / PyTorch fallback for the hot path)")
        L = ctypes.CDLL(_LIB_PATH)
        L.yb_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise YBError(f"yolov5m_b200 native call failed ({rc}): {lib().yb_last_error().decode()}")


def ptr(t):
    """device pointer of a torch tensor (or NULL)."""
    if t is None:
        return c_vp(0)
    assert t.is_cuda, "hot-path tensors must live on the GPU"
    return c_vp(t.data_ptr())


def stream():
    return c_vp(torch.cuda.current_stream().cuda_stream)
