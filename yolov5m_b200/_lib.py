"""ctypes binding of libyolov5m_b200.so (the C ABI in include/yolov5m_b200.h).

The prototypes are read from the header itself, so the binding cannot drift from the
ABI.  There is no fallback: if the shared library is missing or a call fails the caller
gets an exception -- nothing here ever routes to a CPU/PyTorch path.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("YB_LIB") or os.path.join(_HERE, "libyolov5m_b200.so")  # YB_LIB: A/B another build of the same ABI
_HEADER = os.path.join(_HERE, "..", "include", "yolov5m_b200.h")
_lib = None

c_int, c_i64, c_vp, c_f = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float
c_fp = ctypes.POINTER(ctypes.c_float)
c_ip = ctypes.POINTER(ctypes.c_int)


class YBError(RuntimeError):
    pass


_SCALARS = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "double": ctypes.c_double,
            "uint64_t": ctypes.c_uint64, "unsigned": ctypes.c_uint}


def _ctype(decl):
    decl = decl.strip()
    if decl == "void":
        return None
    if "*" in decl:
        return ctypes.c_char_p if decl.replace(" ", "").startswith("constchar*") else ctypes.c_void_p
    base = decl.replace("const", "").split()[0]
    return _SCALARS[base]


def prototypes(header=_HEADER):
    """{name: (restype, [argtypes])} parsed from the C header."""
    src = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?[\s\*])(yb_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argt = []
        if args.strip() != "void":
            for a in args.split(","):
                a = a.strip()
                # drop the parameter name (last identifier) unless the decl ends with '*'
                a = re.sub(r"\b\w+$", "", a).strip() if not a.endswith("*") else a
                argt.append(_ctype(a))
        out[name] = (_ctype(ret), argt)
    return out


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise YBError(f"{_LIB_PATH} not built: run `python -m yolov5m_b200.build` "
                          "(there is no CPU / PyTorch fallback for the hot path)")
        L = ctypes.CDLL(_LIB_PATH)
        for name, (ret, argt) in prototypes().items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = ret
            fn.argtypes = argt
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise YBError(f"yolov5m_b200 native call failed ({rc}): {lib().yb_last_error().decode()}")


def checkp(p, what="plan"):
    if not p:
        raise YBError(f"yolov5m_b200: could not create {what}: {lib().yb_last_error().decode()}")
    return p


def ptr(t):
    """device pointer of a torch tensor (or NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "hot-path tensors must live on the GPU"
    return t.data_ptr()


def stream():
    if not torch.cuda.is_available():
        raise YBError("yolov5m_b200: a CUDA device is required (there is no CPU / PyTorch fallback for the hot path)")
    return torch.cuda.current_stream().cuda_stream
