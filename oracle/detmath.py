"""ctypes access to the host build of csrc/detmath.h (oracle/detmath_host.c).  TEST INFRASTRUCTURE."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_detmath.so")
_lib = None


def build(force=False):
    """Compile the oracle's C helper (gcc); called by __graft_entry__.build() and lazily on first use."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        for name in ("om_exp", "om_sigmoid", "om_log", "om_log1p", "om_atan", "om_pow15"):
            getattr(lib, name).argtypes = [fp, fp, ctypes.c_size_t]
            getattr(lib, name).restype = None
        lib.om_pow.argtypes = [fp, ctypes.c_float, fp, ctypes.c_size_t]
        lib.om_pow.restype = None
        lib.om_bce_logits.argtypes = [fp, fp, fp, ctypes.c_size_t]
        lib.om_bce_logits.restype = None
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _unary(name, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    getattr(_load(), name)(_ptr(x), _ptr(out), x.size)
    return out


def exp(x):
    return _unary("om_exp", x)


def sigmoid(x):
    return _unary("om_sigmoid", x)


def log(x):
    return _unary("om_log", x)


def log1p(x):
    return _unary("om_log1p", x)


def atan(x):
    return _unary("om_atan", x)


def pow15(x):
    return _unary("om_pow15", x)


def pow(x, y):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    _load().om_pow(_ptr(x), ctypes.c_float(y), _ptr(out), x.size)
    return out


def bce_logits(z, x):
    """tf.nn.sigmoid_cross_entropy_with_logits(labels=z, logits=x): (max(x,0) - x*z) + log1p(exp(-|x|))."""
    z, x = np.broadcast_arrays(np.asarray(z, np.float32), np.asarray(x, np.float32))
    z = np.ascontiguousarray(z)
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    _load().om_bce_logits(_ptr(z), _ptr(x), _ptr(out), x.size)
    return out
