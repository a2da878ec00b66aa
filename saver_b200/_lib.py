"""ctypes loader for libsaver_cuda.so (the C ABI in include/saver_cuda.h).

There is no CPU fallback: if the library is missing it is built (nvcc, in-tree); if it
cannot be loaded, or no CUDA device is usable, the error propagates.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


class SaverCudaError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        path = _build.build_library()
    lib = C.CDLL(path)
    H = C.c_void_p
    lib.saver_cuda_abi_version.restype = C.c_int
    lib.saver_cuda_create.argtypes = [C.c_int, C.POINTER(H)]
    lib.saver_cuda_destroy.argtypes = [H]
    lib.saver_cuda_last_error.argtypes = [H]
    lib.saver_cuda_last_error.restype = C.c_char_p
    lib.saver_cuda_synchronize.argtypes = [H]
    lib.saver_cuda_stream.argtypes = [H, C.POINTER(C.c_void_p)]
    lib.saver_cuda_launch_count.argtypes = [H]
    lib.saver_cuda_launch_count.restype = C.c_longlong
    _lib = lib
    return lib


def ptr(a):
    """numpy array / torch tensor / None -> void pointer value."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())  # torch tensor


def is_device(a):
    return (a is not None) and (not isinstance(a, np.ndarray)) and bool(getattr(a, "is_cuda", False))


class Handle:
    """One libsaver_cuda handle = one CUDA device (one process per GPU)."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.saver_cuda_create(int(device), C.byref(self.h))
        if rc != 0:
            raise SaverCudaError(
                "saver_cuda_create(device=%d) failed with code %d: no usable CUDA device "
                "(libsaver_cuda has no CPU fallback)" % (device, rc))
        self.device = int(device)

    def check(self, rc):
        if rc != 0:
            msg = self.lib.saver_cuda_last_error(self.h)
            raise SaverCudaError("libsaver_cuda error %d: %s" % (rc, (msg or b"").decode()))

    def synchronize(self):
        self.check(self.lib.saver_cuda_synchronize(self.h))

    def stream(self):
        s = C.c_void_p()
        self.check(self.lib.saver_cuda_stream(self.h, C.byref(s)))
        return s.value or 0

    def set_option(self, key, value):
        self.check(self.lib.saver_cuda_set_option(self.h, key.encode(), C.c_double(float(value))))

    def launch_count(self):
        return int(self.lib.saver_cuda_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.saver_cuda_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}


def default_handle(device=0):
    if device not in _default:
        _default[device] = Handle(device)
    return _default[device]
