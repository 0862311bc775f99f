"""ctypes binding of the C ABI declared in include/ndfft_b200.h.

The product path is the nvcc-built `ndrustfft_b200/lib/libndfft_b200.so` (hand-written sm_100a kernels).
There is NO CPU fallback: if that library is missing this module raises, and if no CUDA device is usable
every transform raises `NdfftError` (NDFB_E_CUDA).  `CLib(path)` accepts an explicit path only so that the
test-suite can bind the SIMT-emulation build of the same sources (tests/emu/, CPU-side kernel-logic tests);
nothing in this package ever points it there.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "lib", "libndfft_b200.so")

# enums (include/ndfft_b200.h)
C2C, R2C, DCT = 0, 1, 2
F32, F64 = 0, 1
OP_FFT, OP_IFFT, OP_R2C, OP_C2R, OP_DCT1, OP_DCT2, OP_DCT3, OP_DCT4 = range(8)
NORM_NONE, NORM_DEFAULT = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
E_INVALID, E_SIZE_MISMATCH, E_SHAPE, E_AXIS, E_ALLOC, E_CUDA, E_UNSUPPORTED = -1, -2, -3, -4, -5, -6, -7

EXPORTS = (
    "ndfb_plan_create", "ndfb_plan_destroy", "ndfb_plan_describe", "ndfb_exec", "ndfb_exec_scaled", "ndfb_exec_split_out", "ndfb_exec_scatter_out",
    "ndfb_exec_chain", "ndfb_jit_compile_check",
    "ndfb_device_alloc", "ndfb_device_free", "ndfb_memcpy", "ndfb_stream_create", "ndfb_stream_destroy", "ndfb_stream_sync",
    "ndfb_hint_next_launch_smem", "ndfb_hint_next_launch_signal", "ndfb_hint_next_launch_wait", "ndfb_last_error", "ndfb_version", "ndfb_launch_count", "ndfb_release_workspaces",
)


class NdfftError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"ndfft_b200 error {code}: {message}")
        self.code = code
        self.message = message


class SizeMismatch(AssertionError):
    """Mirrors the reference's `assert_size` panic (src/lib.rs:340-347, 533-540, 743-750)."""


class Step(ctypes.Structure):
    """struct ndfb_step (include/ndfft_b200.h)."""
    _fields_ = [("plan", ctypes.c_void_p), ("op", ctypes.c_int), ("norm", ctypes.c_int), ("axis", ctypes.c_int)]


class CLib:
    def __init__(self, path=None):
        path = path or DEFAULT_LIB
        if not os.path.exists(path):
            raise ImportError(
                f"ndrustfft_b200: native library not found at {path}. Build it with `make lib` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        self.path = path
        self.dll = ctypes.CDLL(path)
        d = self.dll
        vp, ci, cz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        d.ndfb_plan_create.argtypes = [ctypes.POINTER(vp), ci, ci, cz, ci]
        d.ndfb_plan_create.restype = ci
        d.ndfb_plan_destroy.argtypes = [vp]
        d.ndfb_plan_destroy.restype = None
        d.ndfb_plan_describe.argtypes = [vp, ctypes.c_char_p, cz]
        d.ndfb_plan_describe.restype = cz
        szp, pdp = ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_ssize_t)
        d.ndfb_exec.argtypes = [vp, ci, ci, vp, vp, ci, szp, pdp, szp, pdp, ci, ci, vp]
        d.ndfb_exec.restype = ci
        d.ndfb_exec_scaled.argtypes = [vp, ci, ci, ctypes.c_double, vp, vp, ci, szp, pdp, szp, pdp, ci, ci, vp]
        d.ndfb_exec_scaled.restype = ci
        d.ndfb_exec_split_out.argtypes = [vp, ci, ci, ctypes.c_double, cz, ctypes.c_ssize_t, vp, vp, ci, szp, pdp, szp, pdp, ci, vp]
        d.ndfb_exec_split_out.restype = ci
        d.ndfb_exec_scatter_out.argtypes = [vp, ci, ci, ctypes.c_double, cz, ci, ctypes.POINTER(vp), vp, ci, szp, pdp, szp, pdp, ci, vp]
        d.ndfb_exec_scatter_out.restype = ci
        d.ndfb_exec_chain.argtypes = [ctypes.POINTER(Step), ci, vp, vp, ci, szp, pdp, szp, pdp, ci, vp]
        d.ndfb_exec_chain.restype = ci
        d.ndfb_jit_compile_check.argtypes = [ci, ci, cz, ci, ctypes.c_char_p, cz]
        d.ndfb_jit_compile_check.restype = ci
        d.ndfb_device_alloc.argtypes = [ctypes.POINTER(vp), cz, ci]
        d.ndfb_device_alloc.restype = ci
        d.ndfb_device_free.argtypes = [vp]
        d.ndfb_device_free.restype = None
        d.ndfb_memcpy.argtypes = [vp, vp, cz, ci, ci, vp]
        d.ndfb_memcpy.restype = ci
        d.ndfb_stream_create.argtypes = [ctypes.POINTER(vp), ci]
        d.ndfb_stream_create.restype = ci
        d.ndfb_stream_destroy.argtypes = [vp]
        d.ndfb_stream_destroy.restype = None
        d.ndfb_stream_sync.argtypes = [vp]
        d.ndfb_stream_sync.restype = ci
        d.ndfb_hint_next_launch_smem.argtypes = [cz]
        d.ndfb_hint_next_launch_smem.restype = None
        d.ndfb_hint_next_launch_signal.argtypes = [vp, ctypes.c_longlong]
        d.ndfb_hint_next_launch_signal.restype = None
        d.ndfb_hint_next_launch_wait.argtypes = [vp, ctypes.c_longlong, ctypes.c_uint, ci]
        d.ndfb_hint_next_launch_wait.restype = None
        d.ndfb_last_error.restype = ctypes.c_char_p
        d.ndfb_version.restype = ctypes.c_char_p
        d.ndfb_launch_count.restype = ctypes.c_uint64
        d.ndfb_release_workspaces.restype = None

    def version(self):
        return self.dll.ndfb_version().decode()

    def last_error(self):
        return self.dll.ndfb_last_error().decode()

    def launch_count(self):
        return int(self.dll.ndfb_launch_count())

    def check(self, rc):
        if rc == 0:
            return
        msg = self.last_error()
        if rc == E_SIZE_MISMATCH:
            raise SizeMismatch(msg)
        if rc == E_AXIS:
            raise IndexError(msg)
        if rc == E_SHAPE:
            raise AssertionError(msg)
        raise NdfftError(rc, msg)


_default = None
_default_lock = threading.Lock()


def default_lib():
    """The CUDA library; raises ImportError when it has not been built."""
    global _default
    with _default_lock:
        if _default is None:
            _default = CLib(DEFAULT_LIB)
        return _default
