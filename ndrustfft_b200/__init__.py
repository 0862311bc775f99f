"""ndrustfft_b200 — B200-native drop-in for ndrustfft's axis-transform hot path.

Host-side mirror of the reference's public API (preiter93/ndrustfft v0.5.0, src/lib.rs):

    Normalization {None, Default, Custom(fn)}          src/lib.rs:89-98
    FftHandler / R2cFftHandler / DctHandler            src/lib.rs:270-311, 452-495, 641-686
    ndfft, ndifft, ndfft_r2c, ndifft_r2c, nddct1..4    src/lib.rs:350-397, 543-587, 753-834
    *_par twins                                        src/lib.rs:399-421, 589-611, 777-844

Same names, argument order `(input, output, handler, axis)` and error behaviour (size-mismatch text, axis range,
shape agreement).  Arrays are numpy arrays (host path: staged through the GPU) or torch CUDA tensors / anything with
`data_ptr()`, `shape`, `stride()` (device path: zero-copy, asynchronous on the current torch stream).  All
arithmetic runs in hand-written sm_100a kernels behind the C ABI of include/ndfft_b200.h; there is no CPU path.
On the GPU every call is already parallel over all lanes, so the `_par` functions are the same entry points.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import NdfftError, SizeMismatch  # noqa: F401

__all__ = [
    "Normalization", "FftHandler", "R2cFftHandler", "DctHandler",
    "ndfft", "ndifft", "ndfft_r2c", "ndifft_r2c", "nddct1", "nddct2", "nddct3", "nddct4",
    "ndfft_par", "ndifft_par", "ndfft_r2c_par", "ndifft_r2c_par",
    "nddct1_par", "nddct2_par", "nddct3_par", "nddct4_par",
    "ndchain", "fft2", "ifft2", "rfft2", "irfft2",
    "NdfftError", "SizeMismatch", "Backend",
]


class Normalization:
    """`Normalization<T>` (src/lib.rs:89-98): `Normalization.None_`, `.Default`, `.Custom(fn)`.

    `fn(lane)` receives one 1-D numpy lane and mutates it in place, exactly where the reference calls it:
    after the inverse C2C transform (src/lib.rs:329), on the spectrum copy before C2R (src/lib.rs:514), on the
    input copy before a DCT (src/lib.rs:695).  It is a host function, so `Custom` stages lanes through the host;
    `None_` and `Default` are fused into the kernels."""

    def __init__(self, kind, func=None):
        self.kind = kind
        self.func = func

    @classmethod
    def Custom(cls, func):
        return cls("custom", func)

    def __repr__(self):
        return f"Normalization.{self.kind}"


Normalization.None_ = Normalization("none")
Normalization.Default = Normalization("default")


def _real_dtype(dtype):
    dt = np.dtype(dtype)
    if dt == np.float32 or dt == np.complex64:
        return np.dtype(np.float32)
    if dt == np.float64 or dt == np.complex128:
        return np.dtype(np.float64)
    raise TypeError(f"unsupported dtype {dt}; ndrustfft supports f32 and f64 (FftNum, src/lib.rs:111)")


class _Handler:
    _kind = None

    def __init__(self, n, dtype=np.float64, device=0, backend=None):
        self.n = int(n)
        self.dtype = _real_dtype(dtype)
        self.device = int(device)
        self.norm = Normalization.Default
        self._backend = backend or _default_backend()
        self._plan = ctypes.c_void_p()
        lib = self._backend.lib
        lib.check(lib.dll.ndfb_plan_create(ctypes.byref(self._plan), self._kind,
                                          _lib.F32 if self.dtype == np.float32 else _lib.F64, self.n, self.device))

    def normalization(self, norm):
        """Builder, like the reference's `fn normalization(mut self, norm) -> Self`."""
        self.norm = norm
        return self

    def describe(self):
        import json
        lib = self._backend.lib
        need = lib.dll.ndfb_plan_describe(self._plan, None, 0)
        buf = ctypes.create_string_buffer(int(need))
        lib.dll.ndfb_plan_describe(self._plan, buf, need)
        return json.loads(buf.value.decode())

    def __del__(self):
        try:
            if self._plan:
                self._backend.lib.dll.ndfb_plan_destroy(self._plan)
                self._plan = ctypes.c_void_p()
        except Exception:
            pass


class FftHandler(_Handler):
    """`FftHandler<T>::new(n)` (src/lib.rs:294-304)."""
    _kind = _lib.C2C


class R2cFftHandler(_Handler):
    """`R2cFftHandler<T>::new(n)` (src/lib.rs:477-488); `m = n/2 + 1`."""
    _kind = _lib.R2C

    def __init__(self, n, dtype=np.float64, device=0, backend=None):
        super().__init__(n, dtype, device, backend)
        self.m = self.n // 2 + 1


class DctHandler(_Handler):
    """`DctHandler<T>::new(n)` (src/lib.rs:665-679)."""
    _kind = _lib.DCT


# ------------------------------------------------------------------------------------------------------
# array adapters
# ------------------------------------------------------------------------------------------------------
class _View:
    __slots__ = ("ptr", "shape", "strides", "dtype", "device", "obj", "stream")


def _span(v):
    """[lo, hi) byte range touched by a view."""
    item = v.dtype.itemsize
    lo = hi = 0
    for n, s in zip(v.shape, v.strides):
        if n == 0:
            return v.ptr, v.ptr
        ext = (n - 1) * s * item
        if ext < 0:
            lo += ext
        else:
            hi += ext
    return v.ptr + lo, v.ptr + hi + item


def _view_of(arr, out=False):
    v = _View()
    v.obj = arr
    v.stream = None
    if isinstance(arr, np.ndarray):
        if out and not arr.flags.writeable:
            raise ValueError("output array is read-only")
        item = arr.dtype.itemsize
        for s in arr.strides:
            if s % item:
                raise ValueError("array strides must be multiples of the element size")
        v.ptr = arr.ctypes.data
        v.shape = tuple(arr.shape)
        v.strides = tuple(s // item for s in arr.strides)
        v.dtype = arr.dtype
        v.device = None
        return v
    if hasattr(arr, "data_ptr") and hasattr(arr, "stride"):  # torch tensor
        import torch
        # lazy conjugate / negative views share the storage of the original: the kernels would read the unflagged values
        if (arr.is_complex() and arr.is_conj()) or arr.is_neg():
            raise ValueError("tensor is a lazy conj/neg view; call .resolve_conj() / .resolve_neg() first")
        v.ptr = arr.data_ptr()
        v.shape = tuple(arr.shape)
        v.strides = tuple(arr.stride())
        v.dtype = np.dtype({torch.float32: np.float32, torch.float64: np.float64,
                            torch.complex64: np.complex64, torch.complex128: np.complex128}[arr.dtype])
        if arr.is_cuda:
            v.device = arr.device.index if arr.device.index is not None else torch.cuda.current_device()
            v.stream = torch.cuda.current_stream(arr.device).cuda_stream
        else:
            v.device = None
        return v
    raise TypeError(f"unsupported array type {type(arr)}")


class Backend:
    """The nd* functions bound to one loaded C library."""

    def __init__(self, lib):
        self.lib = lib

    def _exec(self, handler, op, inp, out, axis, norm_code, extra_scale=1.0):
        vi, vo = _view_of(inp), _view_of(out, out=True)
        if op in (_lib.OP_R2C, _lib.OP_C2R):
            # in place is part of the contract only for the ops that keep shape and element type
            (a0, a1), (b0, b1) = _span(vi), _span(vo)
            if a0 < b1 and b0 < a1:
                raise ValueError("input and output overlap: ndfft_r2c / ndifft_r2c cannot run in place")
        if len(vi.shape) != len(vo.shape):
            raise AssertionError("input and output must have the same number of dimensions")
        ndim = len(vi.shape)
        if ndim < 1:
            raise IndexError("0-dimensional arrays have no axis")
        want_in_cx = op in (_lib.OP_FFT, _lib.OP_IFFT, _lib.OP_C2R)
        want_out_cx = op in (_lib.OP_FFT, _lib.OP_IFFT, _lib.OP_R2C)
        rd = handler.dtype
        cd = np.dtype(np.complex64 if rd == np.float32 else np.complex128)
        if vi.dtype != (cd if want_in_cx else rd) or vo.dtype != (cd if want_out_cx else rd):
            raise TypeError(f"element types do not match the handler: got {vi.dtype} -> {vo.dtype}, "
                            f"expected {(cd if want_in_cx else rd)} -> {(cd if want_out_cx else rd)}")
        if (vi.device is None) != (vo.device is None):
            raise ValueError("input and output must both be host arrays or both be device tensors")
        if vi.device is not None and (vi.device != handler.device or vo.device != handler.device):
            raise ValueError("tensors must live on the handler's device")
        if not (0 <= int(axis) < ndim):
            raise IndexError(f"axis {axis} out of range for {ndim}-dimensional array")
        SZ = ctypes.c_size_t * ndim
        PD = ctypes.c_ssize_t * ndim
        mem = _lib.MEM_HOST if vi.device is None else _lib.MEM_DEVICE
        rc = self.lib.dll.ndfb_exec_scaled(
            handler._plan, op, norm_code, float(extra_scale), ctypes.c_void_p(vi.ptr), ctypes.c_void_p(vo.ptr), ndim,
            SZ(*vi.shape), PD(*vi.strides), SZ(*vo.shape), PD(*vo.strides), int(axis), mem,
            ctypes.c_void_p(vi.stream or 0))
        self.lib.check(rc)

    def ndfft_split_out(self, input, output, handler, axis, out_shape, out_strides, out_block, out_block_stride, inverse=False):
        """ndfft / ndifft on device tensors with the output axis stored in blocks (ndfb_exec_split_out): element k of an
        output lane lands at `(k // out_block) * out_block_stride + (k % out_block) * out_strides[axis]` from the lane base.
        `output` only provides the base pointer; `out_shape` / `out_strides` describe the logical output array."""
        vi, vo = _view_of(input), _view_of(output, out=True)
        if (vi.device is None or vo.device is None) and "emu" not in self.lib.version():
            raise ValueError("split-output transforms take device tensors")
        ndim = len(vi.shape)
        SZ = ctypes.c_size_t * ndim
        PD = ctypes.c_ssize_t * ndim
        norm = _lib.NORM_DEFAULT if handler.norm.kind == "default" else _lib.NORM_NONE
        rc = self.lib.dll.ndfb_exec_split_out(
            handler._plan, _lib.OP_IFFT if inverse else _lib.OP_FFT, norm, 1.0, int(out_block), int(out_block_stride),
            ctypes.c_void_p(vi.ptr), ctypes.c_void_p(vo.ptr), ndim, SZ(*vi.shape), PD(*vi.strides),
            SZ(*out_shape), PD(*out_strides), int(axis), ctypes.c_void_p(vi.stream or 0))
        self.lib.check(rc)

    def ndfft_scatter_out(self, input, handler, axis, out_shape, out_strides, out_block, block_ptrs, inverse=False):
        """ndfft / ndifft whose output blocks go to separate base pointers (ndfb_exec_scatter_out): block p of each lane is
        written relative to `block_ptrs[p]` (ints: device addresses, e.g. peer-mapped buffers of other GPUs)."""
        vi = _view_of(input)
        if vi.device is None and "emu" not in self.lib.version():
            raise ValueError("scatter-output transforms take device tensors")
        ndim = len(vi.shape)
        SZ = ctypes.c_size_t * ndim
        PD = ctypes.c_ssize_t * ndim
        PT = ctypes.c_void_p * len(block_ptrs)
        norm = _lib.NORM_DEFAULT if handler.norm.kind == "default" else _lib.NORM_NONE
        rc = self.lib.dll.ndfb_exec_scatter_out(
            handler._plan, _lib.OP_IFFT if inverse else _lib.OP_FFT, norm, 1.0, int(out_block), len(block_ptrs), PT(*[int(p) for p in block_ptrs]),
            ctypes.c_void_p(vi.ptr), ndim, SZ(*vi.shape), PD(*vi.strides), SZ(*out_shape), PD(*out_strides), int(axis),
            ctypes.c_void_p(vi.stream or 0))
        self.lib.check(rc)

    # -- Custom normalisation plumbing (host callback) --
    @staticmethod
    def _to_host(a):
        return a if isinstance(a, np.ndarray) else a.detach().cpu().numpy()

    @staticmethod
    def _assign(dst, src_np):
        if isinstance(dst, np.ndarray):
            dst[...] = src_np
        else:
            import torch
            dst.copy_(torch.from_numpy(np.ascontiguousarray(src_np)))

    @staticmethod
    def _apply_lanes(func, arr, axis):
        moved = np.moveaxis(arr, axis, -1)
        tmp = np.ascontiguousarray(moved)
        for lane in tmp.reshape(-1, tmp.shape[-1]):
            func(lane)
        moved[...] = tmp

    def _run(self, handler, op, inp, out, axis):
        norm = handler.norm
        if norm.kind == "none":
            return self._exec(handler, op, inp, out, axis, _lib.NORM_NONE)
        if norm.kind == "default":
            return self._exec(handler, op, inp, out, axis, _lib.NORM_DEFAULT)
        # Custom(fn)
        if op in (_lib.OP_FFT, _lib.OP_R2C):          # forward transforms never normalise (src/lib.rs:313-318, 497-503)
            return self._exec(handler, op, inp, out, axis, _lib.NORM_NONE)
        if op == _lib.OP_IFFT:                        # after the transform, on the output lane (src/lib.rs:329)
            self._exec(handler, op, inp, out, axis, _lib.NORM_NONE)
            host = np.array(self._to_host(out))
            self._apply_lanes(norm.func, host, axis)
            return self._assign(out, host)
        # C2R / DCT: on a copy of the input lane, before the transform (src/lib.rs:514, 695)
        host = np.array(self._to_host(inp))
        self._apply_lanes(norm.func, host, axis)
        if isinstance(inp, np.ndarray):
            staged = host
        else:
            import torch
            staged = torch.from_numpy(host).to(inp.device)
        return self._exec(handler, op, staged, out, axis, _lib.NORM_NONE)

    def ndfft(self, input, output, handler, axis): self._run(handler, _lib.OP_FFT, input, output, axis)
    def ndifft(self, input, output, handler, axis): self._run(handler, _lib.OP_IFFT, input, output, axis)
    def ndfft_r2c(self, input, output, handler, axis): self._run(handler, _lib.OP_R2C, input, output, axis)
    def ndifft_r2c(self, input, output, handler, axis): self._run(handler, _lib.OP_C2R, input, output, axis)
    def nddct1(self, input, output, handler, axis): self._run(handler, _lib.OP_DCT1, input, output, axis)
    def nddct2(self, input, output, handler, axis): self._run(handler, _lib.OP_DCT2, input, output, axis)
    def nddct3(self, input, output, handler, axis): self._run(handler, _lib.OP_DCT3, input, output, axis)
    def nddct4(self, input, output, handler, axis): self._run(handler, _lib.OP_DCT4, input, output, axis)

    # -- multi-axis chains (ndfb_exec_chain; the fft2 / rfft2 pattern of examples/fft2.rs:23-27, examples/rfft2.rs:29-33) --
    _CHAIN_OPS = {"ndfft": _lib.OP_FFT, "ndifft": _lib.OP_IFFT, "ndfft_r2c": _lib.OP_R2C, "ndifft_r2c": _lib.OP_C2R,
                  "nddct1": _lib.OP_DCT1, "nddct2": _lib.OP_DCT2, "nddct3": _lib.OP_DCT3, "nddct4": _lib.OP_DCT4}

    def ndchain(self, input, output, steps):
        """Applies `steps` = [(function name, handler, axis), ...] in order, e.g. rfft2 of the reference's example is
        `[("ndfft_r2c", handler_ax1, 1), ("ndfft", handler_ax0, 0)]`.  Same result as the separate calls through `work`
        arrays; the intermediates stay on the GPU and host arrays cross PCIe once each way."""
        steps = [(str(name).removesuffix("_par"), h, int(ax)) for name, h, ax in steps]
        if not steps:
            raise ValueError("at least one step expected")
        for name, _, _ in steps:
            if name not in self._CHAIN_OPS:
                raise ValueError(f"unknown transform {name!r}")
        if any(h.norm.kind == "custom" for _, h, _ in steps):
            return self._chain_stepwise(input, output, steps)
        vi, vo = _view_of(input), _view_of(output, out=True)
        if len(vi.shape) != len(vo.shape):
            raise AssertionError("input and output must have the same number of dimensions")
        ndim = len(vi.shape)
        if ndim < 1:
            raise IndexError("0-dimensional arrays have no axis")
        first, last = self._CHAIN_OPS[steps[0][0]], self._CHAIN_OPS[steps[-1][0]]
        rd = steps[0][1].dtype
        cd = np.dtype(np.complex64 if rd == np.float32 else np.complex128)
        want_in = cd if first in (_lib.OP_FFT, _lib.OP_IFFT, _lib.OP_C2R) else rd
        want_out = cd if last in (_lib.OP_FFT, _lib.OP_IFFT, _lib.OP_R2C) else rd
        if vi.dtype != want_in or vo.dtype != want_out:
            raise TypeError(f"element types do not match the handlers: got {vi.dtype} -> {vo.dtype}, expected {want_in} -> {want_out}")
        if (vi.device is None) != (vo.device is None):
            raise ValueError("input and output must both be host arrays or both be device tensors")
        if vi.device is not None and any(vi.device != h.device or vo.device != h.device for _, h, _ in steps):
            raise ValueError("tensors must live on the handlers' device")
        arr = (_lib.Step * len(steps))()
        for i, (name, h, ax) in enumerate(steps):
            arr[i].plan, arr[i].op, arr[i].axis = h._plan.value if hasattr(h._plan, "value") else h._plan, self._CHAIN_OPS[name], ax
            arr[i].norm = _lib.NORM_DEFAULT if h.norm.kind == "default" else _lib.NORM_NONE
        SZ = ctypes.c_size_t * ndim
        PD = ctypes.c_ssize_t * ndim
        mem = _lib.MEM_HOST if vi.device is None else _lib.MEM_DEVICE
        rc = self.lib.dll.ndfb_exec_chain(arr, len(steps), ctypes.c_void_p(vi.ptr), ctypes.c_void_p(vo.ptr), ndim,
                                          SZ(*vi.shape), PD(*vi.strides), SZ(*vo.shape), PD(*vo.strides), mem,
                                          ctypes.c_void_p(vi.stream or 0))
        self.lib.check(rc)

    def _chain_stepwise(self, input, output, steps):
        # Custom(fn) normalisation runs on the host between steps: one call per axis through temporaries
        cur = input
        for i, (name, h, ax) in enumerate(steps):
            op = self._CHAIN_OPS[name]
            if i == len(steps) - 1:
                dst = output
            else:
                shape = list(cur.shape)
                if op == _lib.OP_R2C: shape[ax] = h.n // 2 + 1
                if op == _lib.OP_C2R: shape[ax] = h.n
                cplx = op in (_lib.OP_FFT, _lib.OP_IFFT, _lib.OP_R2C)
                cd = np.complex64 if h.dtype == np.float32 else np.complex128
                if isinstance(cur, np.ndarray):
                    dst = np.zeros(shape, dtype=cd if cplx else h.dtype)
                else:
                    import torch
                    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                           np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}[np.dtype(cd if cplx else h.dtype)]
                    dst = torch.zeros(shape, dtype=tdt, device=cur.device)
            self._run(h, op, cur, dst, ax)
            cur = dst

    def fft2(self, input, output, handler_ax0, handler_ax1):
        """examples/fft2.rs:23-27: ndfft along axis 1, then along axis 0."""
        self.ndchain(input, output, [("ndfft", handler_ax1, 1), ("ndfft", handler_ax0, 0)])

    def ifft2(self, input, output, handler_ax0, handler_ax1):
        """examples/fft2.rs:55-59: ndifft along axis 0, then along axis 1."""
        self.ndchain(input, output, [("ndifft", handler_ax0, 0), ("ndifft", handler_ax1, 1)])

    def rfft2(self, input, output, handler_ax0, handler_ax1):
        """examples/rfft2.rs:29-33: ndfft_r2c along axis 1 (R2cFftHandler), then ndfft along axis 0."""
        self.ndchain(input, output, [("ndfft_r2c", handler_ax1, 1), ("ndfft", handler_ax0, 0)])

    def irfft2(self, input, output, handler_ax0, handler_ax1):
        """examples/rfft2.rs:49-53: ndifft along axis 0, then ndifft_r2c along axis 1."""
        self.ndchain(input, output, [("ndifft", handler_ax0, 0), ("ndifft_r2c", handler_ax1, 1)])

    ndfft_par, ndifft_par = ndfft, ndifft
    ndfft_r2c_par, ndifft_r2c_par = ndfft_r2c, ndifft_r2c
    nddct1_par, nddct2_par, nddct3_par, nddct4_par = nddct1, nddct2, nddct3, nddct4

    # handler constructors bound to this backend
    def FftHandler(self, n, dtype=np.float64, device=0): return FftHandler(n, dtype, device, backend=self)
    def R2cFftHandler(self, n, dtype=np.float64, device=0): return R2cFftHandler(n, dtype, device, backend=self)
    def DctHandler(self, n, dtype=np.float64, device=0): return DctHandler(n, dtype, device, backend=self)


_backend = None


def _default_backend():
    global _backend
    if _backend is None:
        _backend = Backend(_lib.default_lib())   # raises ImportError if the CUDA library is not built
    return _backend


def ndfft(input, output, handler, axis): handler._backend.ndfft(input, output, handler, axis)
def ndifft(input, output, handler, axis): handler._backend.ndifft(input, output, handler, axis)
def ndfft_r2c(input, output, handler, axis): handler._backend.ndfft_r2c(input, output, handler, axis)
def ndifft_r2c(input, output, handler, axis): handler._backend.ndifft_r2c(input, output, handler, axis)
def nddct1(input, output, handler, axis): handler._backend.nddct1(input, output, handler, axis)
def nddct2(input, output, handler, axis): handler._backend.nddct2(input, output, handler, axis)
def nddct3(input, output, handler, axis): handler._backend.nddct3(input, output, handler, axis)
def nddct4(input, output, handler, axis): handler._backend.nddct4(input, output, handler, axis)
def ndchain(input, output, steps): steps[0][1]._backend.ndchain(input, output, steps)
def fft2(input, output, handler_ax0, handler_ax1): handler_ax0._backend.fft2(input, output, handler_ax0, handler_ax1)
def ifft2(input, output, handler_ax0, handler_ax1): handler_ax0._backend.ifft2(input, output, handler_ax0, handler_ax1)
def rfft2(input, output, handler_ax0, handler_ax1): handler_ax0._backend.rfft2(input, output, handler_ax0, handler_ax1)
def irfft2(input, output, handler_ax0, handler_ax1): handler_ax0._backend.irfft2(input, output, handler_ax0, handler_ax1)


# `_par` twins (feature "parallel", Cargo.toml:39): on the GPU the serial entry points already run every lane in parallel.
ndfft_par, ndifft_par = ndfft, ndifft
ndfft_r2c_par, ndifft_r2c_par = ndfft_r2c, ndifft_r2c
nddct1_par, nddct2_par, nddct3_par, nddct4_par = nddct1, nddct2, nddct3, nddct4
