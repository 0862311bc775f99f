"""CPU oracle for ndrustfft's axis-transform hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import this module.  The product package `ndrustfft_b200` never does: its transforms
run on hand-written sm_100a CUDA kernels behind the C ABI in `include/ndfft_b200.h` and fail
loudly when that library or a GPU is missing.

What this restates (citations are `/root/reference/...`, preiter93/ndrustfft v0.5.0):

* the lane drivers `create_transform!` / `create_transform_par!`          src/lib.rs:100-238
* `FftHandler::{fft_lane, ifft_lane, norm_default}`                       src/lib.rs:313-338
* `R2cFftHandler::{fft_r2c_lane, ifft_r2c_lane, norm_default}`            src/lib.rs:497-531
* `DctHandler::{dct1..4_lane, norm_default}`                              src/lib.rs:688-741
* the size assertions and their panic text                                src/lib.rs:340-347, 533-540, 743-750

The arithmetic itself lives in three crates that are NOT vendored under /root/reference and
cannot be built here (no rustc/cargo, no network): rustfft 6.1.0 (Cargo.lock:462-463),
realfft 3.2.0 (Cargo.lock:414-415), rustdct 0.7.0 (Cargo.lock:453-454).  Their published
definitions are restated here:

* rustfft  `Fft::process`: forward  X[k] = sum_j x[j] exp(-2 pi i jk/n), inverse with +, neither scaled.
* realfft  `RealToComplex::process`: bins 0..n/2 of the forward DFT of real input;
           `ComplexToReal::process`: unscaled inverse of that (output = n * numpy.irfft).
* rustdct  `process_dct1..4`: the textbook sums WITHOUT scipy's factor 2, i.e. exactly half of
           `scipy.fft.dct(x, type=k, norm=None)`.

Engines: `scipy.fft` (pocketfft, f64) for speed, and `naive_*` O(n^2) longdouble sums of the
definitions above as an independent check of pocketfft (tests/test_oracle.py).

Parity pinning: tests/test_oracle.py checks this module against every golden vector the reference
holds for the path (tests/golden/reference_goldens.json, extracted from src/lib.rs:880-1380 and
examples/{fft2,rfft2,fft_norm}.rs by tests/golden/extract_goldens.py).  Those goldens are 6x6 / n=3
f64 cases at abs 1e-3 / 1e-4; beyond them (f32, large n, path B, None/Custom norms for r2c and DCT)
the reference pins nothing, so parity there rests on the mathematical definitions above.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.fft as _sfft

__all__ = [
    "Normalization", "FftHandler", "R2cFftHandler", "DctHandler",
    "ndfft", "ndifft", "ndfft_r2c", "ndifft_r2c", "nddct1", "nddct2", "nddct3", "nddct4",
    "ndfft_par", "ndifft_par", "ndfft_r2c_par", "ndifft_r2c_par",
    "nddct1_par", "nddct2_par", "nddct3_par", "nddct4_par",
    "naive_dft", "naive_dct", "rel_l2",
]


# --------------------------------------------------------------------------------------
# Normalization (src/lib.rs:89-98)
# --------------------------------------------------------------------------------------
class Normalization:
    """`enum Normalization<T> { None, Default, Custom(fn(&mut [T])) }` (src/lib.rs:89-98)."""

    NONE = "none"
    DEFAULT = "default"

    def __init__(self, kind, func=None):
        self.kind = kind
        self.func = func

    @classmethod
    def none(cls):
        return cls(cls.NONE)

    @classmethod
    def default(cls):
        return cls(cls.DEFAULT)

    @classmethod
    def custom(cls, func):
        """`func(lane)` mutates a 1-D numpy array in place, like `fn(&mut [T])`."""
        return cls("custom", func)


Normalization.None_ = Normalization(Normalization.NONE)
Normalization.Default = Normalization(Normalization.DEFAULT)
Normalization.Custom = Normalization.custom


# Precision of the engine calls.  Default: always f64 (the oracle is the yardstick for the f32 kernels too).
# bench.py's CPU-baseline legs switch to the handler's own precision so that the timed port does the same
# arithmetic the reference would (rustfft on Complex<f32>), not an f64 detour.
_NATIVE = False


def set_native_precision(flag):
    global _NATIVE
    _NATIVE = bool(flag)


def _rdt(h):
    return np.float32 if (_NATIVE and h.dtype == np.float32) else np.float64


def _cdt(h):
    return np.complex64 if (_NATIVE and h.dtype == np.float32) else np.complex128


def _workers(par):
    return (os.cpu_count() or 1) if par else 1


def _apply_custom(func, arr, axis):
    """Call `func` on every lane of `arr` along `axis` (in place), as the per-lane closure does."""
    moved = np.moveaxis(arr, axis, -1)
    it = moved.reshape(-1, moved.shape[-1]) if moved.flags.c_contiguous else None
    if it is None:
        tmp = np.ascontiguousarray(moved)
        flat = tmp.reshape(-1, tmp.shape[-1])
        for lane in flat:
            func(lane)
        moved[...] = tmp
    else:
        for lane in it:
            func(lane)


# --------------------------------------------------------------------------------------
# Handlers
# --------------------------------------------------------------------------------------
class FftHandler:
    """`FftHandler<T>` (src/lib.rs:270-348).  `dtype` is the real type T (np.float32 / np.float64)."""

    def __init__(self, n, dtype=np.float64):
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        self.norm = Normalization.Default  # src/lib.rs:302

    def normalization(self, norm):  # src/lib.rs:308-311
        self.norm = norm
        return self

    def assert_size(self, size):  # src/lib.rs:340-347
        assert self.n == size, f"Size mismatch in fft, got {size} expected {self.n}"


class R2cFftHandler:
    """`R2cFftHandler<T>` (src/lib.rs:452-541)."""

    def __init__(self, n, dtype=np.float64):
        self.n = int(n)
        self.m = self.n // 2 + 1  # src/lib.rs:483
        self.dtype = np.dtype(dtype)
        self.norm = Normalization.Default  # src/lib.rs:486

    def normalization(self, norm):  # src/lib.rs:492-495
        self.norm = norm
        return self


class DctHandler:
    """`DctHandler<T>` (src/lib.rs:641-751)."""

    def __init__(self, n, dtype=np.float64):
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        self.norm = Normalization.Default  # src/lib.rs:677

    def normalization(self, norm):  # src/lib.rs:683-686
        self.norm = norm
        return self


def _check_lane(kind, got, expected):
    # src/lib.rs:340-347 / 533-540 ("fft") and 743-750 ("dct"): got first, expected second.
    assert got == expected, f"Size mismatch in {kind}, got {got} expected {expected}"


def _check_shapes(inp, out, axis):
    # ndarray's Zip panics when the non-transformed dimensions differ (src/lib.rs:120-163).
    if inp.ndim != out.ndim:
        raise AssertionError("ndarray: dimension mismatch")
    if not (0 <= axis < inp.ndim):
        raise IndexError("axis out of range")  # index panic at src/lib.rs:116
    for d in range(inp.ndim):
        if d != axis and inp.shape[d] != out.shape[d]:
            raise AssertionError("ndarray: could not zip arrays of different shapes")


# --------------------------------------------------------------------------------------
# The eight transforms.  Computation is always carried out in f64 (the oracle is the yardstick
# for both the f64 and the f32 kernels) and cast to `output.dtype` on assignment.
# --------------------------------------------------------------------------------------
def _c2c(inp, out, h, axis, inverse, par):
    _check_shapes(inp, out, axis)
    _check_lane("fft", inp.shape[axis], h.n)   # src/lib.rs:314 / 322
    _check_lane("fft", out.shape[axis], h.n)   # src/lib.rs:315 / 323
    x = np.asarray(inp, dtype=_cdt(h))
    if not inverse:
        y = _sfft.fft(x, axis=axis, workers=_workers(par))  # src/lib.rs:316-317
    else:
        # plan_bwd.process is unscaled (src/lib.rs:324-325): numpy's ifft * n
        y = _sfft.ifft(x, axis=axis, norm="forward", workers=_workers(par))
        if h.norm.kind == Normalization.DEFAULT:      # src/lib.rs:328, 333-338
            y *= y.dtype.type(1.0 / h.n).real
        elif h.norm.kind == "custom":                 # src/lib.rs:329 (after the transform)
            y = np.array(y)
            _apply_custom(h.norm.func, y, axis)
    out[...] = y


def _r2c(inp, out, h, axis, par):
    _check_shapes(inp, out, axis)
    _check_lane("fft", inp.shape[axis], h.n)   # src/lib.rs:498
    _check_lane("fft", out.shape[axis], h.m)   # src/lib.rs:499
    x = np.asarray(inp, dtype=_rdt(h))
    out[...] = _sfft.rfft(x, axis=axis, workers=_workers(par))  # src/lib.rs:500-502, never scaled


def _c2r(inp, out, h, axis, par):
    _check_shapes(inp, out, axis)
    _check_lane("fft", inp.shape[axis], h.m)   # src/lib.rs:507
    _check_lane("fft", out.shape[axis], h.n)   # src/lib.rs:508
    buf = np.array(inp, dtype=_cdt(h))         # src/lib.rs:509-510 (copy)
    if h.norm.kind == Normalization.DEFAULT:   # src/lib.rs:513, 525-531: spectrum * 1/n BEFORE the transform
        buf *= 1.0 / h.n
    elif h.norm.kind == "custom":              # src/lib.rs:514: f sees the m-long spectrum copy
        _apply_custom(h.norm.func, buf, axis)
    sl = [slice(None)] * buf.ndim
    sl[axis] = 0
    buf[tuple(sl)] = buf[tuple(sl)].real       # src/lib.rs:517  buffer[0].im = 0
    if h.n % 2 == 0:                           # src/lib.rs:519-521
        sl[axis] = h.m - 1
        buf[tuple(sl)] = buf[tuple(sl)].real
    # realfft's ComplexToReal is unscaled: n * irfft (src/lib.rs:522)
    out[...] = _sfft.irfft(buf, n=h.n, axis=axis, norm="forward", workers=_workers(par))


def _dct(inp, out, h, axis, kind, par):
    _check_shapes(inp, out, axis)
    _check_lane("dct", inp.shape[axis], h.n)   # src/lib.rs:689 ...
    _check_lane("dct", out.shape[axis], h.n)   # src/lib.rs:690 ...
    buf = np.array(inp, dtype=_rdt(h))         # src/lib.rs:691 (copy)
    if h.norm.kind == Normalization.DEFAULT:   # src/lib.rs:694, 736-741: input * 2 BEFORE the transform
        buf *= 2.0
    elif h.norm.kind == "custom":              # src/lib.rs:695
        _apply_custom(h.norm.func, buf, axis)
    # rustdct = half of scipy's unnormalised DCT (src/lib.rs:697/709/721/733)
    out[...] = 0.5 * _sfft.dct(buf, type=kind, axis=axis, norm=None, workers=_workers(par))


def ndfft(inp, out, handler, axis):            # src/lib.rs:350-372
    _c2c(inp, out, handler, axis, False, False)


def ndifft(inp, out, handler, axis):           # src/lib.rs:374-397
    _c2c(inp, out, handler, axis, True, False)


def ndfft_par(inp, out, handler, axis):        # src/lib.rs:399-409
    _c2c(inp, out, handler, axis, False, True)


def ndifft_par(inp, out, handler, axis):       # src/lib.rs:411-421
    _c2c(inp, out, handler, axis, True, True)


def ndfft_r2c(inp, out, handler, axis):        # src/lib.rs:543-564
    _r2c(inp, out, handler, axis, False)


def ndifft_r2c(inp, out, handler, axis):       # src/lib.rs:566-587
    _c2r(inp, out, handler, axis, False)


def ndfft_r2c_par(inp, out, handler, axis):    # src/lib.rs:589-599
    _r2c(inp, out, handler, axis, True)


def ndifft_r2c_par(inp, out, handler, axis):   # src/lib.rs:601-611
    _c2r(inp, out, handler, axis, True)


def nddct1(inp, out, handler, axis):           # src/lib.rs:753-775
    _dct(inp, out, handler, axis, 1, False)


def nddct2(inp, out, handler, axis):           # src/lib.rs:789-796
    _dct(inp, out, handler, axis, 2, False)


def nddct3(inp, out, handler, axis):           # src/lib.rs:808-815
    _dct(inp, out, handler, axis, 3, False)


def nddct4(inp, out, handler, axis):           # src/lib.rs:827-834
    _dct(inp, out, handler, axis, 4, False)


def nddct1_par(inp, out, handler, axis):       # src/lib.rs:777-787
    _dct(inp, out, handler, axis, 1, True)


def nddct2_par(inp, out, handler, axis):       # src/lib.rs:798-806
    _dct(inp, out, handler, axis, 2, True)


def nddct3_par(inp, out, handler, axis):       # src/lib.rs:817-825
    _dct(inp, out, handler, axis, 3, True)


def nddct4_par(inp, out, handler, axis):       # src/lib.rs:836-844
    _dct(inp, out, handler, axis, 4, True)


# --------------------------------------------------------------------------------------
# Independent O(n^2) definitions in extended precision (pins pocketfft itself)
# --------------------------------------------------------------------------------------
def naive_dft(x, inverse=False):
    """Unscaled DFT of a 1-D array by direct summation in longdouble (rustfft's definition).

    The phase is reduced exactly with integer arithmetic (jk mod n) before the trig call."""
    x = np.asarray(x)
    n = x.shape[0]
    j = np.arange(n)
    jk = (j[:, None] * j[None, :]) % n
    ang = (2 * np.longdouble(np.pi)) * jk.astype(np.longdouble) / np.longdouble(n)
    c, s = np.cos(ang), np.sin(ang)
    xr = np.real(x).astype(np.longdouble)
    xi = np.imag(x).astype(np.longdouble)
    sg = 1 if inverse else -1
    re = c @ xr - sg * (s @ xi)
    im = sg * (s @ xr) + c @ xi
    return (re + 1j * im).astype(np.complex128)


def naive_dct(x, kind):
    """rustdct's DCT-I..IV definitions (no factor 2) by direct summation in longdouble."""
    x = np.asarray(x, dtype=np.longdouble)
    n = x.shape[0]
    pi = np.longdouble(np.pi)
    j = np.arange(n).astype(np.longdouble)
    k = j[:, None]
    if kind == 1:
        if n == 1:
            return np.array([x[0]], dtype=np.float64)
        y = np.cos(pi * k * j[None, 1:-1] / (n - 1)) @ x[1:-1] if n > 2 else np.zeros(n, np.longdouble)
        y = y + x[0] / 2 + np.where(np.arange(n) % 2 == 0, 1, -1) * x[-1] / 2
    elif kind == 2:
        y = np.cos(pi * k * (j[None, :] + 0.5) / n) @ x
    elif kind == 3:
        y = np.cos(pi * (k + 0.5) * j[None, 1:] / n) @ x[1:] + x[0] / 2
    elif kind == 4:
        y = np.cos(pi * (k + 0.5) * (j[None, :] + 0.5) / n) @ x
    else:
        raise ValueError(kind)
    return np.asarray(y, dtype=np.float64)


def rel_l2(got, want):
    """Relative L2 error ||got - want|| / ||want|| in f64 (the north-star tolerance metric)."""
    got = np.asarray(got)
    want = np.asarray(want)
    d = np.linalg.norm((got.astype(np.complex128) - want.astype(np.complex128)).ravel())
    w = np.linalg.norm(want.astype(np.complex128).ravel())
    return float(d / w) if w > 0 else float(d)
