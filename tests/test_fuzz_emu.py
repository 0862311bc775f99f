"""Property-based layout/length fuzzing of every nd* function against the oracle (CPU: kernels under the SIMT emulator).

The reference accepts any `ArrayBase<_, D>` (any ndim, any strides: src/lib.rs:100-167 paths A/B/C), any length, any axis:
hypothesis draws (function, length, batch shape, axis, memory layout of input and output, normalisation, dtype)."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st, HealthCheck  # noqa: E402

from emu_backend import emu_backend  # noqa: E402
from parity_cases import Harness, TOL, cdt, seeded  # noqa: E402
from oracle import ndrustfft_oracle as orc  # noqa: E402

LENGTHS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 16, 17, 25, 31, 32, 36, 60, 64, 81, 97, 100, 128, 132, 210, 256, 264]
OPS = ["ndfft", "ndifft", "ndfft_r2c", "ndifft_r2c", "nddct1", "nddct2", "nddct3", "nddct4"]


@pytest.fixture(scope="module")
def hs():
    return Harness(emu_backend())


def relayout(a, how, rng):
    """Same values, different memory layout."""
    if how == "C":
        return np.ascontiguousarray(a)
    if how == "F":
        return np.asfortranarray(a)
    if how == "strided":        # every second element of a larger buffer along each dim
        big = np.zeros(tuple(2 * s for s in a.shape), a.dtype)
        view = big[tuple(slice(None, None, 2) for _ in a.shape)]
        view[...] = a
        return view
    if how == "reversed":       # negative strides along every dim
        buf = np.ascontiguousarray(a[tuple(slice(None, None, -1) for _ in a.shape)])
        return buf[tuple(slice(None, None, -1) for _ in a.shape)]
    if how == "transposed":     # permuted axes of a C buffer
        perm = list(rng.permutation(a.ndim))
        buf = np.ascontiguousarray(np.transpose(a, perm))
        return np.transpose(buf, np.argsort(perm))
    raise ValueError(how)


LAYOUTS = ["C", "F", "strided", "reversed", "transposed"]


@settings(max_examples=400, deadline=None, suppress_health_check=list(HealthCheck))
@given(op=st.sampled_from(OPS), n=st.sampled_from(LENGTHS), batch=st.lists(st.integers(1, 5), min_size=0, max_size=3),
       axis_pos=st.integers(0, 3), lin=st.sampled_from(LAYOUTS), lout=st.sampled_from(LAYOUTS),
       norm=st.sampled_from(["default", "none"]), f32=st.booleans(), seed=st.integers(0, 2 ** 16))
def test_any_layout_any_length(hs, op, n, batch, axis_pos, lin, lout, norm, f32, seed):
    if op == "nddct1" and n < 2:
        return
    rd = np.dtype(np.float32 if f32 else np.float64)
    axis = min(axis_pos, len(batch))
    shape = list(batch)
    shape.insert(axis, n)
    hk, icx, ocx = Harness.OPS[op]
    sin, sout = hs.shapes(op, n, tuple(shape), axis)
    rng = np.random.default_rng(seed)
    x = relayout(seeded(seed, sin, rd, icx), lin, rng)
    y = relayout(np.zeros(sout, cdt(rd) if ocx else rd), lout, rng)
    h = getattr(hs.be, hk)(n, rd)
    ho = getattr(orc, hk)(n)
    if norm == "none":
        h.normalization(type(h.norm).None_)
        ho.normalization(orc.Normalization.none())
    x0 = x.copy()
    getattr(hs.be, op)(x, y, h, axis)
    want = np.zeros(sout, np.complex128 if ocx else np.float64)
    getattr(orc, op)(np.asarray(x0), want, ho, axis)
    assert np.array_equal(x, x0), "input was modified"
    # tolerance: the north-star relative L2 bound; longer prime lengths in f32 go through Bluestein (two transforms)
    err = orc.rel_l2(np.asarray(y), want)
    assert err <= TOL[rd], f"{op} n={n} shape={shape} axis={axis} {rd} {lin}->{lout} norm={norm}: rel L2 {err:.3e}"
