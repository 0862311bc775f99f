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


@settings(max_examples=400, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
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


# ---- multi-axis chains: random sequences of steps over random axes and layouts against the step-by-step oracle ----
@settings(max_examples=150, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
@given(data=st.data())
def test_any_chain(hs, data):
    ndim = data.draw(st.integers(1, 3))
    shape = [data.draw(st.sampled_from([2, 3, 4, 5, 6, 8, 9, 12, 16])) for _ in range(ndim)]
    rd = np.dtype(data.draw(st.sampled_from([np.float32, np.float64])))
    nsteps = data.draw(st.integers(1, 4))
    cplx = data.draw(st.booleans())          # element type of the array the chain starts from
    cur_shape = list(shape)
    steps = []
    start_cplx = cplx
    for _ in range(nsteps):
        axis = data.draw(st.integers(0, ndim - 1))
        n_here = cur_shape[axis]
        if cplx:
            ops = ["ndfft", "ndifft"]
            # c2r along this axis: n_here = m = n/2 + 1  ->  n in {2m-2, 2m-1}
            if n_here >= 2:
                ops.append("ndifft_r2c")
        else:
            ops = ["ndfft_r2c", "nddct2", "nddct3", "nddct4"] + (["nddct1"] if n_here >= 2 else [])
        op = data.draw(st.sampled_from(ops))
        if op == "ndifft_r2c":
            n = data.draw(st.sampled_from([2 * n_here - 2, 2 * n_here - 1]))
            if n < 1:
                n = 1
            cur_shape[axis] = n
            cplx = False
        elif op == "ndfft_r2c":
            n = n_here
            cur_shape[axis] = n // 2 + 1
            cplx = True
        else:
            n = n_here
        steps.append((op, n, axis))
    lin = data.draw(st.sampled_from(LAYOUTS)); lout = data.draw(st.sampled_from(LAYOUTS))
    norm = data.draw(st.sampled_from(["default", "none"]))
    seed = data.draw(st.integers(0, 2 ** 16))
    rng = np.random.default_rng(seed)
    x = relayout(seeded(seed, tuple(shape), rd, start_cplx), lin, rng)
    hs_steps, cur = [], np.asarray(x).astype(np.complex128 if start_cplx else np.float64)
    for op, n, axis in steps:
        hk, _, ocx = Harness.OPS[op]
        h = getattr(hs.be, hk)(n, rd)
        ho = getattr(orc, hk)(n)
        if norm == "none":
            h.normalization(type(h.norm).None_)
            ho.normalization(orc.Normalization.none())
        hs_steps.append((op, h, axis))
        _, sout = hs.shapes(op, n, cur.shape, axis)
        nxt = np.zeros(sout, np.complex128 if ocx else np.float64)
        getattr(orc, op)(cur, nxt, ho, axis)
        cur = nxt
    y = relayout(np.zeros(cur.shape, cdt(rd) if cplx else rd), lout, rng)
    x0 = x.copy()
    hs.be.ndchain(x, y, hs_steps)
    assert np.array_equal(x, x0), "input was modified"
    scale = max(1.0, float(np.linalg.norm(cur)))
    err = float(np.linalg.norm(np.asarray(y) - cur)) / scale
    assert err <= 4 * TOL[rd], f"chain {steps} shape={shape} {rd} {lin}->{lout} norm={norm}: rel L2 {err:.3e}"
