"""Pins the CPU oracle (oracle/ndrustfft_oracle.py) against every golden vector the reference holds
for the path (tests/golden/reference_goldens.json <- /root/reference/src/lib.rs:880-1380, examples/),
and against independent O(n^2) longdouble definitions."""
import json
import os

import numpy as np
import pytest

from oracle import ndrustfft_oracle as orc

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))
TOL = G["_tolerance_abs"]  # approx_eq: abs 1e-3 (src/lib.rs:852-878)


def tm():
    return np.array(G["test_matrix"]["values"], dtype=np.float64)


def approx_eq(result, expected, tol=TOL):
    assert np.max(np.abs(np.asarray(result) - np.asarray(expected))) <= tol


@pytest.mark.parametrize("par", [False, True])
def test_fft(par):  # src/lib.rs:903-994
    sol = np.array(G["test_fft"]["solution_re"]) + 1j * np.array(G["test_fft"]["solution_im"])
    v = tm() * (1 + 1j)
    v0 = v.copy()
    vhat = np.zeros_like(v)
    h = orc.FftHandler(6)
    (orc.ndfft_par if par else orc.ndfft)(v, vhat, h, 1)
    (orc.ndifft_par if par else orc.ndifft)(vhat, v, h, 1)
    approx_eq(vhat, sol)
    approx_eq(v, v0)


def test_fft_f_layout():  # src/lib.rs:996-1040
    sol = np.array(G["test_fft_f_layout"]["solution_re"]) + 1j * np.array(G["test_fft_f_layout"]["solution_im"])
    v = np.asfortranarray(tm() * (1 + 1j))
    v0 = v.copy()
    vhat = np.zeros((6, 6), np.complex128)
    h = orc.FftHandler(6)
    orc.ndfft(v, vhat, h, 1)
    orc.ndifft(vhat, v, h, 1)
    approx_eq(vhat, sol)
    approx_eq(v, v0)


@pytest.mark.parametrize("par", [False, True])
def test_fft_r2c(par):  # src/lib.rs:1042-1133
    sol = np.array(G["test_fft_r2c"]["solution_re"]) + 1j * np.array(G["test_fft_r2c"]["solution_im"])
    v = tm()
    v0 = v.copy()
    vhat = np.zeros((6, 4), np.complex128)
    h = orc.R2cFftHandler(6)
    (orc.ndfft_r2c_par if par else orc.ndfft_r2c)(v, vhat, h, 1)
    (orc.ndifft_r2c_par if par else orc.ndifft_r2c)(vhat, v, h, 1)
    approx_eq(vhat, sol)
    approx_eq(v, v0)


def test_ifft_c2r_first_last_element():  # src/lib.rs:1135-1167
    g = G["test_ifft_c2r_first_last_element"]
    h = orc.R2cFftHandler(6)
    v = np.zeros(6)
    vhat = np.zeros(4, np.complex128)
    vhat[0] = 1 + 100j
    orc.ndifft_r2c(vhat, v, h, 0)
    approx_eq(v, g["solution_numpy_first_elem"])
    vhat[:] = 0
    vhat[3] = 1 + 100j
    orc.ndifft_r2c(vhat, v, h, 0)
    approx_eq(v, g["solution_numpy_last_elem"])


def test_fft_r2c_odd():  # src/lib.rs:1169-1202
    v = np.array(G["test_fft_r2c_odd"]["v"], dtype=np.float64)
    v0 = v.copy()
    vhat = np.zeros((3, 2), np.complex128)
    h = orc.R2cFftHandler(3)
    orc.ndfft_r2c(v, vhat, h, 1)
    orc.ndifft_r2c(vhat, v, h, 1)
    approx_eq(v, v0)


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
@pytest.mark.parametrize("par", [False, True])
def test_dct(kind, par):  # src/lib.rs:1204-1406
    sol = np.array(G[f"test_dct{kind}"]["solution"])
    v = tm()
    vhat = np.zeros_like(v)
    h = orc.DctHandler(6)
    getattr(orc, f"nddct{kind}" + ("_par" if par else ""))(v, vhat, h, 1)
    approx_eq(vhat, sol)


def test_example_fft2():  # examples/fft2.rs:14-66
    g = G["example_fft2"]
    v = np.array(g["input_real"]) * (1 + 1j)
    work = np.zeros_like(v)
    vhat = np.zeros_like(v)
    h0, h1 = orc.FftHandler(3), orc.FftHandler(3)
    orc.ndfft(v, work, h1, 1)
    orc.ndfft(work, vhat, h0, 0)      # path B (axis != last, standard layout)
    want = np.array(g["numpy_vhat"])
    approx_eq(vhat, want[..., 0] + 1j * want[..., 1], g["tol"])
    v2 = np.zeros_like(v)
    orc.ndifft(vhat, work, h0, 0)
    orc.ndifft(work, v2, h1, 1)
    approx_eq(v2, v, g["tol"])


def test_example_rfft2():  # examples/rfft2.rs:22-60
    g = G["example_rfft2"]
    v = np.array(g["input_real"])
    work = np.zeros((3, 2), np.complex128)
    vhat = np.zeros((3, 2), np.complex128)
    h0, h1 = orc.FftHandler(3), orc.R2cFftHandler(3)
    orc.ndfft_r2c(v, work, h1, 1)
    orc.ndfft(work, vhat, h0, 0)
    want = np.array(g["numpy_vhat"])
    approx_eq(vhat, want[..., 0] + 1j * want[..., 1], g["tol"])
    v2 = np.zeros_like(v)
    orc.ndifft(vhat, work, h0, 0)
    orc.ndifft_r2c(work, v2, h1, 1)
    approx_eq(v2, v, g["tol"])


def test_example_fft_norm():  # examples/fft_norm.rs:17-40
    g = G["example_fft_norm"]
    v = np.array(g["input_real"]) * (1 + 1j)

    def my_norm(data):
        data *= 2.0 / len(data)

    for norm, key in ((orc.Normalization.default(), "default_roundtrip"),
                      (orc.Normalization.none(), "none_roundtrip"),
                      (orc.Normalization.custom(my_norm), "custom_2_over_len_roundtrip")):
        h = orc.FftHandler(3).normalization(norm)
        vhat = np.zeros_like(v)
        v2 = np.zeros_like(v)
        orc.ndfft(v, vhat, h, 0)
        orc.ndifft(vhat, v2, h, 0)
        approx_eq(v2, np.array(g[key]) * (1 + 1j), 1e-12)


def test_size_mismatch_message():  # src/lib.rs:340-347, 743-750
    with pytest.raises(AssertionError, match="Size mismatch in fft, got 5 expected 6"):
        orc.ndfft(np.zeros((2, 5), complex), np.zeros((2, 5), complex), orc.FftHandler(6), 1)
    with pytest.raises(AssertionError, match="Size mismatch in dct, got 5 expected 6"):
        orc.nddct2(np.zeros((2, 5)), np.zeros((2, 5)), orc.DctHandler(6), 1)


# ---- pocketfft vs the O(n^2) longdouble definitions (independent pin of the engine) -------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 16, 30, 97, 128, 257])
def test_engine_vs_naive(n):
    rng = np.random.default_rng(1000 + n)
    x = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    h = orc.FftHandler(n).normalization(orc.Normalization.none())
    y = np.zeros(n, complex)
    orc.ndfft(x, y, h, 0)
    assert orc.rel_l2(y, orc.naive_dft(x)) < 1e-13
    orc.ndifft(x, y, h, 0)
    assert orc.rel_l2(y, orc.naive_dft(x, inverse=True)) < 1e-13
    xr = rng.uniform(-1, 1, n)
    hr = orc.R2cFftHandler(n)
    yr = np.zeros(n // 2 + 1, complex)
    orc.ndfft_r2c(xr, yr, hr, 0)
    assert orc.rel_l2(yr, orc.naive_dft(xr)[: n // 2 + 1]) < 1e-13
    hd = orc.DctHandler(n).normalization(orc.Normalization.none())
    for kind in (1, 2, 3, 4):
        if kind == 1 and n < 2:
            continue
        yd = np.zeros(n)
        getattr(orc, f"nddct{kind}")(xr, yd, hd, 0)
        assert orc.rel_l2(yd, orc.naive_dct(xr, kind)) < 1e-13, kind


def test_roundtrip_factors():
    rng = np.random.default_rng(7)
    n = 20
    x = rng.uniform(-1, 1, (3, n))
    h = orc.DctHandler(n)
    a, b = np.zeros_like(x), np.zeros_like(x)
    orc.nddct2(x, a, h, 1); orc.nddct3(a, b, h, 1)
    assert orc.rel_l2(b, 2 * n * x) < 1e-13
    orc.nddct1(x, a, h, 1); orc.nddct1(a, b, h, 1)
    assert orc.rel_l2(b, 2 * (n - 1) * x) < 1e-13
    orc.nddct4(x, a, h, 1); orc.nddct4(a, b, h, 1)
    assert orc.rel_l2(b, 2 * n * x) < 1e-13
