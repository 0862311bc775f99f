"""The numpy model of the kernel math (tests/kernel_math_model.py) against the oracle's definitions."""
import numpy as np
import pytest

import kernel_math_model as km
from oracle import ndrustfft_oracle as orc

NS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 17, 30, 31, 32, 33, 64, 97, 100, 126, 128, 129, 130, 360]


@pytest.mark.parametrize("n", NS)
def test_model_all_kinds(n):
    rng = np.random.default_rng(n)
    xc = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    xr = rng.uniform(-1, 1, n)
    none = orc.Normalization.none()
    y = np.zeros(n, complex)
    orc.ndfft(xc, y, orc.FftHandler(n), 0)
    assert orc.rel_l2(km.c2c(xc, False), y) < 1e-12
    orc.ndifft(xc, y, orc.FftHandler(n).normalization(none), 0)
    assert orc.rel_l2(km.c2c(xc, True), y) < 1e-12
    m = n // 2 + 1
    yr = np.zeros(m, complex)
    orc.ndfft_r2c(xr, yr, orc.R2cFftHandler(n), 0)
    assert orc.rel_l2(km.r2c(xr), yr) < 1e-12
    spec = rng.uniform(-1, 1, m) + 1j * rng.uniform(-1, 1, m)
    back = np.zeros(n)
    orc.ndifft_r2c(spec, back, orc.R2cFftHandler(n).normalization(none), 0)
    assert orc.rel_l2(km.c2r(spec, n), back) < 1e-12
    hd = orc.DctHandler(n).normalization(none)
    for kind, f in ((1, km.dct1), (2, km.dct2), (3, km.dct3), (4, km.dct4)):
        if kind == 1 and n < 2:
            continue
        yd = np.zeros(n)
        getattr(orc, f"nddct{kind}")(xr, yd, hd, 0)
        assert orc.rel_l2(f(xr), yd) < 1e-12, (kind, n)
