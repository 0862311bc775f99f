"""Parity cases shared by the CPU (SIMT-emulation) and GPU test modules.

Every function takes a `Backend` (ndrustfft_b200.Backend bound to one loaded C library) and an array factory
`mk(np_array) -> array` (identity for host arrays, torch.cuda tensor for the device path) plus `to_np`.
Comparison is against oracle/ndrustfft_oracle.py; tolerances are the north-star ones:
relative L2 <= 1e-12 (f64) / 1e-5 (f32).
"""
import json
import os

import numpy as np

from oracle import ndrustfft_oracle as orc

TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")))


def cdt(rd):
    return np.complex64 if np.dtype(rd) == np.float32 else np.complex128


def seeded(seed, shape, rd, complex_):
    rng = np.random.default_rng(seed)
    if complex_:
        return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(cdt(rd))
    return rng.uniform(-1, 1, shape).astype(rd)


class Harness:
    def __init__(self, be, mk=None, to_np=None, zeros=None):
        self.be = be
        self.mk = mk or (lambda a: np.array(a))
        self.to_np = to_np or (lambda a: np.asarray(a))
        self.zeros = zeros or (lambda shape, dt: np.zeros(shape, dt))

    # op name -> (backend fn, oracle fn, handler kind, in complex, out complex)
    OPS = {
        "ndfft": ("FftHandler", True, True),
        "ndifft": ("FftHandler", True, True),
        "ndfft_r2c": ("R2cFftHandler", False, True),
        "ndifft_r2c": ("R2cFftHandler", True, False),
        "nddct1": ("DctHandler", False, False),
        "nddct2": ("DctHandler", False, False),
        "nddct3": ("DctHandler", False, False),
        "nddct4": ("DctHandler", False, False),
    }

    def shapes(self, op, n, shape, axis):
        m = n // 2 + 1
        sin, sout = list(shape), list(shape)
        sin[axis] = m if op == "ndifft_r2c" else n
        sout[axis] = m if op == "ndfft_r2c" else n
        return tuple(sin), tuple(sout)

    def run(self, op, n, shape, axis, rd=np.float64, norm="default", seed=0, order="C", par=False, tol=None):
        """One call of `op` on seeded data vs the oracle; returns the relative L2 error."""
        hk, icx, ocx = self.OPS[op]
        rd = np.dtype(rd)
        sin, sout = self.shapes(op, n, shape, axis)
        x = seeded(seed, sin, rd, icx)
        if order == "F":
            x = np.asfortranarray(x)
        h = getattr(self.be, hk)(n, rd)
        ho = getattr(orc, hk)(n)
        if norm == "none":
            h.normalization(type(h.norm).None_)
            ho.normalization(orc.Normalization.none())
        name = op + ("_par" if par else "")
        xin = self.mk(x)
        y = self.zeros(sout, cdt(rd) if ocx else rd)
        getattr(self.be, name)(xin, y, h, axis)
        yo = np.zeros(sout, np.complex128 if ocx else np.float64)
        getattr(orc, name)(x, yo, ho, axis)
        err = orc.rel_l2(self.to_np(y), yo)
        assert np.array_equal(self.to_np(xin), x), "input was modified"
        t = TOL[rd] if tol is None else tol
        assert err <= t, f"{name} n={n} shape={shape} axis={axis} {rd} norm={norm}: rel L2 {err:.3e} > {t:g}"
        return err

    # ---- the reference's own unit tests, transliterated (src/lib.rs:903-1406, examples/) ----
    def reference_unit_tests(self):
        be = self.be
        tol = G["_tolerance_abs"]
        tm = np.array(G["test_matrix"]["values"])

        def approx(a, b, t=tol):
            assert np.max(np.abs(self.to_np(a) - np.asarray(b))) <= t

        for par in ("", "_par"):
            # test_fft / test_fft_par
            sol = np.array(G["test_fft"]["solution_re"]) + 1j * np.array(G["test_fft"]["solution_im"])
            v = self.mk(tm * (1 + 1j)); vhat = self.zeros((6, 6), np.complex128)
            h = be.FftHandler(6)
            getattr(be, "ndfft" + par)(v, vhat, h, 1)
            v2 = self.zeros((6, 6), np.complex128)
            getattr(be, "ndifft" + par)(vhat, v2, h, 1)
            approx(vhat, sol); approx(v2, tm * (1 + 1j))
            # test_fft_r2c
            sol = np.array(G["test_fft_r2c"]["solution_re"]) + 1j * np.array(G["test_fft_r2c"]["solution_im"])
            v = self.mk(tm); vhat = self.zeros((6, 4), np.complex128); back = self.zeros((6, 6), np.float64)
            h = be.R2cFftHandler(6)
            getattr(be, "ndfft_r2c" + par)(v, vhat, h, 1)
            getattr(be, "ndifft_r2c" + par)(vhat, back, h, 1)
            approx(vhat, sol); approx(back, tm)
            # test_fft_r2c_odd
            vo = np.array(G["test_fft_r2c_odd"]["v"], dtype=np.float64)
            v = self.mk(vo); vhat = self.zeros((3, 2), np.complex128); back = self.zeros((3, 3), np.float64)
            h = be.R2cFftHandler(3)
            getattr(be, "ndfft_r2c" + par)(v, vhat, h, 1)
            getattr(be, "ndifft_r2c" + par)(vhat, back, h, 1)
            approx(back, vo)
            # test_dct1..4
            for k in (1, 2, 3, 4):
                v = self.mk(tm); vhat = self.zeros((6, 6), np.float64)
                h = be.DctHandler(6)
                getattr(be, f"nddct{k}" + par)(v, vhat, h, 1)
                approx(vhat, np.array(G[f"test_dct{k}"]["solution"]))
        # test_fft_f_layout: F-order input, C-order output (path C)
        sol = np.array(G["test_fft_f_layout"]["solution_re"]) + 1j * np.array(G["test_fft_f_layout"]["solution_im"])
        vf = self.mk_f(tm * (1 + 1j))
        vhat = self.zeros((6, 6), np.complex128)
        h = be.FftHandler(6)
        be.ndfft(vf, vhat, h, 1)
        be.ndifft(vhat, vf, h, 1)
        approx(vhat, sol); approx(vf, tm * (1 + 1j))
        # test_ifft_c2r_first_last_element
        g = G["test_ifft_c2r_first_last_element"]
        h = be.R2cFftHandler(6)
        spec = np.zeros(4, np.complex128); spec[0] = 1 + 100j
        v = self.zeros((6,), np.float64)
        be.ndifft_r2c(self.mk(spec), v, h, 0)
        approx(v, g["solution_numpy_first_elem"])
        spec[:] = 0; spec[3] = 1 + 100j
        be.ndifft_r2c(self.mk(spec), v, h, 0)
        approx(v, g["solution_numpy_last_elem"])
        # examples/fft2.rs, rfft2.rs
        g = G["example_fft2"]
        v = self.mk(np.array(g["input_real"]) * (1 + 1j))
        work = self.zeros((3, 3), np.complex128); vhat = self.zeros((3, 3), np.complex128)
        h0, h1 = be.FftHandler(3), be.FftHandler(3)
        be.ndfft(v, work, h1, 1); be.ndfft(work, vhat, h0, 0)
        want = np.array(g["numpy_vhat"])
        approx(vhat, want[..., 0] + 1j * want[..., 1], g["tol"])
        v2 = self.zeros((3, 3), np.complex128)
        be.ndifft(vhat, work, h0, 0); be.ndifft(work, v2, h1, 1)
        approx(v2, np.array(g["input_real"]) * (1 + 1j), g["tol"])
        g = G["example_rfft2"]
        v = self.mk(np.array(g["input_real"]))
        work = self.zeros((3, 2), np.complex128); vhat = self.zeros((3, 2), np.complex128)
        h0, h1 = be.FftHandler(3), be.R2cFftHandler(3)
        be.ndfft_r2c(v, work, h1, 1); be.ndfft(work, vhat, h0, 0)
        want = np.array(g["numpy_vhat"])
        approx(vhat, want[..., 0] + 1j * want[..., 1], g["tol"])
        v2 = self.zeros((3, 3), np.float64)
        be.ndifft(vhat, work, h0, 0); be.ndifft_r2c(work, v2, h1, 1)
        approx(v2, np.array(g["input_real"]), g["tol"])
        # examples/fft_norm.rs
        g = G["example_fft_norm"]
        Norm = type(be.FftHandler(3).norm)

        def my_norm(data):
            data *= 2.0 / len(data)

        for norm, key in ((Norm.Default, "default_roundtrip"), (Norm.None_, "none_roundtrip"),
                          (Norm.Custom(my_norm), "custom_2_over_len_roundtrip")):
            h = be.FftHandler(3).normalization(norm)
            v = self.mk(np.array(g["input_real"]) * (1 + 1j))
            vhat = self.zeros((3,), np.complex128); v2 = self.zeros((3,), np.complex128)
            be.ndfft(v, vhat, h, 0); be.ndifft(vhat, v2, h, 0)
            approx(v2, np.array(g[key]) * (1 + 1j), 1e-12)

    # ---- multi-axis chains (ndfb_exec_chain) vs the same steps through the oracle, one call per axis ----
    def run_chain(self, steps, shape_in, rd=np.float64, norm="default", seed=0, inplace=False, tol=None):
        """steps = [(op name, n, axis)]; returns the relative L2 error against the oracle's step-by-step result."""
        rd = np.dtype(rd)
        icx = self.OPS[steps[0][0]][1]
        x = seeded(seed, shape_in, rd, icx)
        hs, cur = [], x.astype(np.complex128 if icx else np.float64)
        for op, n, axis in steps:
            hk, _, ocx = self.OPS[op]
            h = getattr(self.be, hk)(n, rd)
            ho = getattr(orc, hk)(n)
            if norm == "none":
                h.normalization(type(h.norm).None_)
                ho.normalization(orc.Normalization.none())
            hs.append((op, h, axis))
            _, sout = self.shapes(op, n, cur.shape, axis)
            nxt = np.zeros(sout, np.complex128 if ocx else np.float64)
            getattr(orc, op)(cur, nxt, ho, axis)
            cur = nxt
        ocx = self.OPS[steps[-1][0]][2]
        xin = self.mk(x)
        y = xin if inplace else self.zeros(cur.shape, cdt(rd) if ocx else rd)
        self.be.ndchain(xin, y, hs)
        err = orc.rel_l2(self.to_np(y), cur)
        if not inplace:
            assert np.array_equal(self.to_np(xin), x), "input was modified"
        t = TOL[rd] if tol is None else tol
        assert err <= t, f"chain {steps} shape={shape_in} {rd} norm={norm}: rel L2 {err:.3e} > {t:g}"
        return err

    def chain_examples(self):
        """examples/fft2.rs and examples/rfft2.rs through the fused entry points."""
        be = self.be

        def approx(a, b, t):
            assert np.max(np.abs(self.to_np(a) - np.asarray(b))) <= t

        g = G["example_fft2"]
        v = self.mk(np.array(g["input_real"]) * (1 + 1j))
        vhat = self.zeros((3, 3), np.complex128); v2 = self.zeros((3, 3), np.complex128)
        h0, h1 = be.FftHandler(3), be.FftHandler(3)
        be.fft2(v, vhat, h0, h1)
        want = np.array(g["numpy_vhat"])
        approx(vhat, want[..., 0] + 1j * want[..., 1], g["tol"])
        be.ifft2(vhat, v2, h0, h1)
        approx(v2, np.array(g["input_real"]) * (1 + 1j), g["tol"])
        g = G["example_rfft2"]
        v = self.mk(np.array(g["input_real"]))
        vhat = self.zeros((3, 2), np.complex128); v2 = self.zeros((3, 3), np.float64)
        h0, h1 = be.FftHandler(3), be.R2cFftHandler(3)
        be.rfft2(v, vhat, h0, h1)
        want = np.array(g["numpy_vhat"])
        approx(vhat, want[..., 0] + 1j * want[..., 1], g["tol"])
        be.irfft2(vhat, v2, h0, h1)
        approx(v2, np.array(g["input_real"]), g["tol"])

    def mk_f(self, a):
        """An F-ordered array (host: numpy asfortranarray; device harness overrides)."""
        return np.asfortranarray(np.array(a))
