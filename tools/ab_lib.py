#!/usr/bin/env python3
"""A/B timing of two builds of the library on the same box:  python tools/ab_lib.py [libA.so libB.so]
(default: ndrustfft_b200/lib/libndfft_b200.so vs libndfft_b200_alt.so).  One JSON line per case and build."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ndrustfft_b200 as nb
from ndrustfft_b200 import _lib

PEAK = 6547.8


def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def rnd(shape, rt, cx):
    if cx:
        return torch.complex(torch.rand(shape, device="cuda", dtype=rt) * 2 - 1, torch.rand(shape, device="cuda", dtype=rt) * 2 - 1)
    return torch.rand(shape, device="cuda", dtype=rt) * 2 - 1


CASES = [  # name, fn, in shape, out shape, axis, dtype, n, handler
    ("c2 rows 8192 f32", "ndfft", (8192, 8192), None, 1, np.float32, 8192, "FftHandler"),
    ("c2 cols 8192 f32", "ndfft", (8192, 8192), None, 0, np.float32, 8192, "FftHandler"),
    ("c3 r2c 512 f64 axis2", "ndfft_r2c", (512, 512, 512), (512, 512, 257), 2, np.float64, 512, "R2cFftHandler"),
    ("c3 fft 512 f64 axis1", "ndfft", (512, 512, 257), None, 1, np.float64, 512, "FftHandler"),
    ("c3 fft 512 f64 axis0", "ndfft", (512, 512, 257), None, 0, np.float64, 512, "FftHandler"),
    ("c4 dct2 rows", "nddct2", (4096, 4096), None, 1, np.float64, 4096, "DctHandler"),
    ("c4 dct2 cols", "nddct2", (4096, 4096), None, 0, np.float64, 4096, "DctHandler"),
    ("c4 dct1 rows", "nddct1", (4096, 4096), None, 1, np.float64, 4096, "DctHandler"),
    ("c4 dct4 rows", "nddct4", (4096, 4096), None, 1, np.float64, 4096, "DctHandler"),
    ("c4 dct1 cols", "nddct1", (4096, 4096), None, 0, np.float64, 4096, "DctHandler"),
    ("ramp 264 c128 axis0", "ndfft", (264, 264 * 64), None, 0, np.float64, 264, "FftHandler"),
    ("rows 729 f32", "ndfft", (65536, 729), None, 1, np.float32, 729, "FftHandler"),
    ("rows 600 f64", "ndfft", (65536, 600), None, 1, np.float64, 600, "FftHandler"),
    ("c5a 360 axis0", "ndfft", (360, 1000, 384), None, 0, np.float64, 360, "FftHandler"),
    ("c5a 1000 axis1", "ndfft", (360, 1000, 384), None, 1, np.float64, 1000, "FftHandler"),
    ("c5a 384 axis2", "ndfft", (360, 1000, 384), None, 2, np.float64, 384, "FftHandler"),
    ("rows 1024 f32", "ndfft", (65536, 1024), None, 1, np.float32, 1024, "FftHandler"),
    ("rows 4096 f64", "ndfft", (8192, 4096), None, 1, np.float64, 4096, "FftHandler"),
    ("rows 2048 f32", "ndfft", (32768, 2048), None, 1, np.float32, 2048, "FftHandler"),
    ("rows 1024 f64", "ndfft", (32768, 1024), None, 1, np.float64, 1024, "FftHandler"),
    ("cols 2048 f64", "ndfft", (2048, 16384), None, 0, np.float64, 2048, "FftHandler"),
    ("cols 512 f32", "ndfft", (512, 131072), None, 0, np.float32, 512, "FftHandler"),
    ("blu 1009 f32 rows", "ndfft", (33216, 1009), None, 1, np.float32, 1009, "FftHandler"),
]


def main():
    libs = sys.argv[1:3] if len(sys.argv) >= 3 else [os.path.join(ROOT, "ndrustfft_b200", "lib", "libndfft_b200.so"),
                                                     os.path.join(ROOT, "ndrustfft_b200", "lib", "libndfft_b200_alt.so")]
    bes = [nb.Backend(_lib.CLib(p)) for p in libs]
    for name, fn, si, so, axis, dt, n, hk in CASES:
        rt = torch.float32 if dt == np.float32 else torch.float64
        ct = torch.complex64 if dt == np.float32 else torch.complex128
        cx_in = fn in ("ndfft", "ndifft", "ndifft_r2c")
        cx_out = fn in ("ndfft", "ndifft", "ndfft_r2c")
        x = rnd(si, rt, cx_in)
        y = torch.empty(so or si, dtype=ct if cx_out else rt, device="cuda")
        nbytes = x.numel() * x.element_size() + y.numel() * y.element_size()
        row = {"case": name}
        ref = None
        for tag, be in zip("AB", bes):
            h = getattr(be, hk)(n, dt)
            f = getattr(be, fn)
            ms = timeit(lambda: f(x, y, h, axis))
            row[tag + "_ms"] = round(ms, 4)
            row[tag + "_frac"] = round(nbytes / (ms * 1e-3) / 1e9 / PEAK, 3)
            if ref is None: ref = y.clone()
            else:
                row["identical"] = bool(torch.equal(ref, y))
                row["rel_l2_B_vs_A"] = float(torch.linalg.vector_norm((y - ref).flatten()) / torch.linalg.vector_norm(ref.flatten()))
        row["B/A"] = round(row["B_ms"] / row["A_ms"], 3)
        print(json.dumps(row), flush=True)
        del x, y, ref


if __name__ == "__main__":
    main()
