#!/usr/bin/env python3
"""Times lengths that need Bluestein (primes, non-smooth): fused bsfft_kernel vs the general tile kernel.

    python tools/bench_bluestein.py            # one JSON line per case on stdout
Device-resident inputs, CUDA events, 20 timed calls after 5 warm-ups; arrays are 256 MiB+ (larger than L2).
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ndrustfft_b200 as nb  # noqa: E402


def time_call(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    peak = 6547.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for n in (127, 257, 1009, 2039):
        for dt, cdt in ((np.float32, torch.complex64), (np.float64, torch.complex128)):
            cs = 8 if dt == np.float32 else 16
            lanes = (256 << 20) // (n * cs)
            lanes -= lanes % 64
            for axis in (1, 0):
                shape = (lanes, n) if axis == 1 else (n, lanes)
                x = torch.randn(shape, dtype=cdt, device="cuda")
                y = torch.empty_like(x)
                h = nb.FftHandler(n, dt)
                row = {"n": n, "dtype": "f32" if dt == np.float32 else "f64", "axis": axis, "lanes": lanes}
                for tag, env in (("fused", None), ("general", "1")):
                    if env:
                        os.environ["NDFB_DISABLE_BSFFT"] = env
                    else:
                        os.environ.pop("NDFB_DISABLE_BSFFT", None)
                    ms = time_call(lambda: nb.ndfft(x, y, h, axis))
                    row[tag + "_ms"] = round(ms, 4)
                    row[tag + "_frac"] = round(2 * x.numel() * cs / (ms * 1e-3) / 1e9 / peak, 3)
                os.environ.pop("NDFB_DISABLE_BSFFT", None)
                print(json.dumps(row), flush=True)
                del x, y


if __name__ == "__main__":
    main()
