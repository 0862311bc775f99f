#!/usr/bin/env python3
"""One more CTA per SM through a tighter register cap (experimental registry variants, NDFB_GEN_EXPERIMENT=1 build)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ndrustfft_b200 as nb

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


# name, fn, in shape, out shape, axis, dtype, n, handler, env var, core N, (index of the current default, index of the variant)
CASES = [
    ("c4 dct2 cols f64 idx2 vs idx3 (L=4 T=1024 vs L=2 T=512)", "nddct2", (4096, 4096), None, 0, np.float64, 4096, "DctHandler", "NDFB_RSFFT_PICK", 2048, (2, 3)),
    ("c4 dct2 cols f64 idx0 vs idx1 (family A)", "nddct2", (4096, 4096), None, 0, np.float64, 4096, "DctHandler", "NDFB_RSFFT_PICK", 2048, (0, 1)),
    ("c4 dct3 cols f64 idx2 vs idx3", "nddct3", (4096, 4096), None, 0, np.float64, 4096, "DctHandler", "NDFB_RSFFT_PICK", 2048, (2, 3)),
    ("c4 dct4 cols f64 idx2 vs idx3", "nddct4", (4096, 4096), None, 0, np.float64, 4096, "DctHandler", "NDFB_RSFFT_PICK", 2048, (2, 3)),
    ("c4 dct1 cols f64 (new tile; both default)", "nddct1", (4096, 4096), None, 0, np.float64, 4096, "DctHandler", "NDFB_RSFFT_PICK", 1, (0, 0)),
    ("c2-like cols 4096 f64 idx default vs L=2", "ndfft", (4096, 8192), None, 0, np.float64, 4096, "FftHandler", "NDFB_SFFT_PICK", 1, (0, 0)),
]
os.environ["NDFB_TRACE"] = "1"
for name, fn, si, so, axis, dt, n, hk, var, core, idxs in CASES:
    rt = torch.float32 if dt == np.float32 else torch.float64
    ct = torch.complex64 if dt == np.float32 else torch.complex128
    cx_in = fn in ("ndfft", "ndifft", "ndifft_r2c"); cx_out = fn in ("ndfft", "ndifft", "ndfft_r2c")
    x = rnd(si, rt, cx_in); y = torch.empty(so or si, dtype=ct if cx_out else rt, device="cuda")
    h = getattr(nb, hk)(n, dt)
    f = getattr(nb, fn)
    nbytes = x.numel() * x.element_size() + y.numel() * y.element_size()
    row = {"case": name}
    ref = None
    for tag, idx in zip(("base", "more_ctas"), idxs):
        os.environ[var] = f"{core}:{idx}"
        ms = timeit(lambda: f(x, y, h, axis))
        row[tag + "_ms"] = round(ms, 4); row[tag + "_frac"] = round(nbytes / (ms * 1e-3) / 1e9 / PEAK, 3)
        if ref is None: ref = y.clone()
        else: row["identical"] = bool(torch.equal(ref, y))
    os.environ.pop(var, None)
    row["ratio"] = round(row["more_ctas_ms"] / row["base_ms"], 3)
    print(json.dumps(row), flush=True)
    del x, y, ref
