#!/usr/bin/env python3
"""Fused two-pass kernel (fs2_kernel) vs the two-launch four-step on long strided columns: group budget / ring sweep."""
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


def rnd(shape, rt):
    return torch.complex(torch.rand(shape, device="cuda", dtype=rt) * 2 - 1, torch.rand(shape, device="cuda", dtype=rt) * 2 - 1)


CASES = [("c2 cols 8192 f32", (8192, 8192), np.float32, 8192),
         ("cols 16384 f32", (16384, 4096), np.float32, 16384)]
os.environ["NDFB_FS2_ALL"] = "1"
for name, shape, dt, n in CASES:
    rt = torch.float32 if dt == np.float32 else torch.float64
    x = rnd(shape, rt); y = torch.empty_like(x)
    h = nb.FftHandler(n, dt)
    os.environ["NDFB_NO_FS2"] = "1"
    ms = timeit(lambda: nb.ndfft(x, y, h, 0))
    ref = y.clone()
    nbytes = 2 * x.numel() * x.element_size()
    print(json.dumps({"case": name, "variant": "two launches", "ms": round(ms, 4), "frac": round(nbytes / (ms * 1e-3) / 1e9 / PEAK, 3)}), flush=True)
    del os.environ["NDFB_NO_FS2"]
    os.environ["NDFB_FS2_F64"] = "1"
    for kb, ring, dbg, T in ((8192, 3, 0, 256), (8192, 3, 0, 512), (16384, 3, 0, 512), (16384, 4, 0, 512), (32768, 3, 0, 512), (8192, 3, 3, 512), (8192, 3, 3, 256)):
        os.environ["NDFB_FS2_KB"] = str(kb); os.environ["NDFB_FS2_RING"] = str(ring); os.environ["NDFB_FS2_DBG"] = str(dbg); os.environ["NDFB_FS2_T"] = str(T)
        y.zero_()
        ms = timeit(lambda: nb.ndfft(x, y, h, 0))
        print(json.dumps({"case": name, "variant": f"fused {kb} KB ring {ring} dbg {dbg} T {T}", "ms": round(ms, 4), "frac": round(nbytes / (ms * 1e-3) / 1e9 / PEAK, 3),
                          "identical": bool(torch.equal(ref, y))}), flush=True)
    os.environ.pop("NDFB_FS2_KB", None); os.environ.pop("NDFB_FS2_RING", None); os.environ.pop("NDFB_FS2_DBG", None); os.environ.pop("NDFB_FS2_T", None)
    del x, y, ref
