#!/usr/bin/env python3
"""A/B timing of the alternative schedules / tile widths kept in the registry.
NDFB_SFFT_PICK=<N>:<index> (C2C) and NDFB_RSFFT_PICK=<N>:<index> (real kinds) select the index-th registry entry of that
core length; NDFB_TRACE prints what actually ran."""
import json, os, sys, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ndrustfft_b200 as nb

PEAK = 6547.8


def timeit(fn, iters=8):
    for _ in range(2): fn()
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


def case(name, fn_name, shape, axis, dt, n, core, nvariants, real=False):
    rt = torch.float32 if dt == np.float32 else torch.float64
    cx_in = fn_name in ("ndfft", "ndifft")
    x = rnd(shape, rt, cx_in); y = torch.empty_like(x)
    H = nb.FftHandler if cx_in else nb.DctHandler
    h = H(n, dt)
    f = getattr(nb, fn_name)
    es = (8 if dt == np.float32 else 16) if cx_in else (4 if dt == np.float32 else 8)
    ref = None
    var = "NDFB_RSFFT_PICK" if real else "NDFB_SFFT_PICK"
    os.environ["NDFB_STRIDED_FOURSTEP"] = "0"
    for idx in range(-1, nvariants):
        if idx >= 0: os.environ[var] = f"{core}:{idx}"
        ms = timeit(lambda: f(x, y, h, axis))
        if ref is None: ref = y.clone(); err = 0.0
        else: err = (torch.linalg.vector_norm(y - ref) / torch.linalg.vector_norm(ref)).item()
        gbs = 2 * x.numel() * es / (ms * 1e-3) / 1e9
        print(json.dumps({"case": name, "variant": "default" if idx < 0 else idx, "ms": round(ms, 4), "frac": round(gbs / PEAK, 4), "rel_vs_default": err}), flush=True)
    os.environ.pop(var, None); os.environ.pop("NDFB_STRIDED_FOURSTEP", None)


if len(sys.argv) > 1 and sys.argv[1] == "mixed":
    case("c5a cols 360 f64 axis0", "ndfft", (360, 1000, 384), 0, np.float64, 360, 360, 5)
    case("c5a cols 1000 f64 axis1", "ndfft", (360, 1000, 384), 1, np.float64, 1000, 1000, 3)
    case("cols 600 f64", "ndfft", (600, 65536), 0, np.float64, 600, 600, 3)
    case("cols 384 f64", "ndfft", (384, 131072), 0, np.float64, 384, 384, 5)
    sys.exit(0)
case("c3 cols 512 f64 axis1", "ndfft", (512, 512, 257), 1, np.float64, 512, 512, 6)
case("c3 cols 512 f64 axis0", "ndfft", (512, 512, 257), 0, np.float64, 512, 512, 6)
case("c5a cols 360 f64 axis0", "ndfft", (360, 1000, 384), 0, np.float64, 360, 360, 3)
case("c5a cols 1000 f64 axis1", "ndfft", (360, 1000, 384), 1, np.float64, 1000, 1000, 2)
case("c4 dct2 cols 4096 f64 axis0", "nddct2", (4096, 4096), 0, np.float64, 4096, 2048, 4, real=True)
case("c4 dct2 rows 4096 f64 axis1", "nddct2", (4096, 4096), 1, np.float64, 4096, 2048, 2, real=True)
case("c2 pass N=64 f32 cols", "ndfft", (64, 1 << 20), 0, np.float32, 64, 64, 5)
case("c2 pass N=128 f32 cols", "ndfft", (128, 1 << 19), 0, np.float32, 128, 128, 5)
