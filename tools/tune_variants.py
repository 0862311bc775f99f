#!/usr/bin/env python3
"""A/B timing of the alternative Stockham schedules kept in the registry (NDFB_SFFT_PICK=<N>:<index>)."""
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


def case(name, shape, axis, dt, n, nvariants):
    rt = torch.float32 if dt == np.float32 else torch.float64
    x = rnd(shape, rt); y = torch.empty_like(x)
    h = nb.FftHandler(n, dt)
    es = 8 if dt == np.float32 else 16
    ref = None
    for idx in range(nvariants):
        os.environ["NDFB_SFFT_PICK"] = f"{n}:{idx}"
        os.environ["NDFB_STRIDED_FOURSTEP"] = "0"
        ms = timeit(lambda: nb.ndfft(x, y, h, axis))
        if ref is None:
            ref = y.clone()
            err = 0.0
        else:
            err = (torch.linalg.vector_norm(y - ref) / torch.linalg.vector_norm(ref)).item()
        gbs = 2 * x.numel() * es / (ms * 1e-3) / 1e9
        print(json.dumps({"case": name, "variant": idx, "ms": round(ms, 4), "GB/s": round(gbs, 1), "frac": round(gbs / PEAK, 4), "rel_vs_v0": err}), flush=True)
    del os.environ["NDFB_SFFT_PICK"]; del os.environ["NDFB_STRIDED_FOURSTEP"]


case("c2 rows 8192 f32", (8192, 8192), 1, np.float32, 8192, 4)
case("c3 cols 512 f64 (axis1)", (512, 512, 257), 1, np.float64, 512, 5)
case("c3 cols 512 f64 (axis0)", (512, 512, 257), 0, np.float64, 512, 5)
case("rows 512 f64", (512 * 257, 512), 1, np.float64, 512, 3)
