#!/usr/bin/env python3
"""Run-time compiled schedules (csrc/jit.h): parity and roofline fraction over lengths WITHOUT an ahead-of-time instance.
One JSON line per (length, dtype, layout): kernel family from NDFB_TRACE, rel-L2 vs numpy (f64), ms, GB/s, frac of the HBM peak."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ndrustfft_b200 as nb

PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
LENGTHS = [48, 72, 96, 120, 144, 160, 192, 200, 224, 240, 288, 320, 336, 400, 432, 448, 480, 500, 576, 640, 720, 768, 800, 840, 896, 960, 1001,
           1080, 1152, 1200, 1280, 1296, 1440, 1536, 1600, 1728, 1920, 2000, 2160, 2187, 2304, 2400, 2560, 2880, 3000, 3072, 3125, 3200, 3456, 3600,
           3840, 4000, 4320, 4800, 5000, 5120, 5184, 5760, 6000, 6144, 6400, 6561, 6912, 7200, 7680, 8000]
only = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else LENGTHS
TARGET_BYTES = 256 << 20


def timeit(fn, iters=7):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for n in only:
    for dt in (np.float32, np.float64):
        es = 8 if dt == np.float32 else 16
        lanes = max(64, TARGET_BYTES // (n * es) // 64 * 64)
        rt = torch.float32 if dt == np.float32 else torch.float64
        for layout, shape, axis in (("rows", (lanes, n), 1), ("cols", (n, lanes), 0)):
            x = torch.complex(torch.rand(shape, device="cuda", dtype=rt) * 2 - 1, torch.rand(shape, device="cuda", dtype=rt) * 2 - 1)
            y = torch.empty_like(x)
            t0 = time.perf_counter()
            h = nb.FftHandler(n, dt)
            nb.ndfft(x, y, h, axis); torch.cuda.synchronize()
            first_s = time.perf_counter() - t0
            # parity on a lane subset against numpy in f64
            sub = (slice(0, 8), slice(None)) if axis == 1 else (slice(None), slice(0, 8))
            want = np.fft.fft(x[sub].cpu().numpy().astype(np.complex128), axis=axis)
            got = y[sub].cpu().numpy()
            rel = float(np.linalg.norm(got - want) / np.linalg.norm(want))
            ms = timeit(lambda: nb.ndfft(x, y, h, axis))
            nbytes = 2 * x.numel() * es
            gbs = nbytes / (ms * 1e-3) / 1e9
            print(json.dumps({"n": n, "dtype": "f32" if dt == np.float32 else "f64", "layout": layout, "lanes": lanes, "ms": round(ms, 4), "GB/s": round(gbs, 1),
                              "frac": round(gbs / PEAK, 4), "GFLOP/s": round(lanes * 5.0 * n * math.log2(n) / (ms * 1e-3) / 1e9, 1), "rel_l2": rel,
                              "first_call_s": round(first_s, 2)}), flush=True)
            del x, y
