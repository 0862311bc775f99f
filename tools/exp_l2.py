#!/usr/bin/env python3
"""Two-pass (four-step) transforms with the workspace kept L2-resident: group size sweep (NDFB_FS_L2_KB)."""
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


CASES = [("c2 cols 8192 f32", (8192, 8192), 0, np.float32, 8192),
         ("cols 4096 f64", (4096, 8192), 0, np.float64, 4096),
         ("rows 2^20 f32 x64", (64, 1 << 20), 1, np.float32, 1 << 20),
         ("rows 65536 f64 x512", (512, 65536), 1, np.float64, 65536),
         ("cols 16384 f32 x4096", (16384, 4096), 0, np.float32, 16384)]
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, shape, axis, dt, n in CASES:
    if only and only not in name: continue
    rt = torch.float32 if dt == np.float32 else torch.float64
    x = rnd(shape, rt); y = torch.empty_like(x)
    h = nb.FftHandler(n, dt)
    ref = None
    for kb in (0, 16384, 32768, 65536):
        os.environ["NDFB_FS_L2_KB"] = str(kb)
        ms = timeit(lambda: nb.ndfft(x, y, h, axis))
        # the same call replayed from a CUDA graph: GPU time without the host's launch cost
        gms = None
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                nb.ndfft(x, y, h, axis)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    nb.ndfft(x, y, h, axis)
            gms = timeit(lambda: g.replay())
        except Exception as ex:  # noqa
            gms = str(ex)[:80]
        if ref is None: ref = y.clone()
        same = bool(torch.equal(ref, y))
        nbytes = 2 * x.numel() * x.element_size()
        print(json.dumps({"case": name, "group_KB": kb, "ms": round(ms, 4), "graph_ms": gms if isinstance(gms, str) else round(gms, 4), "frac": round(nbytes / (ms * 1e-3) / 1e9 / PEAK, 3), "identical": same}), flush=True)
    os.environ.pop("NDFB_FS_L2_KB", None)
    del x, y, ref
