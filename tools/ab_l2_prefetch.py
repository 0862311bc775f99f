#!/usr/bin/env python3
"""A/B: cp.async.bulk.prefetch.L2 of the row a retiring CTA's successor will load (NDFB_L2_PREFETCH=<waves ahead>), long contiguous rows."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, os
sys.path.insert(0, %r)
import numpy as np, torch, ndrustfft_b200 as nb
def timeit(fn, iters=15):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2]
out = {"waves": os.environ.get("NDFB_L2_PREFETCH", "0")}
for dt, n, lanes in ((np.float32, 8192, 8192), (np.float32, 4096, 16384), (np.float64, 4096, 8192), (np.float64, 2048, 16384), (np.float64, 512, 65536)):
    rt = torch.float32 if dt == np.float32 else torch.float64
    x = torch.complex(torch.rand((lanes, n), device="cuda", dtype=rt) * 2 - 1, torch.rand((lanes, n), device="cuda", dtype=rt) * 2 - 1)
    y = torch.empty_like(x)
    h = nb.FftHandler(n, dt)
    out[f"{n}_{'f32' if dt == np.float32 else 'f64'}_ms"] = round(timeit(lambda: nb.ndfft(x, y, h, 1)), 4)
    del x, y
print(json.dumps(out))
''' % ROOT
for w in ("0", "0.5", "1", "2", "0"):
    e = dict(os.environ); e["NDFB_L2_PREFETCH"] = w
    p = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True)
    print((p.stdout.strip().splitlines() or [p.stderr[-300:]])[-1], flush=True)
