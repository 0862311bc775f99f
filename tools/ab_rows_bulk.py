#!/usr/bin/env python3
"""A/B on one box: register-resident row kernel (two CTAs per SM cover each other's load/store phases) vs the persistent
bulk-async row kernel (TMA loads and stores + mbarrier, NDFB_ROWS_BULK=1).  One JSON line per case."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ndrustfft_b200 as nb
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timeit(fn, iters=15):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]


for dt, n, lanes in ((np.float32, 8192, 8192), (np.float32, 4096, 16384), (np.float32, 2048, 32768), (np.float32, 1024, 65536),
                     (np.float64, 4096, 8192), (np.float64, 2048, 16384), (np.float64, 1024, 32768), (np.float64, 512, 65536)):
    rt = torch.float32 if dt == np.float32 else torch.float64
    es = 8 if dt == np.float32 else 16
    x = torch.complex(torch.rand((lanes, n), device="cuda", dtype=rt) * 2 - 1, torch.rand((lanes, n), device="cuda", dtype=rt) * 2 - 1)
    y0 = torch.empty_like(x); y1 = torch.empty_like(x)
    h = nb.FftHandler(n, dt)
    res = {"n": n, "dtype": "f32" if dt == np.float32 else "f64", "lanes": lanes}
    for key, env, y in (("register_resident", "0", y0), ("bulk_async", "1", y1)):
        os.environ["NDFB_ROWS_BULK"] = env
        ms = timeit(lambda: nb.ndfft(x, y, h, 1))
        res[key + "_ms"] = round(ms, 4)
        res[key + "_frac"] = round(2 * x.numel() * es / (ms * 1e-3) / 1e9 / PEAK, 4)
    res["bit_identical"] = bool(torch.equal(y0, y1))
    want = np.fft.fft(x[:4].cpu().numpy().astype(np.complex128), axis=1)
    res["rel_l2_bulk"] = float(np.linalg.norm(y1[:4].cpu().numpy() - want) / np.linalg.norm(want))
    print(json.dumps(res), flush=True)
    del x, y0, y1
