#!/usr/bin/env python3
"""A/B of two library builds on R2C / DCT-II / DCT-I rows (and one column case): pair epilogue from registers (in-thread
mirror pairs for an even number of butterflies per thread, warp shuffles for one) vs the shared-memory epilogue.
   python tools/ab_pair_epilogue.py [libA.so libB.so]     default: lib/libndfft_b200.so (A) vs lib/libndfft_b200_alt.so (B)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ndrustfft_b200 as nb
from ndrustfft_b200 import _lib
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timeit(fn, iters=15):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]


libs = sys.argv[1:3] if len(sys.argv) >= 3 else [os.path.join(ROOT, "ndrustfft_b200", "lib", "libndfft_b200.so"), os.path.join(ROOT, "ndrustfft_b200", "lib", "libndfft_b200_alt.so")]
bes = [nb.Backend(_lib.CLib(p)) for p in libs]
for dt, n, axis in ((np.float64, 512, 1), (np.float64, 1024, 1), (np.float64, 2048, 1), (np.float64, 4096, 1), (np.float64, 8192, 1), (np.float32, 1024, 1),
                    (np.float32, 2048, 1), (np.float32, 4096, 1), (np.float32, 8192, 1), (np.float64, 1024, 0), (np.float64, 4096, 0)):
    rt = torch.float32 if dt == np.float32 else torch.float64
    es = 4 if dt == np.float32 else 8
    lanes = (256 << 20) // (n * es)
    shape = (lanes, n) if axis == 1 else (n, lanes)
    x = torch.rand(shape, device="cuda", dtype=rt) * 2 - 1
    so = list(shape); so[axis] = n // 2 + 1
    yc = torch.empty(so, device="cuda", dtype=torch.complex64 if dt == np.float32 else torch.complex128)
    yr = torch.empty_like(x)
    row = {"n": n, "dtype": "f32" if dt == np.float32 else "f64", "layout": "rows" if axis == 1 else "cols"}
    refs = {}
    for tag, be in zip("AB", bes):
        hr, hd = be.R2cFftHandler(n, dt), be.DctHandler(n, dt)
        row[f"r2c_{tag}_ms"] = round(timeit(lambda: be.ndfft_r2c(x, yc, hr, axis)), 4)
        rc = yc.clone()
        row[f"dct2_{tag}_ms"] = round(timeit(lambda: be.nddct2(x, yr, hd, axis)), 4)
        rd = yr.clone()
        if tag == "A": refs = {"c": rc, "d": rd}
        else:
            row["r2c_rel_diff"] = float((torch.linalg.vector_norm(rc - refs["c"]) / torch.linalg.vector_norm(refs["c"])).item())
            row["dct2_rel_diff"] = float((torch.linalg.vector_norm(rd - refs["d"]) / torch.linalg.vector_norm(refs["d"])).item())
    row["r2c_frac_A"] = round((x.numel() * es + yc.numel() * 2 * es) / (row["r2c_A_ms"] * 1e-3) / 1e9 / PEAK, 3)
    row["dct2_frac_A"] = round(2 * x.numel() * es / (row["dct2_A_ms"] * 1e-3) / 1e9 / PEAK, 3)
    row["r2c_B/A"] = round(row["r2c_B_ms"] / row["r2c_A_ms"], 3); row["dct2_B/A"] = round(row["dct2_B_ms"] / row["dct2_A_ms"], 3)
    print(json.dumps(row), flush=True)
    del x, yc, yr, refs
