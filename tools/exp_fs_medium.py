#!/usr/bin/env python3
"""Multi-pass C2C rows of 2^16 .. 2^23 points (f32, 4.3 GB per array): default split, transposing second pass, three passes."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os, json
sys.path.insert(0, %r)
import numpy as np, torch, ndrustfft_b200 as nb
f64 = os.environ.get("F64") == "1"
rt = torch.float64 if f64 else torch.float32
for lg in (16, 18, 19, 20, 21, 22, 23):
    n = 1 << lg; b = (1 << (28 if f64 else 29)) // n
    x = torch.complex(torch.rand((b, n), device="cuda", dtype=rt), torch.rand((b, n), device="cuda", dtype=rt)); y = torch.empty_like(x)
    h = nb.FftHandler(n, np.float64 if f64 else np.float32)
    try:
        for _ in range(3): nb.ndfft(x, y, h, 1)
    except Exception as ex:
        print(json.dumps({"n": n, "variant": os.environ.get("V"), "error": str(ex)[:120]})); continue
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); nb.ndfft(x, y, h, 1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    want = np.fft.fft(x[:1].cpu().numpy().astype(np.complex128), axis=1)
    print(json.dumps({"n": n, "batch": b, "dtype": "c128" if f64 else "c64", "variant": os.environ.get("V"), "ms": round(ts[3], 4), "frac_one_pass": round(2 * x.numel() * x.element_size() / (ts[3] * 1e-3) / 1e9 / 6547.8, 3),
                      "rel_l2": float(np.linalg.norm(y[:1].cpu().numpy() - want) / np.linalg.norm(want))}), flush=True)
    del x, y
''' % ROOT
for f64 in ("0", "1"):
    for v, env in (("default", {}), ("two passes wherever both factors fit on chip (NDFB_FS_TWO_PASS=1)", {"NDFB_FS_TWO_PASS": "1"})):
        e = dict(os.environ); e.update(env); e["V"] = v; e["F64"] = f64; e["NDFB_TRACE"] = "1"
        p = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True)
        print(p.stdout, end="", flush=True)
        tr = sorted(set(l for l in p.stderr.splitlines() if l.startswith("[ndfb] four-step")))
        print("    " + " | ".join(t[17:70] for t in tr), flush=True)
