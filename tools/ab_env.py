#!/usr/bin/env python3
"""Same-box A/B of library code paths selected by environment variables (one subprocess per variant):
   python tools/ab_env.py [substring filters...]   ->  one JSON line per (case, variant)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, os
sys.path.insert(0, %r)
import numpy as np, torch, ndrustfft_b200 as nb
shape = tuple(int(v) for v in os.environ["SHAPE"].split("x")); axis = int(os.environ["AXIS"]); f64 = os.environ["F64"] == "1"; op = os.environ["OP"]
rt = torch.float64 if f64 else torch.float32
rd = np.float64 if f64 else np.float32
n = shape[axis]
cx = op in ("ndfft", "ndifft")
x = torch.complex(torch.rand(shape, device="cuda", dtype=rt) * 2 - 1, torch.rand(shape, device="cuda", dtype=rt) * 2 - 1) if cx else torch.rand(shape, device="cuda", dtype=rt) * 2 - 1
y = torch.empty_like(x)
h = (nb.FftHandler if cx else nb.DctHandler)(n, rd)
f = getattr(nb, op)
for _ in range(3): f(x, y, h, axis)
torch.cuda.synchronize()
reps = int(os.environ.get("REPS", "4"))
ts = []
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f(x, y, h, axis)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps)
ts.sort()
# (timing tool only: parity against the oracle is the job of tests/; here a direct O(n^2) f64 sum of ONE output element per kind
# guards against measuring a broken path)
def direct(op, v, k):
    m = np.arange(n)
    if op == "nddct1":
        return 2 * (0.5 * (v[0] + (-1) ** k * v[-1]) + np.sum(v[1:-1] * np.cos(np.pi * m[1:-1] * k / (n - 1))))
    if op == "nddct2": return 2 * np.sum(v * np.cos(np.pi * (m + 0.5) * k / n))
    if op == "nddct3": return 2 * (0.5 * v[0] + np.sum(v[1:] * np.cos(np.pi * m[1:] * (k + 0.5) / n)))
    if op == "nddct4": return 2 * np.sum(v * np.cos(np.pi * (m + 0.5) * (k + 0.5) / n))
    return np.sum(v * np.exp(-2j * np.pi * m * k / n))
idx = [0] * len(shape); idx[axis] = slice(None)
lane_in = x[tuple(idx)].cpu().numpy().astype(np.complex128 if cx else np.float64)
lane_out = y[tuple(idx)].cpu().numpy()
k = n // 3
want = direct(op, lane_in, k)
rel = float(abs(lane_out[k] - want) / (abs(want) + 1e-300))
ms = ts[len(ts) // 2]
nbytes = 2 * x.numel() * x.element_size()
print(json.dumps({"case": os.environ["CASE"], "variant": os.environ["VARIANT"], "ms": round(ms, 4), "ms_min": round(ts[0], 4),
                  "frac": round(nbytes / (ms * 1e-3) / 1e9 / 6547.8, 4), "rel_err_one_element": rel}))
''' % ROOT
V2 = [("staged copy-out", {"NDFB_NO_MIRROR_OUT": "1"}), ("mirror-paired output pass", {})]
D1R = [("13.9.7.5 on 512 threads, 2 CTAs/SM", {"NDFB_RSFFT_PICK": "4095:1"}), ("15.13.7.3 on 320 threads, 2 CTAs/SM", {"NDFB_RSFFT_PICK": "4095:2"}),
       ("15.13.7.3 on 320 threads, 3 CTAs/SM", {"NDFB_RSFFT_PICK": "4095:3"})]
D1C = [("13.9.7.5, two-column tile of 1024 threads", {"NDFB_RSFFT_PICK": "4095:0"}), ("15.13.7.3, two-column tile of 640 threads", {"NDFB_RSFFT_PICK": "4095:1"})]
CASES = [  # case, op, shape, axis, f64, [(variant, env)]
    ("c4 nddct1 rows 4096^2 f64", "nddct1", "4096x4096", 1, 1, D1R),
    ("c4 nddct1 columns 4096^2 f64", "nddct1", "4096x4096", 0, 1, D1C),
    ("nddct1 rows 8192x4096 f32", "nddct1", "8192x4096", 1, 0, D1R),
    ("c4 nddct3 rows 4096^2 f64", "nddct3", "4096x4096", 1, 1, V2),
    ("c4 nddct4 rows 4096^2 f64", "nddct4", "4096x4096", 1, 1, V2),
]
only = sys.argv[1:]
for case, op, shape, axis, f64, variants in CASES:
    if only and not any(o in case for o in only):
        continue
    for vname, env in variants:
        e = dict(os.environ); e.update(env); e.update({"CASE": case, "VARIANT": vname, "SHAPE": shape, "AXIS": str(axis), "F64": str(f64), "OP": op})
        p = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        print(line[0] if line else json.dumps({"case": case, "variant": vname, "error": p.stderr[-400:]}), flush=True)
