#!/usr/bin/env python3
"""A/B of the software-pipelined persistent column kernel (pipe_kernel.cuh) against the register-resident kernels, same box,
one subprocess per variant: c5b (64 x 2^24 c64, both passes), c2 axis 0 (8192-point c64 columns: two passes vs ONE pipelined pass
over 16-byte rows), long f64 columns.  Prints one JSON line per variant."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, os
sys.path.insert(0, %r)
import numpy as np, torch, ndrustfft_b200 as nb
shape = tuple(int(v) for v in os.environ["SHAPE"].split("x")); axis = int(os.environ["AXIS"]); f64 = os.environ["F64"] == "1"
rd = torch.float64 if f64 else torch.float32
x = torch.complex(torch.rand(shape, device="cuda", dtype=rd) * 2 - 1, torch.rand(shape, device="cuda", dtype=rd) * 2 - 1)
y = torch.empty_like(x)
n = shape[axis]
h = nb.FftHandler(n, np.float64 if f64 else np.float32)
for _ in range(3): nb.ndfft(x, y, h, axis)
torch.cuda.synchronize()
ts = []
for _ in range(int(os.environ.get("ITERS", "10"))):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); nb.ndfft(x, y, h, axis); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
# a few lanes against numpy in f64
idx = [slice(None)] * len(shape)
for d in range(len(shape)):
    if d != axis: idx[d] = slice(0, 2)
sub = y[tuple(idx)].cpu().numpy()
want = np.fft.fft(x[tuple(idx)].cpu().numpy().astype(np.complex128), axis=axis)
rel = float(np.linalg.norm(sub - want) / np.linalg.norm(want))
ms = ts[len(ts) // 2]
nbytes = 2 * x.numel() * x.element_size()
print(json.dumps({"variant": os.environ.get("VARIANT"), "shape": os.environ["SHAPE"], "axis": axis, "dtype": "c128" if f64 else "c64", "ms": round(ms, 4), "ms_min": round(ts[0], 4),
                  "frac_one_pass_bytes": round(nbytes / (ms * 1e-3) / 1e9 / 6547.8, 4), "rel_l2": rel}))
''' % ROOT
CASES = [
    ("c5b two-pass, register-resident", "64x16777216", 1, 0, {"NDFB_PIPE": "0"}),
    ("c5b two-pass, pipelined", "64x16777216", 1, 0, {"NDFB_PIPE": "1"}),
    ("c5b two-pass, pipelined, two-lane tiles (2 CTAs/SM, 16-byte rows)", "64x16777216", 1, 0, {"NDFB_PIPE": "2", "NDFB_PIPE_L": "2"}),
    ("c2 axis0 two passes (default)", "8192x8192", 0, 0, {"NDFB_PIPE": "0"}),
    ("c2 axis0 ONE pass, 16-byte rows, register-resident", "8192x8192", 0, 0, {"NDFB_PIPE": "0", "NDFB_STRIDED_FOURSTEP": "0"}),
    ("c2 axis0 ONE pass, 16-byte rows, pipelined", "8192x8192", 0, 0, {"NDFB_PIPE": "1", "NDFB_STRIDED_FOURSTEP": "0"}),
    ("4096-pt c64 columns, register-resident", "4096x16384", 0, 0, {"NDFB_PIPE": "0"}),
    ("4096-pt c64 columns, pipelined", "4096x16384", 0, 0, {"NDFB_PIPE": "1"}),
    ("2048-pt c128 columns, register-resident", "2048x16384", 0, 1, {"NDFB_PIPE": "0"}),
    ("2048-pt c128 columns, pipelined", "2048x16384", 0, 1, {"NDFB_PIPE": "1"}),
    ("4096-pt c128 columns, register-resident", "4096x8192", 0, 1, {"NDFB_PIPE": "0"}),
    ("4096-pt c128 columns, pipelined", "4096x8192", 0, 1, {"NDFB_PIPE": "1"}),
    ("2048-pt c64 columns, register-resident", "2048x32768", 0, 0, {"NDFB_PIPE": "0"}),
    ("2048-pt c64 columns, pipelined (forced)", "2048x32768", 0, 0, {"NDFB_PIPE": "2"}),
    ("c2 rows 8192-pt c64, register-resident", "8192x8192", 1, 0, {"NDFB_PIPE": "0"}),
    ("c2 rows 8192-pt c64, pipelined one-lane tiles", "8192x8192", 1, 0, {"NDFB_PIPE": "2"}),
    ("4096-pt c128 rows, register-resident", "8192x4096", 1, 1, {"NDFB_PIPE": "0"}),
    ("4096-pt c128 rows, pipelined one-lane tiles", "8192x4096", 1, 1, {"NDFB_PIPE": "2"}),
    ("4096-pt c64 rows, register-resident", "16384x4096", 1, 0, {"NDFB_PIPE": "0"}),
    ("4096-pt c64 rows, pipelined one-lane tiles", "16384x4096", 1, 0, {"NDFB_PIPE": "2"}),
    ("c5a axis1 1000-pt c128 columns, register-resident", "360x1000x384", 1, 1, {"NDFB_PIPE": "0"}),
    ("c5a axis1 1000-pt c128 columns, pipelined L=4 (2 CTAs/SM)", "360x1000x384", 1, 1, {"NDFB_PIPE": "2", "NDFB_PIPE_L": "4"}),
    ("c5a axis1 1000-pt c128 columns, pipelined L=8", "360x1000x384", 1, 1, {"NDFB_PIPE": "2", "NDFB_PIPE_L": "8"}),
]
only = sys.argv[1:] 
for name, shape, axis, f64, env in CASES:
    if only and not any(o in name for o in only):
        continue
    e = dict(os.environ); e.update(env); e.update({"VARIANT": name, "SHAPE": shape, "AXIS": str(axis), "F64": str(f64), "NDFB_TRACE": "1"})
    p = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    tr = sorted(set(l for l in p.stderr.splitlines() if l.startswith("[ndfb]")))
    print(line[0] if line else json.dumps({"variant": name, "error": p.stderr[-400:]}), flush=True)
    print("   ", " | ".join(t[7:120] for t in tr), flush=True)
