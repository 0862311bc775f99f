#!/usr/bin/env python3
"""One ndfft call for profilers: SHAPE=64x16777216 AXIS=1 F64=0 [ITERS=1] python tools/run_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ndrustfft_b200 as nb
shape = tuple(int(v) for v in os.environ["SHAPE"].split("x")); axis = int(os.environ["AXIS"]); f64 = os.environ.get("F64", "0") == "1"
rd = torch.float64 if f64 else torch.float32
x = torch.complex(torch.rand(shape, device="cuda", dtype=rd) * 2 - 1, torch.rand(shape, device="cuda", dtype=rd) * 2 - 1)
y = torch.empty_like(x)
h = nb.FftHandler(shape[axis], np.float64 if f64 else np.float32)
for _ in range(int(os.environ.get("ITERS", "2"))):
    nb.ndfft(x, y, h, axis)
torch.cuda.synchronize()
