#!/usr/bin/env python3
"""c3 on one GPU: layout of the intermediate between the axis-1 and axis-0 passes.
natural  b[i0][i1][i2]: the axis-0 pass reads rows 2.16 MB apart (one 2 MiB page per row of a tile);
permuted b[i1][i0][i2]: the axis-0 pass reads rows 4 KB apart, the axis-1 pass WRITES rows 2.16 MB apart instead."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ndrustfft_b200 as nb
n, m, mp = 512, 257, 264
dev = "cuda"
x = torch.rand((n, n, n), device=dev, dtype=torch.float64) * 2 - 1
a = torch.zeros((n, n, mp), dtype=torch.complex128, device=dev)
b_nat = torch.zeros((n, n, mp), dtype=torch.complex128, device=dev)
b_perm = torch.zeros((n, n, mp), dtype=torch.complex128, device=dev).permute(1, 0, 2)     # logical [i0][i1][i2], memory [i1][i0][i2]
out = torch.zeros((n, n, mp), dtype=torch.complex128, device=dev)
h2, h1, h0 = nb.R2cFftHandler(n), nb.FftHandler(n), nb.FftHandler(n)


def T(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]


nb.ndfft_r2c(x, a[:, :, :m], h2, 2)
res = {}
res["axis1_natural_ms"] = T(lambda: nb.ndfft(a, b_nat, h1, 1))
res["axis0_natural_ms"] = T(lambda: nb.ndfft(b_nat, out, h0, 0))
ref = out.clone()
res["axis1_permuted_out_ms"] = T(lambda: nb.ndfft(a, b_perm, h1, 1))
res["axis0_permuted_in_ms"] = T(lambda: nb.ndfft(b_perm, out, h0, 0))
res["same_result"] = bool(torch.equal(ref, out))
out_perm = torch.zeros((n, n, mp), dtype=torch.complex128, device=dev).permute(1, 0, 2)
res["axis0_permuted_in_and_out_ms"] = T(lambda: nb.ndfft(b_perm, out_perm, h0, 0))
print(json.dumps(res))
