#!/usr/bin/env python3
"""c5b (64 x 2^24 c64 rows): two-pass 4096 x 4096 vs three-pass 256 x (256 x 256) with / without the transposing last pass."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ndrustfft_b200 as nb

n = 1 << 24
x = torch.complex(torch.rand((64, n), device="cuda") * 2 - 1, torch.rand((64, n), device="cuda") * 2 - 1)
y = torch.empty_like(x)
h = nb.FftHandler(n, np.float32)
ref = None
for name, env in (("two-pass 4096 x 4096", {}), ("three-pass 256 x 256 x 256", {"NDFB_FS_N1": "256"}),
                  ("three-pass, transposing last pass", {"NDFB_FS_N1": "256", "NDFB_FS_TRANSPOSE": "1"})):
    for k in ("NDFB_FS_N1", "NDFB_FS_TRANSPOSE"): os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(2): nb.ndfft(x, y, h, 1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); nb.ndfft(x, y, h, 1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    if ref is None: ref = y[:, ::4097].clone(); err = 0.0
    else: err = float((torch.linalg.vector_norm(y[:, ::4097] - ref) / torch.linalg.vector_norm(ref)).item())
    print(json.dumps({"variant": name, "ms": round(ts[2], 3), "frac_one_pass": round(2 * x.numel() * 8 / (ts[2] * 1e-3) / 1e9 / 6547.8, 3), "rel_vs_first": err}), flush=True)
