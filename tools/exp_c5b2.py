#!/usr/bin/env python3
"""c5b (64 x 2^24 c64 rows): two-pass 4096 x 4096 vs three-pass 256 x (256 x 256) with / without the transposing last pass."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, os
sys.path.insert(0, %r)
import numpy as np, torch, ndrustfft_b200 as nb
n, b = 1 << 24, 64
x = torch.complex(torch.rand((b, n), device="cuda") * 2 - 1, torch.rand((b, n), device="cuda") * 2 - 1)
y = torch.empty_like(x)
h = nb.FftHandler(n, np.float32)
for _ in range(2): nb.ndfft(x, y, h, 1)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); nb.ndfft(x, y, h, 1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
sub = y[:2].cpu().numpy()
want = np.fft.fft(x[:2].cpu().numpy().astype(np.complex128), axis=1)
rel = float(np.linalg.norm(sub - want) / np.linalg.norm(want))
print(json.dumps({"variant": os.environ.get("VARIANT"), "ms": ts[len(ts)//2], "frac_one_pass_bytes": 2 * x.numel() * 8 / (ts[len(ts)//2] * 1e-3) / 1e9 / 6547.8, "rel_l2": rel}))
''' % ROOT
for name, env in (("two-pass 4096x4096", {"NDFB_FS_N1": "4096", "NDFB_FS_TWO_PASS": "1"}), ("two-pass 4096x4096, per-point twiddle lookups", {"NDFB_FS_N1": "4096", "NDFB_FS_TWO_PASS": "1", "NDFB_NO_FS_FACTORED": "1"}),
                  ("default split", {}),
                  ("three-pass 256x256x256 transposing rows", {"NDFB_FS_N1": "256"}),
                  ("three-pass 256x256x256, per-point twiddle lookups", {"NDFB_FS_N1": "256", "NDFB_NO_FS_FACTORED": "1"}),
                  ("three-pass 256x256x256 capped column tile", {"NDFB_FS_N1": "256", "NDFB_NO_TRANS_STORE": "1"}),
                  ("three-pass 512x(...) transposing", {"NDFB_FS_N1": "512"}), ("three-pass 128x(...) transposing", {"NDFB_FS_N1": "128"}),
                  ("three-pass 64x(...) transposing", {"NDFB_FS_N1": "64"})):
    e = dict(os.environ); e.update(env); e["VARIANT"] = name; e["NDFB_TRACE"] = "1"
    p = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    tr = sorted(set(l for l in p.stderr.splitlines() if l.startswith("[ndfb]")))
    print(line[0] if line else json.dumps({"variant": name, "error": p.stderr[-300:]}), flush=True)
    print("   ", " | ".join(t[7:90] for t in tr), flush=True)
