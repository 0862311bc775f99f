#!/usr/bin/env python3
"""Per-transform timings of every BASELINE.json config on one GPU (device-resident, CUDA events).

Prints one JSON line per (config, call): ms, GB/s of algorithmic bytes (input once + output once),
fraction of the measured HBM peak, GFLOP/s by the 5 N log2 N (complex) / 2.5 N log2 N (real) convention.
Usage: python tools/bench_configs.py [--only c2,c3] [--iters 20]
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import ndrustfft_b200 as nb


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


PEAK = peak()
FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    FLUSH.zero_()


def time_call(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def rnd(shape, dt, cx):
    rt = torch.float32 if dt == np.float32 else torch.float64
    if cx:
        return torch.complex(torch.rand(shape, device="cuda", dtype=rt) * 2 - 1, torch.rand(shape, device="cuda", dtype=rt) * 2 - 1)
    return torch.rand(shape, device="cuda", dtype=rt) * 2 - 1


def report(cfg, name, ms_med, ms_min, nbytes, flops):
    gbs = nbytes / (ms_med * 1e-3) / 1e9
    print(json.dumps({"cfg": cfg, "call": name, "ms": round(ms_med, 4), "ms_min": round(ms_min, 4), "GB/s": round(gbs, 1),
                      "frac_hbm": round(gbs / PEAK, 4), "GFLOP/s": round(flops / (ms_med * 1e-3) / 1e9, 1)}), flush=True)


def run_c2c(cfg, shape, axes, dt, iters, flush=False):
    x = rnd(shape, dt, True); y = torch.empty_like(x)
    es = 8 if dt == np.float32 else 16
    for ax in axes:
        n = shape[ax]
        h = nb.FftHandler(n, dt)
        lanes = x.numel() // n
        fl = lanes * 5.0 * n * math.log2(n)
        for nm, f in (("ndfft", nb.ndfft), ("ndifft", nb.ndifft)):
            med, mn = time_call(lambda: f(x, y, h, ax), iters, flush)
            report(cfg, f"{nm} axis{ax} n={n} {tuple(shape)} c{es*8}", med, mn, 2 * x.numel() * es, fl)
    del x, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    only = set(a.only.split(",")) if a.only else None
    it = a.iters

    def want(c):
        return only is None or c in only

    if want("c1"):
        n = 128
        run_c2c("c1", (n, n), (0, 1), np.float64, it, flush=True)
        x = rnd((n, n), np.float64, False)
        hr = nb.R2cFftHandler(n); hd = nb.DctHandler(n)
        for ax in (0, 1):
            so = [n, n]; so[ax] = n // 2 + 1
            y = torch.empty(so, dtype=torch.complex128, device="cuda")
            med, mn = time_call(lambda: nb.ndfft_r2c(x, y, hr, ax), it, True)
            report("c1", f"ndfft_r2c axis{ax} 128x128 f64", med, mn, x.numel() * 8 + y.numel() * 16, n * 2.5 * n * math.log2(n))
            z = torch.empty_like(x)
            med, mn = time_call(lambda: nb.nddct2(x, z, hd, ax), it, True)
            report("c1", f"nddct2 axis{ax} 128x128 f64", med, mn, 2 * x.numel() * 8, n * 2.5 * n * math.log2(n))
    if want("c2"):
        run_c2c("c2", (8192, 8192), (1, 0), np.float32, it)
    if want("c3"):
        n = 512
        x = rnd((n, n, n), np.float64, False)
        a1 = torch.empty((n, n, n // 2 + 1), dtype=torch.complex128, device="cuda"); a2 = torch.empty_like(a1)
        hr = nb.R2cFftHandler(n); hc = nb.FftHandler(n)
        fl_r = n * n * 2.5 * n * math.log2(n)
        fl_c = n * (n // 2 + 1) * 5.0 * n * math.log2(n)
        med, mn = time_call(lambda: nb.ndfft_r2c(x, a1, hr, 2), it, False)
        report("c3", "ndfft_r2c axis2 512^3 f64", med, mn, x.numel() * 8 + a1.numel() * 16, fl_r)
        med, mn = time_call(lambda: nb.ndfft(a1, a2, hc, 1), it, False)
        report("c3", "ndfft axis1 512x512x257 c128", med, mn, 2 * a1.numel() * 16, fl_c)
        med, mn = time_call(lambda: nb.ndfft(a2, a1, hc, 0), it, False)
        report("c3", "ndfft axis0 512x512x257 c128", med, mn, 2 * a1.numel() * 16, fl_c)
        med, mn = time_call(lambda: nb.ndifft_r2c(a1, x, hr, 2), it, False)
        report("c3", "ndifft_r2c axis2 512^3 f64", med, mn, x.numel() * 8 + a1.numel() * 16, fl_r)
        del x, a1, a2
    if want("c4"):
        n = 4096
        x = rnd((n, n), np.float64, False); y = torch.empty_like(x)
        h = nb.DctHandler(n)
        fl = n * 2.5 * n * math.log2(n)
        for ax in (1, 0):
            for k in (1, 2, 3, 4):
                f = getattr(nb, f"nddct{k}")
                med, mn = time_call(lambda: f(x, y, h, ax), it, False)
                report("c4", f"nddct{k} axis{ax} 4096x4096 f64", med, mn, 2 * x.numel() * 8, fl)
        del x, y
    if want("c5a"):
        run_c2c("c5a", (360, 1000, 384), (0, 1, 2), np.float64, max(3, it // 2))
    if want("c5b"):
        n, b = 1 << 24, 64
        x = rnd((b, n), np.float32, True); y = torch.empty_like(x)
        h = nb.FftHandler(n, np.float32)
        med, mn = time_call(lambda: nb.ndfft(x, y, h, 1), max(3, it // 3), False)
        report("c5b", "ndfft axis1 64x2^24 c64 (four-step; 1-pass byte definition)", med, mn, 2 * x.numel() * 8, b * 5.0 * n * 24)
        del x, y


if __name__ == "__main__":
    main()
