#!/usr/bin/env python3
"""bench.py — axis-transform hot path of ndrustfft on B200.

HEADLINE WORKLOAD (every N): BASELINE.json config c3, the 512^3 f64 real 3-D spectral transform
    ndfft_r2c axis 2  ->  ndfft axis 1  ->  [slab exchange, N > 1]  ->  ndfft axis 0          (examples/rfft2.rs:29-33 pattern)
It is the one config the north_star gives a multi-GPU target for, and the driver derives scaling efficiency from the
per-N `value`s of this file, so the same global transform is timed at N = 1, 2, 4, 8: "scaling": "strong".
  N = 1   three launches on one GPU (ndrustfft_b200.dist.SlabR2cFft3d with one rank = r2c, axis-1, axis-0 pass).
  N > 1   axis-0 slabs per rank, the axis-1 pass's stores ARE the all-to-all (peer-mapped receive buffers over NVLink,
          ndfb_exec_scatter_out), then the axis-0 pass on the rank's axis-1 slab; device time, max over ranks.
`value`        GFLOP/s of the ONE global transform (2.5 n log2 n per real lane, 5 n log2 n per complex lane = 9.08e9).
`e2e`          the same transform from PAGEABLE host arrays (plain np.empty, what ndarray's as_ptr() hands the shim):
               N = 1 one ndfb_exec_chain(mem=HOST) call per step; N > 1 every rank uploads its slab, transforms, downloads.
`roofline`     the dominant kernel of the step (largest share of the device time) at N = 1; at N > 1 also the NVLink side.
`configs`      (N = 1) every other BASELINE config, per call: c1 (incl. the criterion ramp of benches/ndrustfft.rs:9-60 and
               a CUDA-graph replay), c2 (4 calls + the step-weighted fraction), c3 stages, c4 (8 calls), c5a, c5b — each
               with ms, GB/s of algorithmic bytes, fraction of the measured HBM peak, GFLOP/s and the scipy CPU stand-in.
`c2_weak`      (N > 1) round 1's headline (four axis transforms of one 8192^2 c64 array per rank, no collective).
`cpu_baseline` / `--impl reference`: ndrustfft cannot be built here (Rust; rustfft/realfft/rustdct not vendored; the
               harness probes `cargo --version` and says so), so the oracle port (scipy.fft/pocketfft, all host threads =
               what `_par` does with rayon) runs the same three axis passes on the full 512^3 array.  kind = "port".
"""
import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N3 = 512
M3 = N3 // 2 + 1
C3_FLOPS = N3 * N3 * 2.5 * N3 * math.log2(N3) + 2 * N3 * M3 * 5.0 * N3 * math.log2(N3)      # 9.084e9
C3_STAGE_BYTES = [N3 ** 3 * 8 + N3 * N3 * M3 * 16, 2 * N3 * N3 * M3 * 16, 2 * N3 * N3 * M3 * 16]
C3_STAGE_FLOPS = [N3 * N3 * 2.5 * N3 * math.log2(N3), N3 * M3 * 5.0 * N3 * math.log2(N3), N3 * M3 * 5.0 * N3 * math.log2(N3)]
C3_STAGE_NAMES = ["ndfft_r2c axis2 512^3 f64 -> 512x512x257 c128", "ndfft axis1 512x512x257 c128", "ndfft axis0 512x512x257 c128"]
METRIC = "GFLOP/s (2.5 n log2 n real / 5 n log2 n complex lanes) of the 512^3 f64 real 3-D transform (c3: ndfft_r2c + 2 x ndfft)"


def config_of(n_gpus):
    return {"workload": "c3: 512^3 f64 real 3-D spectral transform = ndfft_r2c axis 2, ndfft axis 1, ndfft axis 0 (one step = the three axis passes of ONE global array)",
            "shape": [N3, N3, N3], "l2": "inputs larger than L2 (1 GiB real input, 1.08 GB complex spectrum)",
            "decomposition": "single GPU" if n_gpus == 1 else f"axis-0 slabs over {n_gpus} ranks, one exchange to axis-1 slabs before the last pass"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cargo_probe():
    """the intended CPU arm is ndrustfft's own criterion benches; needs a Rust toolchain AND the crates."""
    exe = shutil.which("cargo")
    if not exe:
        return {"cargo": None, "note": "no cargo/rustc in this image: ndrustfft (rustfft 6.1 / realfft 3.2 / rustdct 0.7, not vendored) cannot be built; CPU arm = oracle port over scipy.fft"}
    try:
        v = subprocess.run([exe, "--version"], capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception as e:  # pragma: no cover
        v = repr(e)
    return {"cargo": v, "note": "cargo present, but the reference's crates are not vendored and there is no network: still the oracle port"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (scipy.fft over lanes, all host threads) on the same workload
# ------------------------------------------------------------------------------------------------------
def cpu_c3_step_factory(frac=1.0):
    """One c3 step on the host: r2c axis 2, fft axis 1, fft axis 0 of a (512*frac, 512, 512) slab... frac < 1 keeps the
    lane LENGTHS (512 on every axis) and cuts the lane COUNT: axes 2 and 1 run on n0*frac planes, axis 0 on n1*frac columns."""
    import numpy as np
    from oracle import ndrustfft_oracle as orc
    rng = np.random.default_rng(0xB200 + 48)
    p0 = max(1, int(round(N3 * frac)))
    x = rng.uniform(-1, 1, (p0, N3, N3))
    a = np.empty((p0, N3, M3), np.complex128)
    b = np.empty_like(a)
    c_in = np.empty((N3, p0, M3), np.complex128)
    c_in[...] = 0.5
    c_out = np.empty_like(c_in)
    hr, hc = orc.R2cFftHandler(N3), orc.FftHandler(N3)

    def step():
        orc.ndfft_r2c_par(x, a, hr, 2)
        orc.ndfft_par(a, b, hc, 1)
        orc.ndfft_par(c_in, c_out, hc, 0)

    return step, C3_FLOPS * p0 / N3, p0


def cpu_c3_run(steps, warmup, budget_s=150.0):
    cores = host_threads()
    step, flops, p0 = cpu_c3_step_factory(1.0)
    t0 = time.perf_counter(); step(); first = time.perf_counter() - t0
    if first * (steps + warmup) > budget_s:          # bounded sample: a quarter of the lanes of every pass
        step, flops, p0 = cpu_c3_step_factory(0.25)
        step()
    for _ in range(max(0, warmup - 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = (f"full 512^3 array ({p0} of 512 planes per pass)" if p0 == N3 else f"{p0} of 512 planes per pass (lane lengths unchanged)") + \
        f", scipy.fft (pocketfft) f64, workers={cores}"
    return flops / dt / 1e9, dt, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to its ranks, which throttles pocketfft's worker pool (measured: 3-5x slower): the CPU
    # arm gets every host core it can use, set before numpy / scipy are first imported
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
    K, W = max(1, args.steps), max(0, args.warmup)
    gf, dt, cores, sample = cpu_c3_run(K, W)
    line = {
        "impl": "reference", "metric": METRIC, "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(args.gpus),
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "ndrustfft itself is not buildable in this image; oracle port = scipy.fft over lanes with all host threads (what `_par` does with rayon)",
                         "toolchain": cargo_probe()},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
# per-config measurements (N = 1): device-resident, CUDA events, median of `iters`
# ------------------------------------------------------------------------------------------------------
class ConfigBench:
    def __init__(self, nb, torch, np, peak, iters, cpu):
        self.nb, self.torch, self.np, self.peak, self.iters, self.cpu = nb, torch, np, peak, iters, cpu
        self.flush_buf = None
        self.rows = []

    def flush(self):
        if self.flush_buf is None:
            self.flush_buf = self.torch.empty(256 << 20, dtype=self.torch.uint8, device="cuda")
        self.flush_buf.zero_()

    def time_call(self, fn, iters=None, flush=False):
        """Median / minimum device time of one call.  A single call bracketed by two events on an idle GPU also times the host
        side of the call (Python -> C ABI -> launch, ~7 us: a tenth of a 0.1 ms kernel), so calls shorter than ~0.5 ms are
        enqueued `reps` times back to back between the events (the arrays are larger than L2; reps = 1 with `flush`)."""
        t = self.torch
        for _ in range(3):
            fn()
        t.cuda.synchronize()
        reps = 1
        if not flush:
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            t.cuda.synchronize()
            reps = max(1, min(8, int(round(0.6 / max(e0.elapsed_time(e1), 1e-3)))))
        ts = []
        for _ in range(iters or self.iters):
            if flush:
                self.flush()
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            t.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    def rnd(self, shape, dt, cx):
        t = self.torch
        rt = t.float32 if dt == self.np.float32 else t.float64
        if cx:
            return t.complex(t.rand(shape, device="cuda", dtype=rt) * 2 - 1, t.rand(shape, device="cuda", dtype=rt) * 2 - 1)
        return t.rand(shape, device="cuda", dtype=rt) * 2 - 1

    def cpu_time(self, fn, flops, reps=1):
        if not self.cpu:
            return None
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt = (time.perf_counter() - t0) / reps
        return {"ms": round(dt * 1e3, 3), "GFLOP/s": round(flops / dt / 1e9, 2)}

    def add(self, cfg, call, dtype, ms, ms_min, nbytes, flops, cpu=None, **extra):
        gbs = nbytes / (ms * 1e-3) / 1e9
        row = {"cfg": cfg, "call": call, "dtype": dtype, "ms": round(ms, 5), "ms_min": round(ms_min, 5), "GB/s": round(gbs, 1),
               "frac": round(gbs / self.peak, 4), "GFLOP/s": round(flops / (ms * 1e-3) / 1e9, 1)}
        if cpu:
            row["cpu"] = cpu
        row.update(extra)
        self.rows.append(row)
        return row

    # ---- c1: 128 x 128 f64, the reference's own bench shape (launch-latency bound) ----
    def c1(self):
        nb, np, t = self.nb, self.np, self.torch
        from oracle import ndrustfft_oracle as orc
        n = 128
        xc = self.rnd((n, n), np.float64, True); yc = t.empty_like(xc)
        xr = self.rnd((n, n), np.float64, False); yr = t.empty_like(xr)
        hc, hr, hd = nb.FftHandler(n), nb.R2cFftHandler(n), nb.DctHandler(n)
        xc_h, xr_h = xc.cpu().numpy(), xr.cpu().numpy()
        oc, orr, od = orc.FftHandler(n), orc.R2cFftHandler(n), orc.DctHandler(n)
        fl_c, fl_r = n * 5.0 * n * math.log2(n), n * 2.5 * n * math.log2(n)
        for ax in (0, 1):
            so = [n, n]; so[ax] = n // 2 + 1
            yh = t.empty(so, dtype=t.complex128, device="cuda")
            cases = [("ndfft", lambda: nb.ndfft(xc, yc, hc, ax), 2 * n * n * 16, fl_c, lambda: orc.ndfft(xc_h, np.empty_like(xc_h), oc, ax)),
                     ("ndfft_r2c", lambda: nb.ndfft_r2c(xr, yh, hr, ax), n * n * 8 + yh.numel() * 16, fl_r,
                      lambda: orc.ndfft_r2c(xr_h, np.empty(so, np.complex128), orr, ax)),
                     ("nddct2", lambda: nb.nddct2(xr, yr, hd, ax), 2 * n * n * 8, fl_r, lambda: orc.nddct2(xr_h, np.empty_like(xr_h), od, ax))]
            for name, fn, nbytes, fl, cfn in cases:
                med, mn = self.time_call(fn, flush=True)
                # the same call replayed 200x from one CUDA graph: the per-call cost without host launch overhead
                g = t.cuda.CUDAGraph()
                side = t.cuda.Stream()
                side.wait_stream(t.cuda.current_stream())
                with t.cuda.stream(side):
                    fn()
                    with t.cuda.graph(g, stream=side):
                        for _ in range(200):
                            fn()
                t.cuda.current_stream().wait_stream(side)
                g.replay(); t.cuda.synchronize()
                e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); t.cuda.synchronize()
                graph_us = e0.elapsed_time(e1) * 1e3 / 200
                self.add("c1", f"{name} axis{ax} 128x128 f64", "f64", med, mn, nbytes, fl, cpu=self.cpu_time(cfn, fl, reps=20),
                         us_per_call_in_cuda_graph=round(graph_us, 3), note="launch-latency bound: 0.04-0.08 us of HBM time")
        # host-array calls (what a Rust caller with ndarray memory pays per call: two PCIe copies + launch)
        xh = np.ascontiguousarray(xc_h); yh = np.empty_like(xh)
        nb.ndfft(xh, yh, hc, 0)
        t0 = time.perf_counter()
        for _ in range(50):
            nb.ndfft(xh, yh, hc, 0)
        self.rows.append({"cfg": "c1", "call": "ndfft axis0 128x128 c128 through ndfb_exec(mem=HOST), pageable numpy arrays",
                          "dtype": "f64", "us_per_call_host": round((time.perf_counter() - t0) / 50 * 1e6, 2)})
        # criterion ramp of the reference's benches (benches/ndrustfft.rs:9-60): n x n f64, axis 0; DCT-I at 2^k + 1
        for n in (128, 264, 512, 1024):
            x = self.rnd((n, n), np.float64, True); y = t.empty_like(x)
            h = nb.FftHandler(n)
            med, mn = self.time_call(lambda: nb.ndfft(x, y, h, 0), flush=True)
            xh_ = x.cpu().numpy(); oh = orc.FftHandler(n)
            fl = n * 5.0 * n * math.log2(n)
            self.add("c1-ramp", f"ndfft axis0 {n}x{n} c128", "f64", med, mn, 2 * n * n * 16, fl,
                     cpu=self.cpu_time(lambda: orc.ndfft(xh_, np.empty_like(xh_), oh, 0), fl, reps=5))
            nd = n + 1
            xr_ = self.rnd((nd, nd), np.float64, False); yr_ = t.empty_like(xr_)
            hd_ = nb.DctHandler(nd)
            med, mn = self.time_call(lambda: nb.nddct1(xr_, yr_, hd_, 0), flush=True)
            xrh = xr_.cpu().numpy(); odh = orc.DctHandler(nd)
            fl = nd * 2.5 * nd * math.log2(nd)
            self.add("c1-ramp", f"nddct1 axis0 {nd}x{nd} f64", "f64", med, mn, 2 * nd * nd * 8, fl,
                     cpu=self.cpu_time(lambda: orc.nddct1(xrh, np.empty_like(xrh), odh, 0), fl, reps=5))

    # ---- c2: 8192 x 8192 c64, both axes, forward and inverse ----
    def c2(self):
        nb, np, t = self.nb, self.np, self.torch
        from oracle import ndrustfft_oracle as orc
        n = 8192
        x = self.rnd((n, n), np.float32, True); a = t.empty_like(x); b = t.empty_like(x)
        h = nb.FftHandler(n, np.float32)
        fl, nbytes = n * 5.0 * n * math.log2(n), 2 * n * n * 8
        cpu = {}
        if self.cpu:
            orc.set_native_precision(True)
            xh = x.cpu().numpy(); yh = np.empty_like(xh); oh = orc.FftHandler(n, np.float32)
            for ax in (1, 0):
                cpu[ax] = self.cpu_time(lambda: orc.ndfft_par(xh, yh, oh, ax), fl)
            orc.set_native_precision(False)
            del xh, yh
        calls = [("ndfft axis1 (contiguous)", lambda: nb.ndfft(x, a, h, 1), 1), ("ndfft axis0 (strided)", lambda: nb.ndfft(a, b, h, 0), 0),
                 ("ndifft axis0 (strided)", lambda: nb.ndifft(b, a, h, 0), 0), ("ndifft axis1 (contiguous)", lambda: nb.ndifft(a, b, h, 1), 1)]
        tot = 0.0
        for name, fn, ax in calls:
            med, mn = self.time_call(fn)
            tot += med
            self.add("c2", f"{name} 8192x8192 c64", "f32", med, mn, nbytes, fl, cpu=cpu.get(ax))
        rel = (t.linalg.vector_norm(b - x) / t.linalg.vector_norm(x)).item()
        assert rel < 1e-5, rel
        self.rows.append({"cfg": "c2", "call": "step = the four calls above", "dtype": "f32", "ms": round(tot, 5),
                          "GFLOP/s": round(4 * fl / (tot * 1e-3) / 1e9, 1), "step_weighted_frac": round(4 * nbytes / (tot * 1e-3) / 1e9 / self.peak, 4),
                          "roundtrip_rel_l2": rel})
        del x, a, b

    # ---- c3 stages through the public single calls (unpadded arrays), forward and the inverse r2c ----
    def c3(self):
        nb, np, t = self.nb, self.np, self.torch
        n = N3
        x = self.rnd((n, n, n), np.float64, False)
        a1 = t.empty((n, n, M3), dtype=t.complex128, device="cuda"); a2 = t.empty_like(a1)
        hr, hc = nb.R2cFftHandler(n), nb.FftHandler(n)
        for name, fn, nbytes, fl in (
                ("ndfft_r2c axis2 512^3 f64", lambda: nb.ndfft_r2c(x, a1, hr, 2), C3_STAGE_BYTES[0], C3_STAGE_FLOPS[0]),
                ("ndfft axis1 512x512x257 c128", lambda: nb.ndfft(a1, a2, hc, 1), C3_STAGE_BYTES[1], C3_STAGE_FLOPS[1]),
                ("ndfft axis0 512x512x257 c128", lambda: nb.ndfft(a2, a1, hc, 0), C3_STAGE_BYTES[2], C3_STAGE_FLOPS[2]),
                ("ndifft axis0 512x512x257 c128", lambda: nb.ndifft(a1, a2, hc, 0), C3_STAGE_BYTES[2], C3_STAGE_FLOPS[2]),
                ("ndifft_r2c axis2 512^3 f64", lambda: nb.ndifft_r2c(a1, x, hr, 2), C3_STAGE_BYTES[0], C3_STAGE_FLOPS[0])):
            med, mn = self.time_call(fn)
            self.add("c3", name, "f64", med, mn, nbytes, fl)
        del x, a1, a2

    # ---- c4: DCT-I..IV on 4096 x 4096 f64, both axes ----
    def c4(self):
        nb, np, t = self.nb, self.np, self.torch
        from oracle import ndrustfft_oracle as orc
        n = 4096
        x = self.rnd((n, n), np.float64, False); y = t.empty_like(x); z = t.empty_like(x)
        h = nb.DctHandler(n)
        fl, nbytes = n * 2.5 * n * math.log2(n), 2 * n * n * 8
        xh = x.cpu().numpy() if self.cpu else None
        oh = orc.DctHandler(n)
        for ax in (1, 0):
            for k in (1, 2, 3, 4):
                f = getattr(nb, f"nddct{k}")
                med, mn = self.time_call(lambda: f(x, y, h, ax))
                cpu = self.cpu_time(lambda: getattr(orc, f"nddct{k}_par")(xh, np.empty_like(xh), oh, ax), fl) if self.cpu else None
                self.add("c4", f"nddct{k} axis{ax} 4096x4096 f64", "f64", med, mn, nbytes, fl, cpu=cpu)
            nb.nddct2(x, y, h, ax); nb.nddct3(y, z, h, ax)       # Chebyshev round trip: dct3(dct2(x)) = 2n x
            rel = (t.linalg.vector_norm(z / (2.0 * n) - x) / t.linalg.vector_norm(x)).item()
            assert rel < 1e-12, rel
        del x, y, z

    # ---- c5a: 360 x 1000 x 384 c128, every axis ----
    def c5a(self):
        nb, np, t = self.nb, self.np, self.torch
        from oracle import ndrustfft_oracle as orc
        shape = (360, 1000, 384)
        x = self.rnd(shape, np.float64, True); y = t.empty_like(x)
        for ax in (0, 1, 2):
            n = shape[ax]
            h = nb.FftHandler(n)
            lanes = x.numel() // n
            fl = lanes * 5.0 * n * math.log2(n)
            cpu = None
            if self.cpu:   # a quarter of the lanes (cut along another axis), lane length unchanged
                cut = [slice(None)] * 3
                cut[(ax + 1) % 3] = slice(0, shape[(ax + 1) % 3] // 4)
                xs = np.ascontiguousarray(x[tuple(cut)].cpu().numpy()); ys = np.empty_like(xs); oh = orc.FftHandler(n)
                cpu = self.cpu_time(lambda: orc.ndfft_par(xs, ys, oh, ax), fl / 4)
                cpu["sample"] = "1/4 of the lanes"
                del xs, ys
            for nm, f in (("ndfft", nb.ndfft), ("ndifft", nb.ndifft)):
                med, mn = self.time_call(lambda: f(x, y, h, ax), iters=max(3, self.iters // 2))
                self.add("c5a", f"{nm} axis{ax} n={n} 360x1000x384 c128", "f64", med, mn, 2 * x.numel() * 16, fl, cpu=cpu if nm == "ndfft" else None)
        del x, y
        # the shape above never needs Bluestein: a 1009-point (prime) axis of a comparable array
        x = self.rnd((1009, 4096), np.float64, True); y = t.empty_like(x)
        h = nb.FftHandler(1009)
        med, mn = self.time_call(lambda: nb.ndfft(x, y, h, 0), iters=max(3, self.iters // 2))
        self.add("c5a+", "ndfft axis0 n=1009 (prime: fused Bluestein) 1009x4096 c128", "f64", med, mn, 2 * x.numel() * 16, 4096 * 5.0 * 1009 * math.log2(1009))
        del x, y

    # ---- c5b: 2^24-point rows, batch 64, c64 (multi-pass through a device workspace) ----
    def c5b(self):
        nb, np, t = self.nb, self.np, self.torch
        from oracle import ndrustfft_oracle as orc
        n, b = 1 << 24, 64
        x = self.rnd((b, n), np.float32, True); y = t.empty_like(x)
        h = nb.FftHandler(n, np.float32)
        fl = b * 5.0 * n * 24
        cpu = None
        if self.cpu:
            orc.set_native_precision(True)
            xs = x[:4].cpu().numpy(); ys = np.empty_like(xs); oh = orc.FftHandler(n, np.float32)
            cpu = self.cpu_time(lambda: orc.ndfft_par(xs, ys, oh, 1), fl * 4 / b)
            cpu["sample"] = "4 of 64 rows"
            orc.set_native_precision(False)
            del xs, ys
        med, mn = self.time_call(lambda: nb.ndfft(x, y, h, 1), iters=max(3, self.iters // 3))
        nbytes = 2 * x.numel() * 8
        self.add("c5b", "ndfft axis1 64 x 2^24 c64 (multi-pass: 256 x 256 x 256)", "f32", med, mn, nbytes, fl, cpu=cpu,
                 frac_of_two_pass_bound=round(2 * nbytes / (med * 1e-3) / 1e9 / self.peak, 4),
                 note="frac uses the one-pass byte definition; at least two HBM passes are unavoidable at this length, and the library takes "
                      "three over 256-byte rows (two passes over 32-byte rows measure slower: DESIGN.md 4.4)")
        del x, y


def c2_step_bench(nb, torch, np, dev, local, steps, warmup):
    """Round 1's headline: ndfft axis 1, ndfft axis 0, ndifft axis 0, ndifft axis 1 on one 8192^2 c64 array per rank."""
    n = 8192
    h = nb.FftHandler(n, np.float32, device=local)
    g = torch.Generator(device=dev); g.manual_seed(0xB200 + 32 + local)
    x = torch.complex(torch.rand((n, n), generator=g, device=dev) * 2 - 1, torch.rand((n, n), generator=g, device=dev) * 2 - 1)
    a = torch.empty_like(x); b = torch.empty_like(x)

    def step():
        nb.ndfft(x, a, h, 1); nb.ndfft(a, b, h, 0); nb.ndifft(b, a, h, 0); nb.ndifft(a, b, h, 1)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# per-launch DRAM traffic of the dominant kernel from one `ncu --set full` capture of this command (profiles/)
C3_TRAFFIC = {"bytes": 2.156e9, "source": "ncu --set full, profiles/round2/r2e_ncu_c3_summary.txt: sfft_kernel<double, Sched<512,64,8,8,8>, cols> dram__bytes_read 1.116 GB + "
                                              "dram__bytes_write 1.039 GB per launch = 1.00 x the algorithmic 2.156 GB; the r2c kernel: 1.107 + 1.027 GB for 2.152 GB"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU stand-in legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config table (c1, c2, c4, c5a, c5b)")
    ap.add_argument("--only", default="", help="comma list of configs for the table, e.g. c2,c4")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--chunks", type=int, default=int(os.environ.get("NDFB_C3_CHUNKS", "0")), help="i2-chunks of the N>1 exchange pipeline (0 = auto)")
    ap.add_argument("--no-graph", action="store_true", help="N>1: launch the slab pipeline from the host instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import ndrustfft_b200 as nb
    from ndrustfft_b200.dist import SlabR2cFft3d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and "NDFB_HOST_THREADS" not in os.environ:
        # the ranks of one node share its cores: split them between the ranks' staging-copy pools
        os.environ["NDFB_HOST_THREADS"] = str(max(2, host_threads() // world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    peak, peak_src = measured_peak()
    lib = nb._default_backend().lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        tt = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item()

    # ---- the headline transform ----
    # i2-chunked overlap of the exchange with the last pass loses at every N measured (8 GPUs, CUDA graph: 0.295 ms unchunked,
    # 0.315 / 0.341 / 0.383 ms with 2 / 4 / 8 chunks; profiles/round2/r2b_c3_probe_n8.jsonl), so the default pipeline is unchunked
    chunks = args.chunks or 1
    plan = SlabR2cFft3d((N3, N3, N3), np.float64, device=dev, chunks=chunks)
    s0, s1 = N3 // world, N3 // world
    g = torch.Generator(device=dev); g.manual_seed(0xB200 + 48 + rank)
    x = torch.rand((s0, N3, N3), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    for _ in range(W):
        out = plan.forward(x)              # result = view of the plan's padded array (aligned tile rows in the last pass)
    back = plan.inverse(out)
    rel = (torch.linalg.vector_norm(back - x) / torch.linalg.vector_norm(x)).item()
    assert rel < 1e-12, f"round trip rel L2 {rel}"
    del back
    barrier()
    lc0 = lib.launch_count()
    plan.forward(x)
    torch.cuda.synchronize()
    launches_per_step = lib.launch_count() - lc0
    use_graph = world > 1 and not args.no_graph
    run = lambda: plan.forward(x)
    per_replay = 1
    if use_graph:
        try:
            if getattr(plan, "peer", False) and plan._call % 2:
                plan.forward(x)                           # keep the double-buffer parity of capture and replay aligned
            gr = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                with torch.cuda.graph(gr, stream=side):
                    plan.forward(x)
                    plan.forward(x)                       # two transforms per replay: both receive buffers
            torch.cuda.current_stream(dev).wait_stream(side)
            per_replay = 2
            run = lambda: gr.replay()
            run(); torch.cuda.synchronize()
        except Exception as e:                            # pragma: no cover - depends on the box
            use_graph = False
            per_replay = 1
            run = lambda: plan.forward(x)
            if rank == 0:
                print(f"[bench] CUDA-graph capture of the slab pipeline failed ({e!r}); host launches", file=sys.stderr)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stage_ms = None
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nrep = (K + per_replay - 1) // per_replay
    if world == 1:
        # same three launches, with an event between the stages
        be = plan.be
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
        t_start.record()
        for k in range(K):
            evs[k][0].record(); be.ndfft_r2c(x, plan.a, plan.h2, 2)
            evs[k][1].record(); be.ndfft(plan.a_pad, plan.b_pad, plan.h1, 1)
            evs[k][2].record(); be.ndfft(plan.b_pad, plan.out_pad, plan.h0, 0)
            evs[k][3].record()
        t_end.record()
        barrier()
        stage_ms = [sum(evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(K)) / K for i in range(3)]
        steps_done = K
    else:
        if getattr(plan, "peer", False):
            plan._symm[0][1].barrier()             # device-side rendezvous: every GPU enters the timed region within microseconds
        t_start.record()
        for _ in range(nrep):
            run()
        t_end.record()
        barrier()
        steps_done = nrep * per_replay
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = allmax(t_start.elapsed_time(t_end) / steps_done)
    value = C3_FLOPS / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: pageable host arrays in, pageable host arrays out ----
    e2e = None
    if not args.no_e2e:
        KE = max(1, min(args.e2e_steps, K))
        rng = np.random.default_rng(0xB200 + 48 + rank)
        if world == 1:
            hx = rng.uniform(-1, 1, (N3, N3, N3))                      # plain pageable numpy memory
            hy = np.empty((N3, N3, M3), np.complex128)
            hr, hc = nb.R2cFftHandler(N3), nb.FftHandler(N3)
            chain = [("ndfft_r2c", hr, 2), ("ndfft", hc, 1), ("ndfft", hc, 0)]
            nb.ndchain(hx, hy, chain)
            t0 = time.perf_counter()
            for _ in range(KE):
                nb.ndchain(hx, hy, chain)
            dt = (time.perf_counter() - t0) / KE
            # spot check against the device result of the same data
            xd = torch.from_numpy(hx).to(dev); od = torch.empty((N3, N3, M3), dtype=torch.complex128, device=dev)
            nb.ndchain(xd, od, chain)
            relh = (torch.linalg.vector_norm(torch.from_numpy(hy[:8]).to(dev) - od[:8]) / torch.linalg.vector_norm(od[:8])).item()
            assert relh < 1e-12, relh
            del xd, od
            # the same with caller-pinned arrays (an extra: what a caller that registers its memory gets)
            px = torch.empty((N3, N3, N3), dtype=torch.float64).pin_memory(); py = torch.empty((N3, N3, M3), dtype=torch.complex128).pin_memory()
            px.copy_(torch.from_numpy(hx))
            nb.ndchain(px.numpy(), py.numpy(), chain)
            t0 = time.perf_counter()
            for _ in range(max(2, KE // 2)):
                nb.ndchain(px.numpy(), py.numpy(), chain)
            dtp = (time.perf_counter() - t0) / max(2, KE // 2)
            del px, py
            e2e = {"value": C3_FLOPS / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": N3 ** 3 * 8, "d2h_bytes_per_step": N3 * N3 * M3 * 16,
                   "ms_per_step": dt * 1e3, "steps": KE,
                   "path": "ndfb_exec_chain(mem=HOST) on PAGEABLE numpy arrays: caller memory -> pinned ring (copy threads) -> H2D | r2c pieces ... axis-0 pieces | D2H -> caller memory",
                   "pinned_arrays": {"value": C3_FLOPS / dtp / 1e9, "ms_per_step": dtp * 1e3, "note": "same call on caller-pinned arrays (no staging ring)"}}
            del hx, hy
        else:
            hx = rng.uniform(-1, 1, (s0, N3, N3))
            hy = np.empty((N3, s1, M3), np.complex128)

            def host_step():
                plan.forward_host(hx, hy)          # pageable slab in, pageable spectrum slab out (pinned ring + copy threads)
            host_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(KE):
                host_step()
            torch.cuda.synchronize()
            dt = allmax(time.perf_counter() - t0) / KE
            e2e = {"value": C3_FLOPS / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": N3 ** 3 * 8, "d2h_bytes_per_step": N3 * N3 * M3 * 16,
                   "ms_per_step": dt * 1e3, "steps": KE,
                   "path": "per rank: SlabR2cFft3d.forward_host = pageable numpy slab -> pinned ring -> device, forward, device -> pinned ring -> pageable numpy; wall clock, max over ranks",
                   "host_copy_threads_per_rank": int(os.environ.get("NDFB_HOST_THREADS", "0"))}
            del hx, hy

    # ---- c2 weak line (N > 1) ----
    c2_weak = None
    if world > 1:
        ms_c2 = allmax(c2_step_bench(nb, torch, np, dev, local, max(3, K // 2), 3))
        fl = 4 * 8192 * 5.0 * 8192 * 13
        c2_weak = {"value": world * fl / (ms_c2 * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": ms_c2, "scaling": "weak",
                   "workload": "c2 step (4 axis transforms of an 8192^2 c64 array) on an independent array per rank, no collective"}

    # ---- every other BASELINE config, per call (N = 1) ----
    configs = None
    if world == 1 and not args.no_configs:
        del x, out, plan
        torch.cuda.empty_cache()
        cb = ConfigBench(nb, torch, np, peak, args.iters, cpu=not args.no_cpu)
        only = set(args.only.split(",")) if args.only else None
        for name in ("c1", "c2", "c3", "c4", "c5a", "c5b"):
            if only is None or name in only:
                getattr(cb, name)()
                lib.dll.ndfb_release_workspaces()
                torch.cuda.empty_cache()
        configs = cb.rows

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": steps_done, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_of(world),
            "gpu_launches": launches_per_step * steps_done, "launches_per_step": launches_per_step,
            "clocks": clocks, "e2e": e2e, "roundtrip_rel_l2": rel, "library": lib.version(),
        }
        step_bytes = sum(C3_STAGE_BYTES)
        if world == 1:
            dom = max(range(3), key=lambda i: stage_ms[i])
            ach = C3_STAGE_BYTES[dom] / (stage_ms[dom] * 1e-3) / 1e9
            kernels = ["rsfft_kernel<double, Sched<256,...>, rows, RK_R2C> (r2c of 512-point real rows: 256-point complex core + paired epilogue)",
                       "sfft_kernel<double, Sched<512,64,8,8,8>, cols> (512-point c128 columns, stride 264 elements)",
                       "sfft_kernel<double, Sched<512,64,8,8,8>, cols> (512-point c128 columns, stride 512*264 elements)"]
            line["roofline"] = {"bound": "hbm", "kernel": kernels[dom], "stage": C3_STAGE_NAMES[dom], "achieved": ach, "peak": peak, "unit": "GB/s",
                                "frac": ach / peak, "traffic": C3_TRAFFIC["bytes"], "traffic_source": C3_TRAFFIC["source"], "peak_source": peak_src,
                                "algorithmic_bytes_per_launch": C3_STAGE_BYTES[dom], "share_of_step": stage_ms[dom] / sum(stage_ms),
                                "step_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                                "stages": [{"name": C3_STAGE_NAMES[i], "ms": stage_ms[i], "GB/s": C3_STAGE_BYTES[i] / (stage_ms[i] * 1e-3) / 1e9,
                                            "frac": C3_STAGE_BYTES[i] / (stage_ms[i] * 1e-3) / 1e9 / peak,
                                            "GFLOP/s": C3_STAGE_FLOPS[i] / (stage_ms[i] * 1e-3) / 1e9} for i in range(3)]}
        else:
            sent = plan.bytes_sent_per_rank()
            line["roofline"] = {"bound": "hbm", "achieved": step_bytes / world / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": step_bytes / world / (ms_per_step * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                                "note": "per-GPU algorithmic HBM bytes of the three passes / step time (the step also contains the exchange)",
                                "nvlink": {"bytes_sent_per_rank": sent, "peak_GBps_per_direction": 770.0, "time_at_peak_ms": sent / 770e9 * 1e3,
                                           "frac_of_step": sent / 770e9 * 1e3 / ms_per_step,
                                           "exchange": ("peer stores fused into the axis-1 kernel (ndfb_exec_scatter_out over symmetric memory)"
                                                        if getattr(plan, "peer", False) else "NCCL all_to_all_single")}}
            line["pipeline"] = {"chunks": chunks, "cuda_graph": use_graph}
            line["c2_weak"] = c2_weak
        if configs is not None:
            line["configs"] = configs
        if world == 1 and not args.no_cpu:
            gf, dt, cores, sample = cpu_c3_run(2, 1)
            line["cpu_baseline"] = {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample, "toolchain": cargo_probe()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
