#!/usr/bin/env python3
"""bench.py — BASELINE.json config c2: 2-D complex f32 ndfft/ndifft along both axes of 8192 x 8192.

One STEP = the four axis transforms of that config on one synthetic array, all through the public API
(ndrustfft_b200.ndfft / ndifft  ->  C ABI  ->  sm_100a kernels):

    ndfft  axis 1 (contiguous rows, path A)      x -> a
    ndfft  axis 0 (stride 8192 elements, path B) a -> b
    ndifft axis 0                                b -> a
    ndifft axis 1                                a -> b     (b == x up to rounding)

`value`   whole-job GFLOP/s (5 N log2 N per lane), inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`     the same step through the HOST-array path of the C ABI (pinned numpy arrays in, pinned numpy arrays
          out): H2D + kernel + D2H inside the timed region, every call.
`roofline` for the slowest of the four launches: algorithmic bytes (input once + output once = 1 GiB) / its
          CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` / `--impl reference`: the reference's CPU path cannot be built here (Rust + un-vendored crates, no
          toolchain), so the oracle port (scipy.fft/pocketfft over lanes, all host threads — what `_par` does with
          rayon) is timed on a bounded lane sample of the same workload.  kind = "port".

N > 1 (torchrun): lanes are independent (src/lib.rs:120-124), so each rank transforms its own 8192 x 8192 array
with no collective on the data path — "scaling": "weak".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_AXIS = 8192
FLOPS_PER_TRANSFORM = N_AXIS * 5.0 * N_AXIS * math.log2(N_AXIS)   # 8192 lanes x 5 n log2 n = 4.362e9
BYTES_PER_TRANSFORM = 2 * N_AXIS * N_AXIS * 8                      # read once + write once = 1 GiB
STEP_NAMES = ["ndfft axis1 (contiguous)", "ndfft axis0 (strided)", "ndifft axis0 (strided)", "ndifft axis1 (contiguous)"]
METRIC = "GFLOP/s (5*N*log2N) per axis transform, c2: 8192x8192 c64 ndfft/ndifft both axes"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
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
        time.sleep(0.15)
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


def cpu_port_run(steps, warmup, sample_lanes=1024):
    """The oracle port on host cores: scipy.fft over a bounded lane sample of the c2 step (all threads)."""
    import numpy as np
    from oracle import ndrustfft_oracle as orc
    orc.set_native_precision(True)   # time complex64 arithmetic, as rustfft on Complex<f32> would
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0xB200 + 32)
    rows = (rng.uniform(-1, 1, (sample_lanes, N_AXIS)) + 1j * rng.uniform(-1, 1, (sample_lanes, N_AXIS))).astype(np.complex64)
    cols = np.ascontiguousarray(rows.T)            # (8192, sample) : axis-0 lanes with stride = sample elements
    h = orc.FftHandler(N_AXIS, np.float32)
    ra, ca = np.empty_like(rows), np.empty_like(cols)
    rb, cb = np.empty_like(rows), np.empty_like(cols)

    def step():
        orc.ndfft_par(rows, ra, h, 1)
        orc.ndfft_par(cols, ca, h, 0)
        orc.ndifft_par(ca, cb, h, 0)
        orc.ndifft_par(ra, rb, h, 1)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    flops = 4 * sample_lanes * 5.0 * N_AXIS * math.log2(N_AXIS)
    return flops / dt / 1e9, dt, cores, f"{sample_lanes} of 8192 lanes per transform (4 transforms/step), scipy.fft complex64, workers={cores}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, args.steps)
    gf, dt, cores, sample = cpu_port_run(steps, max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": max(1, min(args.warmup, 3)), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: 8192x8192 c64 ndfft axis1, ndfft axis0, ndifft axis0, ndifft axis1 (bounded lane sample)"},
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "ndrustfft itself is not buildable in this image (no rustc/cargo, crates not vendored); oracle port = scipy.fft over lanes"},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# per-launch DRAM traffic of the two bench kernels from `ncu --set full` captures of this same command (profiles/)
ROWS_TRAFFIC = 1.028e9
ROWS_TRAFFIC_SRC = "ncu --set full, profiles/r1z_ncu_bench_summary.txt: dram__bytes_read 537 MB + dram__bytes_write 491 MB per launch"
COLS2_TRAFFIC = 2.05e9
COLS2_TRAFFIC_SRC = "ncu, profiles/r1l_ncu_bench_summary.txt + r1t l2 probe: 537 MB read + 488-492 MB written per pass kernel, two pass kernels per call"
COLS_TRAFFIC = 1.130e9
COLS_TRAFFIC_SRC = "ncu --set full, profiles/r1z_ncu_bench_summary.txt: dram__bytes_read 537 MB + dram__bytes_write 593 MB per launch (both passes; the intermediate stays in L2)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import ndrustfft_b200 as nb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    lib = nb._default_backend().lib
    h = nb.FftHandler(N_AXIS, np.float32, device=local)
    g = torch.Generator(device=dev); g.manual_seed(0xB200 + 32 + rank)
    x = torch.complex(torch.rand((N_AXIS, N_AXIS), generator=g, device=dev) * 2 - 1,
                      torch.rand((N_AXIS, N_AXIS), generator=g, device=dev) * 2 - 1)
    a = torch.empty_like(x)
    b = torch.empty_like(x)

    def step(evs=None):
        if evs: evs[0].record()
        nb.ndfft(x, a, h, 1)
        if evs: evs[1].record()
        nb.ndfft(a, b, h, 0)
        if evs: evs[2].record()
        nb.ndifft(b, a, h, 0)
        if evs: evs[3].record()
        nb.ndifft(a, b, h, 1)
        if evs: evs[4].record()

    for _ in range(W):
        step()
    torch.cuda.synchronize()
    lc0 = lib.launch_count()
    nb.ndfft(a, b, h, 0)
    strided_launches = lib.launch_count() - lc0          # kernels one strided-axis call launches
    nb.ndifft(b, a, h, 0)
    torch.cuda.synchronize()
    step()
    torch.cuda.synchronize()
    # correctness guard on the bench's own data: the four transforms are a round trip
    rel = (torch.linalg.vector_norm(b - x) / torch.linalg.vector_norm(x)).item()
    assert rel < 1e-5, f"round trip rel L2 {rel}"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
    launches0 = lib.launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(K):
        step(events[k])
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    per = [sum(events[k][i].elapsed_time(events[k][i + 1]) for k in range(K)) / K for i in range(4)]
    tmax = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = tmax.item()
    ms_per_step = total_ms_max / K
    value = world * 4 * FLOPS_PER_TRANSFORM / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host arrays through the C ABI's NDFB_MEM_HOST path (pinned staging) ----
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((N_AXIS, N_AXIS), dtype=torch.complex64).pin_memory()
        ha = torch.empty_like(hx).pin_memory()
        hb = torch.empty_like(hx).pin_memory()
        hx.copy_(x)
        nx, na, nbuf = hx.numpy(), ha.numpy(), hb.numpy()

        call_s = [0.0, 0.0, 0.0, 0.0]

        def host_step():
            t = [time.perf_counter()]
            nb.ndfft(nx, na, h, 1); t.append(time.perf_counter())
            nb.ndfft(na, nbuf, h, 0); t.append(time.perf_counter())
            nb.ndifft(nbuf, na, h, 0); t.append(time.perf_counter())
            nb.ndifft(na, nbuf, h, 1); t.append(time.perf_counter())
            for i in range(4):
                call_s[i] += t[i + 1] - t[i]

        host_step()
        barrier()
        call_s[:] = [0.0, 0.0, 0.0, 0.0]
        KE = max(1, min(args.e2e_steps, K))
        t0 = time.perf_counter()
        for _ in range(KE):
            host_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        et = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        dt = et.item() / KE
        relh = float(np.linalg.norm(nbuf - nx) / np.linalg.norm(nx))
        assert relh < 1e-5, relh
        # the same four transforms as two multi-axis calls (ndfb_exec_chain: fft2 then ifft2): two PCIe round trips, not four
        def chain_step():
            nb.fft2(nx, na, h, h)
            nb.ifft2(na, nbuf, h, h)

        chain_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(KE):
            chain_step()
        torch.cuda.synchronize()
        ct = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        cdt_s = ct.item() / KE
        relc = float(np.linalg.norm(nbuf - nx) / np.linalg.norm(nx))
        assert relc < 1e-5, relc
        e2e = {"value": world * 4 * FLOPS_PER_TRANSFORM / dt / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": 4 * N_AXIS * N_AXIS * 8, "d2h_bytes_per_step": 4 * N_AXIS * N_AXIS * 8,
               "ms_per_step": dt * 1e3, "steps": KE, "ms_per_call": [round(c / KE * 1e3, 3) for c in call_s],
               "path": "ndfb_exec(mem=HOST) on pinned numpy arrays, 4 calls/step; each call pipelines H2D | kernel | D2H in pieces",
               "chained": {"value": world * 4 * FLOPS_PER_TRANSFORM / cdt_s / 1e9, "unit": "GFLOP/s", "ms_per_step": cdt_s * 1e3,
                           "h2d_bytes_per_step": 2 * N_AXIS * N_AXIS * 8, "d2h_bytes_per_step": 2 * N_AXIS * N_AXIS * 8,
                           "path": "same four transforms as two ndfb_exec_chain calls (fft2, ifft2): intermediates stay on the GPU"}}
        del hx, ha, hb

    if rank == 0:
        peak, peak_src = measured_peak()
        # The two contiguous-axis calls are ONE launch each of the 8192-point row kernel (2 launches/step, 1 GiB of
        # algorithmic bytes per launch); each strided-axis call is one persistent launch that runs both column passes
        # (64- and 128-point) with the intermediate kept in L2.  Shares are in `launches`.
        rows_ms = 0.5 * (per[0] + per[3])
        cols_ms = 0.5 * (per[1] + per[2])
        rows = {"kernel": "sfft_kernel<float, Sched<8192,512,16,16,16,2>, rows> (ndfft/ndifft along the contiguous axis: one launch per call)",
                "ms": rows_ms, "achieved": BYTES_PER_TRANSFORM / (rows_ms * 1e-3) / 1e9,
                "frac": BYTES_PER_TRANSFORM / (rows_ms * 1e-3) / 1e9 / peak, "share_of_step": (per[0] + per[3]) / sum(per),
                "traffic": ROWS_TRAFFIC, "traffic_source": ROWS_TRAFFIC_SRC}
        cols = {"kernel": ("fs2_kernel<float, Sched<64,...>, 64, Sched<128,...>, 32> (NDFB_FS2=1: both column passes of 8192 = 64 x 128 in one "
                           "persistent launch, workspace ring in L2)") if strided_launches == 1 else
                          "two sfft_kernel launches per call: 64-point then 128-point column passes through an HBM workspace (2 x 1 GiB moved)",
                "launches_per_call": strided_launches, "ms": cols_ms, "achieved": BYTES_PER_TRANSFORM / (cols_ms * 1e-3) / 1e9,
                "frac": BYTES_PER_TRANSFORM / (cols_ms * 1e-3) / 1e9 / peak, "share_of_step": (per[1] + per[2]) / sum(per),
                "traffic": COLS_TRAFFIC if strided_launches == 1 else COLS2_TRAFFIC,
                "traffic_source": COLS_TRAFFIC_SRC if strided_launches == 1 else COLS2_TRAFFIC_SRC}
        if strided_launches == 2:
            # each pass kernel moves the whole array once in and once out: its own roofline fraction
            cols["frac_of_bytes_moved_by_the_two_passes"] = 2 * BYTES_PER_TRANSFORM / (cols_ms * 1e-3) / 1e9 / peak
        # dominant kernel = the single kernel with the largest share of the step
        dom, other, other_key = (cols, rows, "contiguous_axis_kernel") if (strided_launches == 1 and cols["share_of_step"] >= rows["share_of_step"]) \
            else (rows, cols, "strided_axis_call")
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c2: 8192x8192 c64 ndfft axis1, ndfft axis0, ndifft axis0, ndifft axis1 (one step = 4 axis transforms)",
                       "l2": "inputs larger than L2 (512 MiB per array, 3 arrays cycled)", "sharding": "independent array per rank, no collective"},
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                         "frac": dom["frac"], "traffic": dom["traffic"], "traffic_source": dom["traffic_source"],
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_TRANSFORM,
                         "share_of_step": dom["share_of_step"], other_key: other},
            "launches": [{"name": STEP_NAMES[i], "ms": per[i], "GB/s": BYTES_PER_TRANSFORM / (per[i] * 1e-3) / 1e9,
                          "frac": BYTES_PER_TRANSFORM / (per[i] * 1e-3) / 1e9 / peak,
                          "GFLOP/s": FLOPS_PER_TRANSFORM / (per[i] * 1e-3) / 1e9} for i in range(4)],
            "gpu_launches": launches, "clocks": clocks, "e2e": e2e,
            "roundtrip_rel_l2": rel, "library": lib.version(),
        }
        if world == 1 and not args.no_cpu:
            gf, dt, cores, sample = cpu_port_run(2, 1)
            line["cpu_baseline"] = {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
