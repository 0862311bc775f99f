#!/usr/bin/env python3
"""Phase-by-phase (serialised) timing of the slab pipeline: where do the milliseconds go at N ranks?"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
from ndrustfft_b200.dist import SlabR2cFft3d
n = 512
for chunks in (1, 2, 4, 8):
    plan = SlabR2cFft3d((n, n, n), np.float64, device=dev, chunks=chunks)
    be = plan.be
    x = torch.rand((n // world, n, n), device=dev, dtype=torch.float64)
    out = torch.empty((n, n // world, n // 2 + 1), dtype=torch.complex128, device=dev)
    for _ in range(3): plan.forward(x, out)
    torch.cuda.synchronize()
    def T(fn, reps=5):
        fn(); torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3
    res = {"chunks": chunks, "world": world}
    res["forward"] = T(lambda: plan.forward(x, out))
    res["r2c"] = T(lambda: be.ndfft_r2c(x, plan.a, plan.h2, 2))
    if world > 1:
        s0, s1, n0, n1 = plan.s0, plan.s1, plan.n0, plan.n1
        def fft1_all():
            for c, (lo, hi) in enumerate(plan.chunks):
                mc = hi - lo
                be.ndfft_split_out(plan.a[:, :, lo:hi], plan.send[c], plan.h1, 1, out_shape=(s0, n1, mc), out_strides=(s1 * mc, mc, 1), out_block=s1, out_block_stride=s0 * s1 * mc)
        def a2a_all():
            ws = [plan._a2a(plan.recv[c], plan.send[c]) for c in range(len(plan.chunks))]
            for w in ws: w.wait()
        def fft0_all():
            for c, (lo, hi) in enumerate(plan.chunks):
                be.ndfft(plan.recv[c].view(n0, s1, hi - lo), out[:, :, lo:hi], plan.h0, 0)
        res["fft1_split"] = T(fft1_all); res["a2a"] = T(a2a_all); res["fft0"] = T(fft0_all)
        res["fft1_plain"] = T(lambda: be.ndfft(plan.a, plan.b, plan.h1, 1))
    if rank == 0: print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
    del plan
if world > 1: dist.destroy_process_group()
