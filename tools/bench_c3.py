#!/usr/bin/env python3
"""BASELINE config c3: 3-D f64 real transform 512^3 (ndfft_r2c last axis, ndfft axes 1 and 0), slab-decomposed over
the ranks of one node.  Launch: python tools/bench_c3.py            (1 GPU)
        python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c3.py
Strong scaling: the global array is fixed, each rank owns n0/N slabs.  Prints one JSON line from rank 0 with the
per-phase device times (max over ranks), GFLOP/s, fraction of the HBM roofline and, for N > 1, of the NVLink roofline."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--peer", default="auto", help="auto | on | off: fused peer-store all-to-all vs NCCL all_to_all_single")
    ap.add_argument("--chunks", type=int, default=1)
    ap.add_argument("--row-chunks", type=int, default=1)
    ap.add_argument("--scatter-smem", type=int, default=-1, help="shared-memory floor (bytes) of the scatter launch; -1 = auto")
    ap.add_argument("--graph", action="store_true", help="replay the forward transform from a CUDA graph (removes host launch overhead)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ndrustfft_b200.dist import SlabR2cFft3d
    n = a.n
    plan = SlabR2cFft3d((n, n, n), np.float64, device=dev, chunks=a.chunks, peer={'auto': 'auto', 'on': True, 'off': False}[a.peer],
                        row_chunks=a.row_chunks, scatter_smem=None if a.scatter_smem < 0 else a.scatter_smem)
    g = torch.Generator(device=dev); g.manual_seed(0xB200 + 48 + rank)
    x = torch.rand((n // world, n, n), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    out = torch.empty((n, n // world, n // 2 + 1), dtype=torch.complex128, device=dev)
    for _ in range(a.warmup):
        plan.forward(x, out)
    back = plan.inverse(out)
    rel = (torch.linalg.vector_norm(back - x) / torch.linalg.vector_norm(x)).item()
    assert rel < 1e-12, rel
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    run = lambda: plan.forward(x, out)
    if a.graph:
        if plan.peer and plan._call % 2:
            plan.forward(x, out)                       # keep the double-buffer parity of capture and replay aligned
        gr = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                plan.forward(x, out)
                plan.forward(x, out)                   # two calls per replay: both receive buffers
        torch.cuda.current_stream(dev).wait_stream(side)
        run = lambda: gr.replay()
        run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps / (2 if a.graph else 1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    m = n // 2 + 1
    flops = n * n * 2.5 * n * math.log2(n) + 2 * n * m * 5.0 * n * math.log2(n)
    nbytes = (n ** 3 * 8 + n * n * m * 16) + 2 * (2 * n * n * m * 16)     # the three axis passes, input once + output once each
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    if rank == 0:
        line = {"cfg": "c3", "call": f"rfft3d {n}^3 f64 slab x{world}", "n_gpus": world, "ms": ms, "GFLOP/s": flops / (ms * 1e-3) / 1e9,
                "hbm_frac_per_gpu": nbytes / world / (ms * 1e-3) / 1e9 / peak, "roundtrip_rel_l2": rel, "cuda_graph": bool(a.graph), "chunks": a.chunks, "row_chunks": a.row_chunks, "scatter_smem": getattr(plan, "scatter_smem", None), "exchange": ("peer stores fused into the axis-1 kernel" if plan.peer else ("NCCL all_to_all_single" if world > 1 else "none")),
                "a2a_bytes_sent_per_rank": plan.bytes_sent_per_rank(),
                "nvlink_time_at_770GBs_ms": plan.bytes_sent_per_rank() / 770e9 * 1e3 if world > 1 else 0.0}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
