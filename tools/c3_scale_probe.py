#!/usr/bin/env python3
"""c3 slab pipeline at N ranks: serialised phase times and whole-pipeline variants (i2-chunks x CUDA graph) in ONE launch.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c3_scale_probe.py [--variants 1,2,4:0,1]
One JSON line per measurement from rank 0 (device time, max over ranks)."""
import argparse, json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", default="1,2,4,8")
ap.add_argument("--graph", default="0,1")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--phases", type=int, default=1)
ap.add_argument("--blocked", default="0")
ap.add_argument("--overlap", default="1,0")
ap.add_argument("--ctas", default="1")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from ndrustfft_b200.dist import SlabR2cFft3d
n = 512; m = n // 2 + 1
FLOPS = n * n * 2.5 * n * math.log2(n) + 2 * n * m * 5.0 * n * math.log2(n)


def allmax(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def timed(fn, steps, per=1):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return allmax(e0.elapsed_time(e1) / steps / per)


def emit(d):
    if rank == 0:
        print(json.dumps(d), flush=True)


g = torch.Generator(device=dev); g.manual_seed(0xB200 + 48 + rank)
x = torch.rand((n // world, n, n), generator=g, device=dev, dtype=torch.float64) * 2 - 1

if a.phases and world > 1:
    plan = SlabR2cFft3d((n, n, n), np.float64, device=dev, chunks=1, overlap=False)
    be = plan.be
    for _ in range(3):
        plan.forward(x)
    s0, s1, mp = plan.s0, plan.s1, plan.mp
    buf, hdl = plan._symm[0]
    chunk = s0 * s1 * mp * 16
    ptrs = [int(hdl.buffer_ptrs[p]) + rank * chunk for p in range(world)]
    recv = buf.view(n, s1, mp)
    ph = {"world": world, "what": "serialised phases (each phase alone, barrier between repetitions)"}
    ph["r2c_ms"] = timed(lambda: be.ndfft_r2c(x, plan.a, plan.h2, 2), 10)
    ph["axis1_local_ms"] = timed(lambda: be.ndfft(plan.a_pad, plan.b_pad, plan.h1, 1), 10)
    ph["axis1_scatter_ms"] = timed(lambda: be.ndfft_scatter_out(plan.a_pad, plan.h1, 1, out_shape=(s0, n, mp), out_strides=(s1 * mp, mp, 1), out_block=s1, block_ptrs=ptrs), 10)
    lb_ = plan.lanes128; nb_ = mp // lb_
    ph["axis1_scatter_blocked_ms"] = timed(lambda: be.ndfft_scatter_out(plan.a_pad.view(s0, n, nb_, lb_), plan.h1, 1, out_shape=(s0, n, nb_, lb_), out_strides=(s1 * mp, lb_, s1 * lb_, 1), out_block=s1, block_ptrs=ptrs), 10)
    ph["axis0_blocked_ms"] = timed(lambda: be.ndfft(buf.view(n, nb_, s1, lb_).permute(0, 2, 1, 3), plan.out_pad.view(n, s1, nb_, lb_), plan.h0, 0), 10)
    ph["barrier_ms"] = timed(lambda: hdl.barrier(), 10)
    ph["axis0_ms"] = timed(lambda: be.ndfft(recv, plan.out_pad, plan.h0, 0), 10)
    sent = plan.bytes_sent_per_rank()
    ph["bytes_sent_per_rank"] = sent
    ph["scatter_GBps_out"] = sent / (ph["axis1_scatter_ms"] * 1e-3) / 1e9
    emit(ph)
    del plan

for chunks, blocked, overlap, ctas in [(int(v), int(b), int(o), int(c)) for v in a.chunks.split(",") for b in a.blocked.split(",") for o in a.overlap.split(",") for c in a.ctas.split(",")]:
    if (world == 1 and chunks > 1) or (chunks > 1 and (blocked or overlap)) or (not overlap and ctas != int(a.ctas.split(",")[0])):
        continue
    plan = SlabR2cFft3d((n, n, n), np.float64, device=dev, chunks=chunks, blocked=bool(blocked), overlap=bool(overlap))
    plan.consumer_ctas = ctas
    for _ in range(3):
        out = plan.forward(x)
    back = plan.inverse(out)
    rel = (torch.linalg.vector_norm(back - x) / torch.linalg.vector_norm(x)).item()
    del back
    for graph in [int(v) for v in a.graph.split(",")]:
        run, per = (lambda: plan.forward(x)), 1
        if graph:
            try:
                if getattr(plan, "peer", False) and plan._call % 2:
                    plan.forward(x)
                gr = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    with torch.cuda.graph(gr, stream=side):
                        plan.forward(x); plan.forward(x)
                torch.cuda.current_stream(dev).wait_stream(side)
                run, per = (lambda: gr.replay()), 2
            except Exception as e:
                emit({"chunks": chunks, "graph": 1, "error": repr(e)[:300]})
                continue
        ms = timed(run, a.steps, per)
        emit({"cfg": "c3", "n_gpus": world, "chunks": chunks, "blocked": bool(blocked), "overlap": bool(overlap), "consumer_ctas": ctas, "cuda_graph": bool(graph), "ms": ms, "GFLOP/s": FLOPS / (ms * 1e-3) / 1e9,
              "peer": bool(getattr(plan, "peer", False)), "roundtrip_rel_l2": rel,
              "nvlink_ms_at_770": plan.bytes_sent_per_rank() / 770e9 * 1e3 if world > 1 else 0.0})
    del plan
if world > 1:
    dist.destroy_process_group()
