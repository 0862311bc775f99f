#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29921 tools/c3_scale_probe.py --chunks 1 --graph 0,1 --blocked 1,0 2> gpurun_out/r2d.err | grep -E '^\{' > gpurun_out/r2d_c3_probe_n$N.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29922 bench.py --gpus $N --steps 20 --warmup 5 2>> gpurun_out/r2d.err | grep -E '^\{' > gpurun_out/r2d_bench_n$N.json
if [ "$N" -ge 8 ]; then
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29923 tools/c3_scale_probe.py --chunks 1 --graph 1 --blocked 1,0 --phases 1 2>> gpurun_out/r2d.err | grep -E '^\{' > gpurun_out/r2d_c3_probe_n4.jsonl
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29924 tools/c3_scale_probe.py --chunks 1 --graph 1 --blocked 1,0 --phases 1 2>> gpurun_out/r2d.err | grep -E '^\{' > gpurun_out/r2d_c3_probe_n2.jsonl
fi
cat gpurun_out/r2d_c3_probe_n*.jsonl | cut -c1-420; head -c 400 gpurun_out/r2d_bench_n$N.json; grep -iE "error|Traceback" gpurun_out/r2d.err | head -5
