#!/bin/bash
# 4 GPUs: i2-chunked overlap at N = 4 and N = 2 (bigger kernels than at N = 8)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29951 tools/c3_scale_probe.py --chunks 1,2,3,4 --graph 1 --blocked 0 --overlap 0 --phases 0 2> gpurun_out/r2h.err | grep -E '^\{' > gpurun_out/r2h_probe_n4.jsonl
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29952 tools/c3_scale_probe.py --chunks 1,2,3,4 --graph 1 --blocked 0 --overlap 0 --phases 0 2>> gpurun_out/r2h.err | grep -E '^\{' > gpurun_out/r2h_probe_n2.jsonl
cat gpurun_out/r2h_probe_n4.jsonl gpurun_out/r2h_probe_n2.jsonl | cut -c1-200
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29953 bench.py --gpus 2 --steps 20 --warmup 5 2>> gpurun_out/r2h.err | grep -E '^\{' > gpurun_out/r2h_bench_n2.json
python -c "
import json; b=json.load(open('gpurun_out/r2h_bench_n2.json')); print('bench n2', b['value'], b['ms_per_step'], b['e2e'])"
grep -iE "error|Traceback" gpurun_out/r2h.err | head -5
