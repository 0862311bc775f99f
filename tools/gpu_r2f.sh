#!/bin/bash
# 8 GPUs: bulk-async (TMA) scatter stores vs per-thread stores, the bench line at N = 8 / 4 / 2, NCCL test of the peer inverse
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29931 tools/c3_scale_probe.py --chunks 1 --graph 0,1 --blocked 1 2> gpurun_out/r2f.err | grep -E '^\{' > gpurun_out/r2f_probe_bulk_n$N.jsonl
NDFB_NO_BULK_STORE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29932 tools/c3_scale_probe.py --chunks 1 --graph 1 --blocked 1,0 --phases 0 2>> gpurun_out/r2f.err | grep -E '^\{' > gpurun_out/r2f_probe_nobulk_n$N.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29933 bench.py --gpus $N --steps 20 --warmup 5 2>> gpurun_out/r2f.err | grep -E '^\{' > gpurun_out/r2f_bench_n$N.json
if [ "$N" -ge 8 ]; then
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29934 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e 2>> gpurun_out/r2f.err | grep -E '^\{' > gpurun_out/r2f_bench_n4.json
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29935 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e 2>> gpurun_out/r2f.err | grep -E '^\{' > gpurun_out/r2f_bench_n2.json
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2f_pytest_dist.log 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 100 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu --no-e2e > gpurun_out/r2f_bench_n1.json 2>> gpurun_out/r2f.err
fi
cat gpurun_out/r2f_probe_*.jsonl | cut -c1-300; for n in 1 2 4 8; do python -c "
import json,sys; b=json.load(open('gpurun_out/r2f_bench_n$n.json')); print($n, b['value'], b['ms_per_step'])"; done; tail -3 gpurun_out/r2f_pytest_dist.log; grep -iE "error|Traceback" gpurun_out/r2f.err | head -5
