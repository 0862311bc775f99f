#!/bin/bash
# 8 GPUs: r2c || exchange overlap through per-plane counters (persistent consumer CTAs) vs the serial pipeline; then N = 4, 2
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29941 tools/c3_scale_probe.py --chunks 1 --graph 0,1 --blocked 0 --overlap 1,0 --ctas 1 --phases 0 2> gpurun_out/r2g.err | grep -E '^\{' > gpurun_out/r2g_probe_n$N.jsonl
cat gpurun_out/r2g_probe_n$N.jsonl | cut -c1-260
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29942 tools/c3_scale_probe.py --chunks 1 --graph 1 --blocked 0 --overlap 1 --ctas 2 --phases 0 2>> gpurun_out/r2g.err | grep -E '^\{' > gpurun_out/r2g_probe_ctas2_n$N.jsonl
cat gpurun_out/r2g_probe_ctas2_n$N.jsonl | cut -c1-260
if [ "$N" -ge 8 ]; then
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29943 tools/c3_scale_probe.py --chunks 1 --graph 1 --blocked 0 --overlap 1,0 --ctas 1 --phases 0 2>> gpurun_out/r2g.err | grep -E '^\{' > gpurun_out/r2g_probe_n4.jsonl
CUDA_VISIBLE_DEVICES=0,1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29944 tools/c3_scale_probe.py --chunks 1 --graph 1 --blocked 0 --overlap 1,0 --ctas 1 --phases 0 2>> gpurun_out/r2g.err | grep -E '^\{' > gpurun_out/r2g_probe_n2.jsonl
cat gpurun_out/r2g_probe_n4.jsonl gpurun_out/r2g_probe_n2.jsonl | cut -c1-260
fi
grep -iE "error|Traceback" gpurun_out/r2g.err | head -5
