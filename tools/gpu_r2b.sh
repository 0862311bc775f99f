#!/bin/bash
# round 2: c3 slab pipeline at N GPUs (N = number of visible GPUs): phases, chunk x graph variants, the bench line itself
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/r2b_topo_n$N.txt 2>&1
timeout 120 python bench.py --no-configs --no-cpu --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1_on$N.json 2> gpurun_out/r2b.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29911 tools/c3_scale_probe.py --chunks 1,2,4,8 --graph 0,1 2>> gpurun_out/r2b.err | grep -E '^\{' > gpurun_out/r2b_c3_probe_n$N.jsonl
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29912 bench.py --gpus $N --steps 20 --warmup 5 2>> gpurun_out/r2b.err | grep -E '^\{' > gpurun_out/r2b_bench_n$N.json
cat gpurun_out/r2b_c3_probe_n$N.jsonl | cut -c1-400; head -c 600 gpurun_out/r2b_bench_n$N.json; grep -iE "error|Traceback" gpurun_out/r2b.err | head -5
