#!/bin/bash
# final multi-GPU check on one 8-GPU box: the bench line at N = 8, 4, 2, 1 (what the driver's scale run does) + NCCL dist test
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for n in 8 4 2; do
  if [ "$n" -le "$N" ]; then
    DEV=$(seq -s, 0 $((n-1)))
    CUDA_VISIBLE_DEVICES=$DEV timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29960+n)) bench.py --gpus $n --steps 20 --warmup 5 2>> gpurun_out/r2p.err | grep -E '^\{' > gpurun_out/r2p_bench_n$n.json
  fi
done
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --no-cpu > gpurun_out/r2p_bench_n1.json 2>> gpurun_out/r2p.err
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2p_pytest_dist.log 2>&1
CUDA_VISIBLE_DEVICES=0,1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29970 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>> gpurun_out/r2p.err | grep -E '^\{' > gpurun_out/r2p_reference_n2.json
for n in 1 2 4 8; do python -c "
import json; b=json.load(open('gpurun_out/r2p_bench_n$n.json')); print($n, round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1) if b.get('e2e') else None, (b.get('c2_weak') or {}).get('value'))"; done
tail -2 gpurun_out/r2p_pytest_dist.log; head -c 300 gpurun_out/r2p_reference_n2.json; echo; grep -iE "error|Traceback" gpurun_out/r2p.err | head -5
