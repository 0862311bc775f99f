#!/bin/bash
# 2-GPU sanity of the final code: NCCL dist test + the bench line at N = 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r3e_pytest_dist.log 2>&1; tail -2 gpurun_out/r3e_pytest_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29962 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r3e.err | grep -E '^\{' > gpurun_out/r3e_bench_n2.json
python -c "
import json; b=json.load(open('gpurun_out/r3e_bench_n2.json')); print(2, round(b['value'],1), round(b['ms_per_step'],4), b['scaling'], 'e2e', round(b['e2e']['value'],1) if b.get('e2e') else None, 'rt', b.get('roundtrip_rel_l2'))"
tail -3 gpurun_out/r3e.err
