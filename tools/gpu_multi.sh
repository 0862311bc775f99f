#!/bin/bash
# Multi-GPU session: NCCL slab test, c3 strong scaling and c2 weak scaling at 1..NG GPUs.
NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_dist.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dist.log
for N in 1 2 4 8; do
 [ $N -gt $NG ] && break
 if [ $N = 1 ]; then
   timeout 300 python tools/bench_c3.py > gpurun_out/c3_n$N.json 2> gpurun_out/c3_n$N.err
   timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
 else
   timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2970$N tools/bench_c3.py > gpurun_out/c3_n$N.json 2> gpurun_out/c3_n$N.err
   timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
 fi
done
tail -3 gpurun_out/pytest_dist.log
for N in 1 2 4 8; do [ -f gpurun_out/c3_n$N.json ] && grep cfg gpurun_out/c3_n$N.json; done
for N in 1 2 4 8; do [ -f gpurun_out/bench_n$N.json ] && grep metric gpurun_out/bench_n$N.json | cut -c1-330; done
tail -2 gpurun_out/c3_n$NG.err
