mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu2.txt 2>&1
timeout 600 python -m pytest tests/test_dist.py tests/test_parity_gpu.py -m gpu -q --maxfail=10 --timeout 600 -p no:cacheprovider -k "nccl or four_step or config2 or config5b" > gpurun_out/pytest_gpu2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu2.log
for N in 1 2; do
 if [ $N = 1 ]; then timeout 300 python tools/bench_c3.py > gpurun_out/c3_n$N.json 2> gpurun_out/c3_n$N.err
 else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 tools/bench_c3.py > gpurun_out/c3_n$N.json 2> gpurun_out/c3_n$N.err; fi
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
timeout 600 python tools/bench_configs.py --only c2,c5b > gpurun_out/configs_r1e.jsonl 2>&1
tail -5 gpurun_out/pytest_gpu2.log; cat gpurun_out/c3_n1.json gpurun_out/c3_n2.json; tail -3 gpurun_out/c3_n2.err; cat gpurun_out/bench_n2.json | cut -c1-400; tail -2 gpurun_out/bench_n2.err; cat gpurun_out/configs_r1e.jsonl
