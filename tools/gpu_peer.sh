mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dist.log
tail -4 gpurun_out/pytest_dist.log
for MODE in off on; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29821 tools/bench_c3.py --peer $MODE 2> gpurun_out/c3_peer_$MODE.err | grep cfg | tee gpurun_out/c3_n2_peer_$MODE.json
done
tail -3 gpurun_out/c3_peer_on.err
