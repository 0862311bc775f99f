#!/bin/bash
# round 2, first GPU pass: parity suite, smoke, the new bench line (all configs), launch list of the headline step
mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; lscpu | head -25 >> gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt; nvidia-smi topo -m >> gpurun_out/r2a_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
NDFB_NO_STAGING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu > gpurun_out/r2a_bench_nostaging.json 2>> gpurun_out/r2a_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.json 2>> gpurun_out/r2a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu --no-e2e > /dev/null 2>> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_pytest.log; tail -2 gpurun_out/r2a_smoke.log; head -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
