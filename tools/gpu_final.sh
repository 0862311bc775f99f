#!/bin/bash
# Final evidence run: bench line, config sweep, ncu launch list and full captures summarised ON the box (the .ncu-rep
# files are deleted afterwards: gpurun copies back at most 64 MiB).
TAG=${1:-final}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 900 python tools/bench_configs.py > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fs2_kernel|sfft_kernel' -s 4 -c 4 -o /tmp/prof_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_bench_summary.txt 2>&1
python tools/ncu_opmix.py /tmp/prof_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_bench_opmix.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rsfft_kernel|sfft_kernel' -c 16 -o /tmp/prof_${TAG}_sweep python tools/bench_configs.py --only c3,c4 --iters 1 > gpurun_out/ncu_sweep.log 2>&1
python tools/ncu_summary.py /tmp/prof_${TAG}_sweep.ncu-rep > gpurun_out/${TAG}_ncu_sweep_summary.txt 2>&1
rm -f /tmp/prof_$TAG*.ncu-rep
cat gpurun_out/bench_$TAG.json | cut -c1-400; wc -l gpurun_out/configs_$TAG.jsonl gpurun_out/launches_$TAG.csv gpurun_out/${TAG}_ncu_*.txt; du -sh gpurun_out
