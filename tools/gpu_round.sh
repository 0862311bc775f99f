#!/bin/bash
rm -f gpurun_out/*.ncu-rep
# One GPU session: smoke, parity tests, bench, config sweep, ncu launch list + full capture of the bench kernels.
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
if [ -z "$SKIP_TESTS" ]; then
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 900 python tools/bench_configs.py ${SWEEP_ARGS} > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fs2_kernel|sfft_kernel|tile_kernel' -s 4 -c 4 -o /tmp/prof_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
# summarise on the box and drop the report: gpurun copies back at most 64 MiB
python tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_bench_summary.txt 2>&1
python tools/ncu_opmix.py /tmp/prof_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_bench_opmix.txt 2>&1
if [ -n "$EXTRA_NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rsfft_kernel|sfft_kernel' -c 14 -o /tmp/prof_${TAG}_sweep python tools/bench_configs.py --only c3,c4 --iters 1 > gpurun_out/ncu_sweep.log 2>&1
python tools/ncu_summary.py /tmp/prof_${TAG}_sweep.ncu-rep > gpurun_out/${TAG}_ncu_sweep_summary.txt 2>&1
fi
tail -3 gpurun_out/smoke.log; tail -12 gpurun_out/pytest_gpu.log 2>/dev/null; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench.err; cat gpurun_out/configs_$TAG.jsonl; tail -5 gpurun_out/configs.err
