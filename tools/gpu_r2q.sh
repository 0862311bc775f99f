#!/bin/bash
# pipelined column kernel: GPU parity test, same-box A/B, ncu full capture of both c5b passes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pipelined or config5b" > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
timeout 900 python tools/ab_pipe.py > gpurun_out/r2q_ab_pipe.jsonl 2> gpurun_out/r2q.err; cat gpurun_out/r2q_ab_pipe.jsonl
SHAPE=16x16777216 AXIS=1 F64=0 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sfft_pipe_kernel' -s 2 -c 2 -o /tmp/prof_r2q python tools/run_one.py > gpurun_out/r2q_ncu.log 2>&1
python tools/ncu_summary.py /tmp/prof_r2q.ncu-rep > gpurun_out/r2q_ncu_c5b_pipe_summary.txt 2>&1
grep -E "Kernel Name|time_duration|dram__bytes|dram_throughput|issue_active|stalled|warps_active|bank_conflicts|wavefronts|registers|l1tex__throughput|lts__throughput" gpurun_out/r2q_ncu_c5b_pipe_summary.txt
NDFB_PIPE=0 SHAPE=16x16777216 AXIS=1 F64=0 ITERS=2 timeout 600 ncu --set full --clock-control none -k regex:'sfft_kernel' -s 2 -c 2 -o /tmp/prof_r2q0 python tools/run_one.py >> gpurun_out/r2q_ncu.log 2>&1
python tools/ncu_summary.py /tmp/prof_r2q0.ncu-rep > gpurun_out/r2q_ncu_c5b_nopipe_summary.txt 2>&1
grep -E "Kernel Name|time_duration|dram_throughput|issue_active|stalled" gpurun_out/r2q_ncu_c5b_nopipe_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rsfft_kernel' -c 8 -o /tmp/prof_r2q_c4 python tools/bench_configs.py --only c4 --iters 1 > gpurun_out/r2q_ncu_c4.log 2>&1
python tools/ncu_summary.py /tmp/prof_r2q_c4.ncu-rep > gpurun_out/r2q_ncu_c4_summary.txt 2>&1
python tools/ncu_opmix.py /tmp/prof_r2q_c4.ncu-rep > gpurun_out/r2q_ncu_c4_opmix.txt 2>&1
grep -E "Kernel Name|time_duration|l1tex__throughput|issue_active|bank_conflicts|wavefronts_mem_shared" gpurun_out/r2q_ncu_c4_summary.txt | head -60
