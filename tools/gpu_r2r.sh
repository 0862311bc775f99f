#!/bin/bash
# pipelined column kernel after the lane-table / factored four-step twiddle changes: parity, A/B, ncu; odd-radix twiddle powers A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pipelined or config5b or four_step" > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
timeout 900 python tools/ab_pipe.py > gpurun_out/r2r_ab_pipe.jsonl 2> gpurun_out/r2r.err; grep -E "^\{" gpurun_out/r2r_ab_pipe.jsonl | cut -c1-260
timeout 600 python tools/ab_lib.py > gpurun_out/r2r_ab_tw_pow_odd.jsonl 2>> gpurun_out/r2r.err; cat gpurun_out/r2r_ab_tw_pow_odd.jsonl
SHAPE=16x16777216 AXIS=1 F64=0 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sfft_pipe_kernel' -s 2 -c 2 -o /tmp/prof_r2r python tools/run_one.py > gpurun_out/r2r_ncu.log 2>&1
python tools/ncu_summary.py /tmp/prof_r2r.ncu-rep > gpurun_out/r2r_ncu_c5b_pipe_summary.txt 2>&1
grep -E "Kernel Name|time_duration|dram__bytes|dram_throughput|issue_active|inst_executed.sum|stalled|global_op_ld|l1tex__throughput" gpurun_out/r2r_ncu_c5b_pipe_summary.txt
