#!/bin/bash
# ncu --set full of the final code: the three c5b passes (factored twiddle, transposing last pass) and the c4 row kernels
mkdir -p gpurun_out
SHAPE=16x16777216 AXIS=1 F64=0 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sfft_kernel' -s 3 -c 3 -o /tmp/prof_r3f python tools/run_one.py > gpurun_out/r3f_ncu.log 2>&1
python tools/ncu_summary.py /tmp/prof_r3f.ncu-rep > gpurun_out/r3f_ncu_c5b_summary.txt 2>&1
grep -E "Kernel Name|time_duration|dram__bytes|dram_throughput|issue_active|l1tex__throughput|global_op_ld.sum" gpurun_out/r3f_ncu_c5b_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rsfft_kernel' -c 16 -o /tmp/prof_r3f_c4 python tools/bench_configs.py --only c4 --iters 1 > gpurun_out/r3f_ncu_c4.log 2>&1
python tools/ncu_summary.py /tmp/prof_r3f_c4.ncu-rep > gpurun_out/r3f_ncu_c4_summary.txt 2>&1
grep -E "Kernel Name|time_duration|l1tex__throughput|bank_conflicts|wavefronts_mem_shared|dram__bytes" gpurun_out/r3f_ncu_c4_summary.txt | awk '!seen[$0]++' | head -80
