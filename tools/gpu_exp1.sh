#!/bin/bash
# experiment: c5b factor splits; ncu full capture of the c4 DCT kernels
mkdir -p gpurun_out
for N1 in default 4096 8192 1024; do
  if [ "$N1" = default ]; then unset NDFB_FS_N1; else export NDFB_FS_N1=$N1; fi
  echo "N1=$N1" >> gpurun_out/exp1_c5b.txt
  NDFB_TRACE=1 timeout 300 python tools/bench_configs.py --only c5b --iters 5 >> gpurun_out/exp1_c5b.txt 2> gpurun_out/exp1_c5b_$N1.err
  grep -m3 "four-step\|sfft" gpurun_out/exp1_c5b_$N1.err >> gpurun_out/exp1_c5b.txt
done
unset NDFB_FS_N1
cat gpurun_out/exp1_c5b.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rsfft_kernel' -c 8 -o gpurun_out/prof_r1q_c4 python tools/bench_configs.py --only c4 --iters 1 > gpurun_out/ncu_c4.log 2>&1
tail -3 gpurun_out/ncu_c4.log
