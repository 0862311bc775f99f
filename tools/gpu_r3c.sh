#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "four_step or config5b or three_pass or fused_two_pass or staged or in_place or long_bluestein or jit_lengths_c2c" > gpurun_out/r3c_pytest.log 2>&1; tail -3 gpurun_out/r3c_pytest.log
timeout 900 python tools/exp_fs_medium.py > gpurun_out/r3c_fs_medium.txt 2> gpurun_out/r3c.err; grep -E "^\{" gpurun_out/r3c_fs_medium.txt | cut -c1-200; tail -3 gpurun_out/r3c.err
