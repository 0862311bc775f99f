#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_fs_medium.py > gpurun_out/r2x_fs_medium.txt 2> gpurun_out/r2x.err; cat gpurun_out/r2x_fs_medium.txt; tail -3 gpurun_out/r2x.err
