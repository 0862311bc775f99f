#!/bin/bash
mkdir -p gpurun_out
NDFB_TRACE=1 timeout 600 python tools/ab_env.py nddct1 > gpurun_out/r3i_ab_dct1_schedules.jsonl 2> gpurun_out/r3i.err; cat gpurun_out/r3i_ab_dct1_schedules.jsonl; tail -2 gpurun_out/r3i.err
