#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "mirror_paired_output or config4 or jit_lengths_real" > gpurun_out/r3g_pytest.log 2>&1; tail -2 gpurun_out/r3g_pytest.log
timeout 600 python tools/ab_env.py > gpurun_out/r3g_ab_mirror_out_wide.jsonl 2> gpurun_out/r3g.err; cat gpurun_out/r3g_ab_mirror_out_wide.jsonl; tail -2 gpurun_out/r3g.err
