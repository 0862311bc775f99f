#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_c5b2.py > gpurun_out/r2i_c5b_variants.txt 2>&1; cat gpurun_out/r2i_c5b_variants.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.log
