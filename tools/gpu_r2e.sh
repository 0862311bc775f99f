#!/bin/bash
# 1 GPU: full parity suite (JIT included), JIT column sweep after the single-pass fix, ncu full capture of the c3 kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
timeout 900 python tools/sweep_jit.py 96,192,480,720,768,1080,1200,1296,1536,1920,2160,2400,3072,3600,4000,4800,6000,6561,7200,8000 > gpurun_out/r2e_jit_sweep.jsonl 2> gpurun_out/r2e.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sfft_kernel|rsfft_kernel" -s 9 -c 3 -o gpurun_out/r2e_c3_prof python bench.py --steps 2 --warmup 3 --no-configs --no-cpu --no-e2e > /dev/null 2>> gpurun_out/r2e.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu > gpurun_out/r2e_bench_n1.json 2>> gpurun_out/r2e.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2e_jit_sweep.jsonl')]
for r in rows: print(r['n'], r['dtype'], r['layout'], r['frac'], r['rel_l2'])
PY
tail -3 gpurun_out/r2e.err
