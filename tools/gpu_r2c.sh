#!/bin/bash
# round 2: run-time compiled schedules on the GPU: parity tests + sweep over unlisted lengths
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "jit" > gpurun_out/r2c_pytest_jit.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c_pytest_jit.log
timeout 1200 python tools/sweep_jit.py > gpurun_out/r2c_jit_sweep.jsonl 2> gpurun_out/r2c_sweep.err; echo "sweep rc=$?"
NDFB_NO_JIT=1 timeout 600 python tools/sweep_jit.py 96,192,720,768,1200,1536,3072,6561 > gpurun_out/r2c_nojit_sweep.jsonl 2>> gpurun_out/r2c_sweep.err
tail -3 gpurun_out/r2c_sweep.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2c_jit_sweep.jsonl')]
print(len(rows), 'rows; min frac', min(r['frac'] for r in rows), 'max rel f64', max(r['rel_l2'] for r in rows if r['dtype']=='f64'), 'max rel f32', max(r['rel_l2'] for r in rows if r['dtype']=='f32'))
bad=[(r['n'],r['dtype'],r['layout'],r['frac']) for r in rows if r['frac']<0.5]
print('below 0.5:', len(bad), bad[:40])
PY
