#!/bin/bash
# mirror-paired last pass (R2C / DCT-I / DCT-II): parity + the c3 / c4 / c1 config rows; full bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2l_bench_n1.json'))
print(b['value'], b['ms_per_step'], [(round(s['ms'],4), round(s['frac'],3)) for s in b['roofline']['stages']], 'e2e', b['e2e']['ms_per_step'], b['e2e']['pinned_arrays']['ms_per_step'])
for r in b['configs']:
    if r['cfg'] in ('c2','c3','c4','c5b') and r.get('frac') is not None: print(r['cfg'], r['call'][:44], r['ms'], r['frac'])
PY
tail -3 gpurun_out/r2l.err
