#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --only c3,c4 --no-cpu --no-e2e > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m.err; echo "bench rc=$?"
NDFB_NO_MIRROR_PRO=1 timeout 900 python bench.py --steps 20 --warmup 5 --only c3,c4 --no-cpu --no-e2e > gpurun_out/r2m_bench_n1_nomirrorpro.json 2>> gpurun_out/r2m.err
python - <<'PY'
import json
for f in ('gpurun_out/r2m_bench_n1.json','gpurun_out/r2m_bench_n1_nomirrorpro.json'):
    b=json.load(open(f)); print(f, b['ms_per_step'])
    for r in b['configs']:
        if r.get('frac') is not None and ('dct3' in r['call'] or 'ifft' in r['call']): print('  ', r['cfg'], r['call'][:44], r['ms'], r['frac'])
PY
tail -3 gpurun_out/r2m.err
