#!/bin/bash
# driver-style verification of the final state: GPU parity suite, smoke, bench line (N = 1), reference arm, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3h_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3h_bench_n1.json 2> gpurun_out/r3h.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3h_reference_n1.json 2>> gpurun_out/r3h.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r3h_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu --no-e2e > /dev/null 2>> gpurun_out/r3h.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r3h_bench_n1.json')); r=json.load(open('gpurun_out/r3h_reference_n1.json'))
print('ours', b['value'], b['ms_per_step'], 'e2e', b['e2e']['value'], b['e2e']['ms_per_step'], 'roofline', b['roofline']['frac'], b['roofline']['step_frac'])
print('ref ', r['value'], r['ms_per_step'], 'same config', b['config']==r['config'], 'ratio', b['value']/r['value'], 'e2e ratio', b['e2e']['value']/r['value'])
PY
tail -3 gpurun_out/r3h.err
python - <<'PY'
import json
b = json.load(open('gpurun_out/r3h_bench_n1.json'))
for r in b.get('configs', []):
    print(r['cfg'], r['call'], r.get('ms'), r.get('frac'))
PY
