#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "config4 or lengths_f64 or lengths_f32 or plan_families or mirror_paired" > gpurun_out/r3j_pytest.log 2>&1; tail -2 gpurun_out/r3j_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --only c4 --no-cpu --no-e2e > gpurun_out/r3j_bench_c4.json 2> gpurun_out/r3j.err
python - <<'PY'
import json
b = json.load(open('gpurun_out/r3j_bench_c4.json'))
print(b['ms_per_step'], b['roofline']['step_frac'])
for r in b.get('configs', []):
    print(r['cfg'], r['call'], r.get('ms'), r.get('frac'))
PY
tail -2 gpurun_out/r3j.err
