#!/bin/bash
# mirror-paired DCT-III / DCT-IV output pass: parity + same-box A/B; what HBM gives tile-wise copies of narrow strided rows (probe)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "mirror_paired_output or pipelined or config4" > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
timeout 600 python tools/ab_env.py > gpurun_out/r2s_ab_mirror_out.jsonl 2> gpurun_out/r2s.err; cat gpurun_out/r2s_ab_mirror_out.jsonl
timeout 300 ./tools/probes/strided_copy_probe > gpurun_out/r2s_strided_copy_probe.jsonl 2>> gpurun_out/r2s.err; cat gpurun_out/r2s_strided_copy_probe.jsonl
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --only c4,c5a --no-cpu --no-e2e > gpurun_out/r2s_bench_c4_c5a.json 2>> gpurun_out/r2s.err
python - <<'PY'
import json
b = json.load(open('gpurun_out/r2s_bench_c4_c5a.json'))
for r in b.get('configs', []):
    print(r['cfg'], r['call'], r.get('ms'), r.get('frac'))
PY
tail -3 gpurun_out/r2s.err
