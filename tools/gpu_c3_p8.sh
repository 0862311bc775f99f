#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/c3_p8_variants.jsonl
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 tools/bench_c3.py --peer on ${@:2} 2>gpurun_out/c3p8.err | grep cfg >> gpurun_out/c3_p8_variants.jsonl; grep -iE "error" gpurun_out/c3p8.err | head -2; }
run 29901 --row-chunks 1
run 29902 --row-chunks 1 --graph
run 29903 --row-chunks 2 --scatter-smem 118784
run 29904 --row-chunks 4 --scatter-smem 118784
run 29905 --row-chunks 4 --scatter-smem 0
run 29906 --row-chunks 4 --scatter-smem 118784 --graph
python - <<'PY'
import json
for l in open('gpurun_out/c3_p8_variants.jsonl'):
    d=json.loads(l); print(round(d['ms'],4), 'rows',d['row_chunks'],'smem',d['scatter_smem'],'graph',d['cuda_graph'], 'rel', d['roundtrip_rel_l2'])
PY
