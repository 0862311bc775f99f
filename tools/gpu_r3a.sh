#!/bin/bash
# medium multi-pass rows: which pass is slow?  (ncu launch lists), plus the pipelined-kernel GPU test after its fix
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pipelined or three_pass" > gpurun_out/r3a_pytest.log 2>&1; tail -2 gpurun_out/r3a_pytest.log
for cfg in "8192x65536:default" "8192x65536:trans" "2048x262144:default" "2048x262144:trans" "512x1048576:default"; do
  shape=${cfg%%:*}; v=${cfg##*:}
  if [ "$v" = trans ]; then export NDFB_FS_TRANSPOSE=1; else unset NDFB_FS_TRANSPOSE; fi
  SHAPE=$shape AXIS=1 F64=0 ITERS=1 NDFB_TRACE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3a_launches_${shape}_$v.csv python tools/run_one.py 2> gpurun_out/r3a_trace.txt > /dev/null
  echo "== $shape $v"; grep "^\[ndfb\]" gpurun_out/r3a_trace.txt | sort -u | cut -c1-140
  python - "gpurun_out/r3a_launches_${shape}_$v.csv" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
for r in rows[hdr + 1:]:
    if 'sfft' in r[4]: print('   ', r[4][:100], r[-1], r[-2])
PY
done
