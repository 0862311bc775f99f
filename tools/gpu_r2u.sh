#!/bin/bash
# c5b: per-launch times of the two-pass and three-pass splits (ncu launch list, one call each)
mkdir -p gpurun_out
for v in default 256 64 512; do
  if [ "$v" = default ]; then unset NDFB_FS_N1; else export NDFB_FS_N1=$v; fi
  SHAPE=64x16777216 AXIS=1 F64=0 ITERS=2 NDFB_TRACE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_c5b_launches_$v.csv python tools/run_one.py 2> gpurun_out/r2u_trace_$v.txt > /dev/null
  echo "== FS_N1=$v"; grep "^\[ndfb\]" gpurun_out/r2u_trace_$v.txt | sort -u | cut -c1-150
  python - "$v" <<'PY'
import csv, sys
rows = list(csv.reader(open('gpurun_out/r2u_c5b_launches_%s.csv' % sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
for r in rows[hdr + 1:]:
    name = r[4]
    if 'sfft' in name or 'fs' in name:
        print('   ', name[:110], r[-1], r[-2])
PY
done
