#!/bin/bash
# factored four-step twiddle in the column kernels (MODE 4): parity, c5b splits with / without it, per-launch times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "four_step or config5b or three_pass or fused_two_pass or config2 or in_place" > gpurun_out/r2w_pytest.log 2>&1; tail -3 gpurun_out/r2w_pytest.log
timeout 900 python tools/exp_c5b2.py > gpurun_out/r2w_c5b_variants.txt 2> gpurun_out/r2w.err; grep -E "^\{" gpurun_out/r2w_c5b_variants.txt
for v in 4096 256 64; do
  export NDFB_FS_N1=$v
  SHAPE=64x16777216 AXIS=1 F64=0 ITERS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2w_c5b_launches_$v.csv python tools/run_one.py > /dev/null 2>&1
  echo "== FS_N1=$v"
  python - "$v" <<'PY'
import csv, sys
rows = list(csv.reader(open('gpurun_out/r2w_c5b_launches_%s.csv' % sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
for r in rows[hdr + 1:]:
    if 'sfft' in r[4]: print('   ', r[4][:100], r[-1], r[-2])
PY
done
unset NDFB_FS_N1
python - <<'PY'
import subprocess, sys, os, json
# medium four-step lengths: factored vs per-point lookups
code = r'''
import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ndrustfft_b200 as nb
for n, b in ((1 << 16, 4096), (1 << 18, 1024), (1 << 20, 256), (1 << 22, 64)):
    x = torch.complex(torch.rand((b, n), device="cuda"), torch.rand((b, n), device="cuda")); y = torch.empty_like(x)
    h = nb.FftHandler(n, np.float32)
    for _ in range(3): nb.ndfft(x, y, h, 1)
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); nb.ndfft(x, y, h, 1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    want = np.fft.fft(x[:1].cpu().numpy().astype(np.complex128), axis=1)
    print(json.dumps({"n": n, "batch": b, "variant": os.environ.get("V"), "ms": round(ts[3], 4), "frac_one_pass": round(2 * x.numel() * 8 / (ts[3] * 1e-3) / 1e9 / 6547.8, 3), "rel_l2": float(np.linalg.norm(y[:1].cpu().numpy() - want) / np.linalg.norm(want))}))
'''
for v, env in (("factored", {}), ("per-point lookups", {"NDFB_NO_FS_FACTORED": "1"})):
    e = dict(os.environ); e.update(env); e["V"] = v
    print(subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True).stdout, end="")
PY
