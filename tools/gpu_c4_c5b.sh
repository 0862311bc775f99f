mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python tools/bench_configs.py --only c4 --iters 10 | tee gpurun_out/c4_r1n.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['call'][:40], d['ms'], d['frac_hbm'])"
for N1 in 256 2048 4096; do echo "N1=$N1"; NDFB_FS_N1=$N1 NDFB_TRACE=1 python tools/bench_configs.py --only c5b --iters 3 2>gpurun_out/c5b_trace_$N1.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['ms'], d['frac_hbm'])"; grep four-step gpurun_out/c5b_trace_$N1.txt | sort | uniq -c | head -3; done
