mkdir -p gpurun_out
for F in A B; do
  NDFB_SFFT_FAMILY=$F timeout 600 python tools/bench_configs.py --only c1,c2,c3,c4,c5a --iters 10 > gpurun_out/fam_$F.jsonl 2> gpurun_out/fam_$F.err
done
python - <<'PY'
import json
A=[json.loads(l) for l in open('gpurun_out/fam_A.jsonl')]
B=[json.loads(l) for l in open('gpurun_out/fam_B.jsonl')]
for a,b in zip(A,B):
    print(f"{a['cfg']:4s} {a['call'][:48]:48s} A {a['ms']:8.4f} ({a['frac_hbm']:.3f})  B {b['ms']:8.4f} ({b['frac_hbm']:.3f})  B/A {b['ms']/a['ms']:.2f}")
PY
