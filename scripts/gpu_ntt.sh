for lg in 21 22 23 24; do for tp in 20 24; do
  BZ_NTT_2PASS_MAX=$tp python bench.py --workload ntt --log2n $lg --steps 10 --warmup 3 > gpurun_out/ntt_${lg}_tp${tp}.log 2>&1
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ntt_*_tp*.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True; print(f, round(d['value'],1), d['unit'], round(d['ms_per_step'],3), d.get('verified'))
    if not ok: print(f, open(f).read()[-400:])
PY
python -m pytest tests/test_gpu_arith.py -q -x -k "ntt or fft" 2>&1 | tail -2
BZ_NTT_2PASS_MAX=24 python -m pytest tests/test_gpu_arith.py -q -x -k "ntt or fft" 2>&1 | tail -2
