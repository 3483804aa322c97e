# headline bench with the lanes free-running across the K timed steps
mkdir -p gpurun_out
export BZ_NO_CPU_BASELINE=1
timeout 200 python bench.py --no-extras > gpurun_out/last_shot.log 2>&1
timeout 200 python bench.py --workload board --no-extras > gpurun_out/last_board.log 2>&1
python - <<'PY'
import json
for f in ['gpurun_out/last_shot.log','gpurun_out/last_board.log']:
    ok=False
    for l in open(f):
        if l.startswith('{'):
            ok=True
            d=json.loads(l); print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), d['verified'], d['verify'] and d['verify']['accepted'], d['gpu_launches'], d['clocks'])
    if not ok: print(f, open(f).read()[-1500:])
PY
