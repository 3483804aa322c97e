set -o pipefail
BZ_FIXED_WINDOW=16 python bench.py --no-extras > gpurun_out/win_shot_c16.log 2>&1
BZ_FIXED_WINDOW=15 python bench.py --no-extras --workload board > gpurun_out/win_board_c15.log 2>&1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > gpurun_out/win_mem.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/win_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v/d['steps'],2) for k,v in d['roofline']['kernel_ms'].items()}, d.get('verified'))
    if 'log' in f: print(open(f).read()[-300:] if not any(l.startswith('{') for l in open(f)) else '')
PY
bash profiles/capture.sh r1g > gpurun_out/r1g_capture.log 2>&1; tail -3 gpurun_out/r1g_capture.log
