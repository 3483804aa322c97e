# batch x lanes sweep of the Shot workload (and Board at the best few), one bench line per point
set -o pipefail
for b in 64 128 256; do for fl in 2 3 4 6; do
  python bench.py --no-extras --batch $b --inflight $fl --steps 4 --warmup 3 > gpurun_out/sweep_shot_b${b}_l${fl}.log 2>&1
done; done
for b in 64 128; do for fl in 2 4; do
  python bench.py --no-extras --workload board --batch $b --inflight $fl --steps 4 --warmup 3 > gpurun_out/sweep_board_b${b}_l${fl}.log 2>&1
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sweep_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d.get('single_proof_ms'))
PY
