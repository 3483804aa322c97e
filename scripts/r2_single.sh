# single-GPU: full GPU suite, then bench A/B (generated quotient on/off), MSM / NTT lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash scripts/r2_ab.sh "BZ_X=0" "BZ_QUOTIENT_GENERATED=0"
python bench.py --workload msm --steps 3 --warmup 3 > gpurun_out/r2_msm22.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/r2_msm22.log'):
    if l.startswith('{'):
        d=json.loads(l); print('msm 2^22', round(d['value']/1e6,1), 'M pts/s', d['roofline']['kernel_ms'])
PY
