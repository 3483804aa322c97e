# A/B of env-var switches on the headline bench (no extras).  usage: bash scripts/r2_ab.sh "NAME=VAL ..." "NAME=VAL ..." ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --no-extras > gpurun_out/ab_$i.log 2>&1
  python - "$cfg" gpurun_out/ab_$i.log <<'PY'
import json,sys
for l in open(sys.argv[2]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], '| value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,2) for k,v in d['roofline']['kernel_ms'].items()}, 'single', d.get('single_proof_ms'))
PY
done
