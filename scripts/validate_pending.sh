# Round 2, first GPU call: validate what round 1 wrote after its GPU budget was spent, then A/B it.
#   gpurun --timeout 900 -- 'bash scripts/validate_pending.sh'
set -o pipefail
BZ_VALIDATE_PENDING=1 python -m pytest tests/test_gpu_prover.py tests/test_gpu_arith.py -x -q -k "narrow_geometry or binary_gcd" 2>&1 | tail -3
for nm in 0 1; do
  BZ_GP_FINISH_NARROW=$nm python bench.py --no-extras > gpurun_out/pending_narrow$nm.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/pending_narrow?.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],1), round(d['e2e']['value'],1), {k:round(v/d['steps'],2) for k,v in d['roofline']['kernel_ms'].items()}, d.get('single_proof_ms'))
PY
