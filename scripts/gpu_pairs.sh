set -o pipefail
BZ_FB_PAIRS=1 timeout 600 python -m pytest tests/test_gpu_prover.py tests/test_gpu_verifier.py tests/test_gpu_params.py -x -q > gpurun_out/pairs_tests.log 2>&1; tail -6 gpurun_out/pairs_tests.log
for pm in 0 1; do
  BZ_FB_PAIRS=$pm timeout 300 python bench.py --no-extras > gpurun_out/pairs_shot_$pm.log 2>&1
  BZ_FB_PAIRS=$pm timeout 300 python bench.py --no-extras --workload board > gpurun_out/pairs_board_$pm.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/pairs_*_?.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            ok=True; d=json.loads(l); print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v/d['steps'],2) for k,v in d['roofline']['kernel_ms'].items()}, d.get('verified'), d.get('single_proof_ms'))
    if not ok: print(f, open(f).read()[-500:])
PY
