set -o pipefail
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_verifier.py -x -q > gpurun_out/gp_tests.log 2>&1; tail -5 gpurun_out/gp_tests.log
for sq in 1 0; do
  BZ_GP_SEQUENTIAL=$sq timeout 300 python bench.py --no-extras > gpurun_out/gp_shot_seq$sq.log 2>&1
  BZ_GP_SEQUENTIAL=$sq timeout 300 python bench.py --no-extras --workload board > gpurun_out/gp_board_seq$sq.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/gp_*_seq?.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            ok=True; d=json.loads(l); print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v/d['steps'],2) for k,v in d['roofline']['kernel_ms'].items()}, d.get('verified'), d.get('single_proof_ms'))
    if not ok: print(f, open(f).read()[-500:])
PY
