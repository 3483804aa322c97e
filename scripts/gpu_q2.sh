set -o pipefail
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_verifier.py -x -q -s > gpurun_out/q2_tests.log 2>&1; tail -3 gpurun_out/q2_tests.log; grep -a "multiplications per proof" gpurun_out/q2_tests.log
timeout 300 python bench.py > gpurun_out/q2_bench.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/q2_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v/d['steps'],2) for k,v in d['roofline']['kernel_ms'].items()}, d.get('verified'), d.get('single_proof_ms'))
        print({k:(v.get('value'),v.get('unit'),v.get('error')) for k,v in d.get('extras',{}).items()})
PY
