set -o pipefail
python -m pytest tests/test_gpu_prover.py -x -q -s -k "quotient_dag or shot_proof or board_proof or goldens or tiny" > gpurun_out/s41_tests.log 2>&1; tail -5 gpurun_out/s41_tests.log
for cse in 1 0; do
  BZ_QUOTIENT_CSE=$cse python bench.py --no-extras > gpurun_out/s41_shot_cse$cse.log 2>&1
  BZ_QUOTIENT_CSE=$cse python bench.py --no-extras --workload board > gpurun_out/s41_board_cse$cse.log 2>&1
done
for fl in 6 8; do python bench.py --no-extras --inflight $fl > gpurun_out/s41_shot_fl$fl.log 2>&1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s41_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), {k:round(v/d['steps'],2) for k,v in d['roofline']['kernel_ms'].items()}, d.get('verified'))
PY
