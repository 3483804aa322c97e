# single-GPU: full GPU suite after the sliced IPA inner product, then the k = 20 proof
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
export BZ_NO_CPU_BASELINE=1
timeout 600 python bench.py --workload board_scaled --k 20 --steps 2 --warmup 3 > gpurun_out/scaled20_n1.log 2>&1
python - <<'PY'
import json
for f in ['gpurun_out/scaled20_n1.log']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), d['roofline']['kernel_ms'] if d['roofline'] else None, d['verified'], d['single_proof_ms'])
PY
