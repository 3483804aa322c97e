# single-GPU: full GPU suite after the two-level table-MSM fold, headline bench, scaled proofs on the table path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
export BZ_NO_CPU_BASELINE=1
timeout 600 bash scripts/r2_ab.sh "BZ_X=0"
for k in 16 18; do
timeout 600 python bench.py --workload board_scaled --k $k --steps 2 --warmup 3 > gpurun_out/scaled${k}_n1.log 2>&1
done
timeout 600 python bench.py --workload board --no-extras > gpurun_out/board_n1.log 2>&1
python - <<'PY'
import json,glob
for f in ['gpurun_out/scaled16_n1.log', 'gpurun_out/scaled18_n1.log', 'gpurun_out/board_n1.log']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), d['roofline']['kernel_ms'] if d['roofline'] else None, d['verified'], d['single_proof_ms'])
PY
