# single-GPU: prover tests incl. the materialised-G' IPA, then the scaled proofs and the headline bench (RNG mirror change)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_prover.py tests/test_gpu_finegrained.py tests/test_gpu_large_parity.py -x -q 2>&1 | tail -6
export BZ_NO_CPU_BASELINE=1
for k in 16 18 20; do
timeout 600 python bench.py --workload board_scaled --k $k --steps 2 --warmup 3 > gpurun_out/scaled${k}_n1.log 2>&1
done
timeout 600 bash scripts/r2_ab.sh "BZ_X=0"
python - <<'PY'
import json,glob
for f in ['gpurun_out/scaled16_n1.log', 'gpurun_out/scaled18_n1.log', 'gpurun_out/scaled20_n1.log']:
    ok=False
    for l in open(f):
        if l.startswith('{'):
            ok=True
            d=json.loads(l); print(f, round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), d['roofline']['kernel_ms'] if d['roofline'] else None, d['verified'], d['single_proof_ms'])
    if not ok: print(f, [l[:300] for l in open(f)][-4:])
PY
