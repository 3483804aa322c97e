# single-GPU: MSM / prover / fine-grained / large parity tests after the batched bucket MSM, then k = 18 (forced bucket MSM) and k = 20 proofs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
export BZ_NO_CPU_BASELINE=1
timeout 300 python bench.py --workload msm --steps 3 --warmup 3 > gpurun_out/r2_msm22.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/r2_msm22.log'):
    if l.startswith('{'):
        d=json.loads(l); print('msm 2^22', round(d['value']/1e6,1), 'M pts/s e2e', round(d['e2e']['value']/1e6,1), d['roofline']['kernel_ms'], d['verified'])
PY
tail -2 gpurun_out/r2_msm22.log | cut -c1-300
BZ_FORCE_GENERAL_MSM=1 timeout 600 python bench.py --workload board_scaled --k 18 --steps 2 --warmup 2 > gpurun_out/scaled18g_n1.log 2>&1
timeout 600 python bench.py --workload board_scaled --k 20 --steps 2 --warmup 3 > gpurun_out/scaled20_n1.log 2>&1
python - <<'PY'
import json,glob
for f in ['gpurun_out/scaled18g_n1.log', 'gpurun_out/scaled20_n1.log']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],3), round(d['ms_per_step'],1), d['n_gpus'], d['scaling'], d['roofline']['kernel_ms'] if d['roofline'] else None, d['verified'], d['single_proof_ms'])
PY
tail -3 gpurun_out/scaled20_n1.log | cut -c1-300
