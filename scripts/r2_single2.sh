# single-GPU: full GPU suite, A/B of the 128-register build of the generated h(X) kernels, MSM line, k = 20 proof, ncu evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
export BZ_NO_CPU_BASELINE=1
timeout 600 bash scripts/r2_ab.sh "BZ_X=0" "BZ_LIB=battlezips-halo2_b200/lib_ab/libbzhalo2.so"
timeout 300 python bench.py --workload msm --steps 3 --warmup 3 > gpurun_out/r2_msm22.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/r2_msm22.log'):
    if l.startswith('{'):
        d=json.loads(l); print('msm 2^22', round(d['value']/1e6,1), 'M pts/s', d['roofline']['kernel_ms'])
PY
for k in 18 20; do
timeout 600 python bench.py --workload board_scaled --k $k --steps 2 --warmup 3 > gpurun_out/scaled${k}_n1.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scaled*_n1.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value'],3), round(d['ms_per_step'],1), d['n_gpus'], d['scaling'], d['roofline']['kernel_ms'] if d['roofline'] else None, d['verified'], d['single_proof_ms'])
PY
unset BZ_NO_CPU_BASELINE
timeout 1200 bash profiles/capture.sh r2 2>&1 | tail -20
