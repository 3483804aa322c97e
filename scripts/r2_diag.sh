# where a k = 20 proof's wall time goes (host phases with the device drained at the boundaries), and the witness-like MSM line
mkdir -p gpurun_out
export BZ_NO_CPU_BASELINE=1
BZ_PHASE_TIMES=1 timeout 600 python bench.py --workload board_scaled --k 20 --steps 1 --warmup 1 > gpurun_out/k20_phases.log 2> gpurun_out/k20_phases.err
grep "^\[phase\]" gpurun_out/k20_phases.err | tail -7
for L in 20 22; do
timeout 300 python bench.py --workload msm --scalars witness --log2n $L --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/msm_w$L.log
done
python - <<'PY'
import json
for L in (20,22):
    for l in open(f'gpurun_out/msm_w{L}.log'):
        if l.startswith('{'):
            d=json.loads(l); print('msm witness 2^%d'%L, round(d['value']/1e6,1), 'M pts/s e2e', round(d['e2e']['value']/1e6,1), d['roofline']['kernel_ms'])
PY
