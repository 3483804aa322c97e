# N-GPU run (N = $1): sharded-proof parity test, then the scaled proof at world N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
grep -v "^\[W\|^W1018\|^\*\*\*" gpurun_out/multi_gpu_worker.log 2>/dev/null | tail -3 | cut -c1-300
export BZ_NO_CPU_BASELINE=1
for k in ${KS:-16 18 20}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload board_scaled --k $k --steps 2 --warmup 3 > gpurun_out/scaled${k}_n$N.log 2>&1
done
python - $N <<'PY'
import json,glob,sys
N=sys.argv[1]
for f in sorted(glob.glob(f'gpurun_out/scaled*_n{N}.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            ok=True
            d=json.loads(l); print(f, round(d['value'],3), round(d['ms_per_step'],1), d['n_gpus'], d['scaling'], d['roofline']['kernel_ms'] if d['roofline'] else None, d['verified'], d['single_proof_ms'])
    if not ok: print(f, 'NO LINE:', [l[:300] for l in open(f) if 'Error' in l or 'error' in l][-3:])
PY
